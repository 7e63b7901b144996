#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3; do
REMO_BENCH_DEBUG=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --contexts $c > gpurun_out/bench_5M_c$c.json 2> gpurun_out/bench_5M_c$c.err; tail -2 gpurun_out/bench_5M_c$c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_5M_c$c.json')); print('contexts', $c, 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['stage_ms'])"
done
