#!/bin/bash
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 4 --warmup 2 --no-cpu-baseline --contexts 1 > gpurun_out/bench_5M_pat1.json 2> gpurun_out/bench_5M_pat1.err
python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_5M_pat2.json 2> gpurun_out/bench_5M_pat2.err
python -c "
import json
for f in ('pat1','pat2'):
    d=json.load(open('gpurun_out/bench_5M_%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['config']['stage_ms_one_context_alone'])"
