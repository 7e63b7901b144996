#!/bin/bash
mkdir -p gpurun_out
timeout 285 python bench.py --size 20M --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_20M_ebe.json 2> gpurun_out/bench_20M_ebe.err; echo "rc=$?"
tail -3 gpurun_out/bench_20M_ebe.err
python -c "
import json
d=json.load(open('gpurun_out/bench_20M_ebe.json')); print(d['value'], d['ms_per_step'], d['config']['iterations'], d['config']['max_relres'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['kernel'][:12])"
