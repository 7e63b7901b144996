#!/bin/bash
mkdir -p gpurun_out
export REMO_MESH_CACHE=$PWD/.meshcache
timeout 900 python bench.py --size 20M --steps 2 --warmup 1 --no-cpu-baseline --maxit 5000 > gpurun_out/bench_20M.json 2> gpurun_out/bench_20M.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/bench_20M.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['ndof'], d['config']['nnz'], d['config']['iterations'], d['config']['stage_ms_one_context_alone'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
nvidia-smi --query-gpu=memory.used --format=csv,noheader
