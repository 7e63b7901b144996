#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
REMO_BENCH_DEBUG=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_5M_flat.json 2> gpurun_out/bench_5M_flat.err; tail -3 gpurun_out/bench_5M_flat.err
REMO_PROBE_SIZE=5M python tools/spmm_probe.py --ks 1,2,5,8 2>&1 | grep "^k="
