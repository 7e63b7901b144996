#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep10.log
run() { echo "$*" >> gpurun_out/sweep10.log; env "$@" timeout 300 python tools/spmm_probe.py --ks 5,8 2>&1 | grep "^k=\|rror" >> gpurun_out/sweep10.log; }
for w in 0 2 3 4; do run REMO_SELL_WIDE=$w REMO_PROBE_SIZE=5M; done
for w in 0 3; do run REMO_SELL_WIDE=$w REMO_PROBE_SIZE=1M; done
cat gpurun_out/sweep10.log
REMO_SELL_WIDE=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmm or block or rhs" 2>&1 | tail -3
