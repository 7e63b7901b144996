#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/ebe5.log
: > $L
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmm or block_sizes" 2>&1 | tail -6 >> $L
REMO_PROBE_SIZE=5M timeout 400 python tools/spmm_probe.py --ks 5,6,8,2,1 2>&1 | grep "^k=\|rror\|ndof" >> $L
REMO_PROBE_SIZE=1M timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_spmm_ebe -s 30 -c 1 -o gpurun_out/ebe5_k5_1M -f \
  python tools/spmm_probe.py --ks 5 > gpurun_out/ncu_ebe5.log 2>&1
grep "^k=" gpurun_out/ncu_ebe5.log >> $L
cat $L
