#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/ebe8.log
: > $L
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmm or block_sizes" 2>&1 | tail -4 >> $L
for sp in 8 16; do
echo "SPLIT=$sp" >> $L
REMO_EBE_SPLIT=$sp REMO_PROBE_SIZE=5M timeout 400 python tools/spmm_probe.py --ks 5,6,2,1 2>&1 | grep "^k=\|rror" >> $L
done
cat $L
