#!/bin/bash
# round-2 session 15: order-3 product at 3 vs 4 resident CTAs per SM; order-3 full solves
mkdir -p gpurun_out
L=gpurun_out/s15.log
: > $L
for o in 3 4; do
  echo "== order 3, ebe_p3_ctas=$o" >> $L
  REMO_BENCH_OPTS=ebe_p3_ctas=$o timeout 600 python tools/spmm_probe.py --size 1M --order 3 --ks 1,2,3,4,5,6 >> $L 2>&1
done
cat $L
