#!/bin/bash
# round-2 session 13: order-3 element-wise product: parity tests, product time against the SELL kernel (order 3, 1M-size mesh)
mkdir -p gpurun_out
L=gpurun_out/s13.log
: > $L
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_c3.py -q -x > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -15 gpurun_out/s13_pytest.log >> $L
for ebe in 1 0; do
  echo "== order 3, REMO_SPMM_EBE=$ebe, size 1M mesh" >> $L
  REMO_SPMM_EBE=$ebe timeout 600 python tools/spmm_probe.py --size 1M --order 3 --ks 1,2,5,6 >> $L 2>&1
done
echo "== order 2 (unchanged kernel) size 5M" >> $L
timeout 600 python tools/spmm_probe.py --size 5M --order 2 --ks 5 >> $L 2>&1
cat $L
