#!/bin/bash
# round-2 session 14: full GPU suite with the order-3 element-wise product, default bench with the order-3 companion,
# ncu full capture of the order-3 product, Model in 3D at the reference's order 3
mkdir -p gpurun_out
L=gpurun_out/s14.log
: > $L
rm -f gpurun_out/golden_stats.json
REMO_GOLDEN_STATS=gpurun_out/golden_stats.json timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s14_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -6 gpurun_out/s14_pytest.log >> $L
echo "== default bench" >> $L
timeout 1200 python bench.py --steps 12 --warmup 3 > gpurun_out/s14_bench_default.json 2> gpurun_out/s14_bench_default.err; echo "rc=$?" >> $L
tail -3 gpurun_out/s14_bench_default.err >> $L
python - >> $L 2>&1 <<PY
import json
d = json.load(open('gpurun_out/s14_bench_default.json'))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'frac', round(d['roofline']['frac'],3), 'spmm ms', round(d['roofline']['avg_launch_ms'],4))
print('order3', d['config'].get('order3_companion'))
print('plain', d['config'].get('value_plain_mesh')); print('parity', d.get('parity')); print('like', d.get('like_for_like'))
PY
echo "== ncu order-3 product" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_ebe -s 30 -c 2 -o gpurun_out/r02_ebe_p3_k5_1M -f \
  python tools/spmm_probe.py --size 1M --order 3 --ks 5 > gpurun_out/s14_ncu_p3.log 2>&1; echo "ncu full rc=$?" >> $L
cat $L
