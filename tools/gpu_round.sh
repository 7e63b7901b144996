#!/bin/bash
# The standard evidence round (one GPU): GPU test-suite with the reference-log statistics, smoke, the default bench line with all
# its legs, the reference arm, pipeline + example01 modes, the ncu launch list of one step and the full captures of the two
# product kernels.  Outputs under gpurun_out/ (scratch); the summaries worth keeping are copied to profiles/ by hand.
mkdir -p gpurun_out
T=${1:-round}
L=gpurun_out/${T}.log
: > $L
rm -f gpurun_out/golden_stats.json
REMO_GOLDEN_STATS=gpurun_out/golden_stats.json timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -4 gpurun_out/${T}_pytest.log >> $L
echo "== smoke" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
echo "== default bench" >> $L
timeout 1200 python bench.py --steps 12 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "rc=$?" >> $L
tail -3 gpurun_out/${T}_bench_default.err >> $L
python - $T >> $L 2>&1 <<PY
import json, sys
d = json.load(open('gpurun_out/%s_bench_default.json' % sys.argv[1]))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'frac', round(d['roofline']['frac'],3), 'spmm ms', round(d['roofline']['avg_launch_ms'],4), 'traffic', d['roofline']['traffic'])
o = d['config'].get('order3_companion') or {}
print('order3', {k: o.get(k) for k in ('value', 'ms_per_step', 'iterations', 'ndof', 'product_avg_launch_ms', 'product_frac_of_measured_hbm')})
print('plain', d['config'].get('value_plain_mesh')); print('parity', d.get('parity')); print('like', d.get('like_for_like')); print('cpu', d['cpu_baseline']['value'])
PY
echo "== reference arm" >> $L
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "rc=$?" >> $L; tail -1 gpurun_out/${T}_bench_reference.err >> $L
echo "== pipeline / example01" >> $L
timeout 900 python bench.py --mode pipeline > gpurun_out/${T}_pipeline.json 2> gpurun_out/${T}_pipeline.err; echo "rc=$?" >> $L; cut -c1-260 gpurun_out/${T}_pipeline.json >> $L
timeout 600 python bench.py --mode example01 > gpurun_out/${T}_example01.json 2> gpurun_out/${T}_example01.err; echo "rc=$?" >> $L; cut -c1-260 gpurun_out/${T}_example01.json >> $L
echo "== ncu" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches_bench_5M.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-companions --contexts 1 > gpurun_out/${T}_ncu_launch.log 2>&1; echo "launch list rc=$?" >> $L
python tools/summarize_launches.py gpurun_out/${T}_launches_bench_5M.csv gpurun_out/${T}_launches_bench_5M_summary.csv >> $L 2>&1
gzip -f gpurun_out/${T}_launches_bench_5M.csv
REMO_PROBE_SIZE=5M timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_ebe -s 30 -c 2 -o gpurun_out/${T}_ebe_k5_5M -f \
  python tools/spmm_probe.py --ks 5 > gpurun_out/${T}_ncu_ebe.log 2>&1; echo "ncu full (order 2) rc=$?" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_ebe -s 30 -c 2 -o gpurun_out/${T}_ebe_p3_k5_1M -f \
  python tools/spmm_probe.py --size 1M --order 3 --ks 5 > gpurun_out/${T}_ncu_ebe_p3.log 2>&1; echo "ncu full (order 3) rc=$?" >> $L
cat $L
