#!/bin/bash
# round-2 session 4: all GPU tests (with the reference-log statistics), the default bench line, pipeline mode, ncu evidence
mkdir -p gpurun_out
L=gpurun_out/s4.log
: > $L
rm -f gpurun_out/golden_stats.json
REMO_GOLDEN_STATS=gpurun_out/golden_stats.json timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -12 gpurun_out/s4_pytest.log >> $L
echo "== default bench" >> $L
timeout 900 python bench.py > gpurun_out/s4_bench_default.json 2> gpurun_out/s4_bench_default.err; echo "rc=$?" >> $L
tail -4 gpurun_out/s4_bench_default.err >> $L
python - >> $L 2>&1 <<PY
import json
d = json.load(open('gpurun_out/s4_bench_default.json'))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'frac', round(d['roofline']['frac'],3), 'spmm ms', round(d['roofline']['avg_launch_ms'],4), 'share', round(d['roofline']['spmm_share_of_step'],3))
print('stages', {k: round(v,2) for k,v in d['config']['stage_ms_one_context_alone'].items()}, 'levels', d['config'].get('amg_levels'))
print('plain', d['config'].get('value_plain_mesh')); print('parity', d.get('parity')); print('like', d.get('like_for_like')); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['exponent_measured'])
print('assembly', d.get('assembly'))
PY
echo "== reference arm" >> $L
timeout 600 python bench.py --impl reference > gpurun_out/s4_bench_reference.json 2> gpurun_out/s4_bench_reference.err; echo "rc=$?" >> $L; tail -2 gpurun_out/s4_bench_reference.err >> $L
echo "== pipeline" >> $L
timeout 900 python bench.py --mode pipeline > gpurun_out/s4_pipeline.json 2> gpurun_out/s4_pipeline.err; echo "rc=$?" >> $L; tail -3 gpurun_out/s4_pipeline.err >> $L; cat gpurun_out/s4_pipeline.json >> $L
echo "== ncu" >> $L
REMO_PROBE_SIZE=5M timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_ebe -s 30 -c 2 -o gpurun_out/r02_ebe_k5_5M -f \
  python tools/spmm_probe.py --ks 5 > gpurun_out/s4_ncu_ebe.log 2>&1; echo "ncu full rc=$?" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench_5M.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-companions --contexts 1 > gpurun_out/s4_ncu_launch.log 2>&1; echo "launch list rc=$?" >> $L
python tools/summarize_launches.py gpurun_out/r02_launches_bench_5M.csv gpurun_out/r02_launches_bench_5M_summary.csv >> $L 2>&1
gzip -f gpurun_out/r02_launches_bench_5M.csv
cat $L
