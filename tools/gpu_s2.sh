#!/bin/bash
# round-2 session 2: full GPU tests, launch list of one 5M step, Morton-aggregate iteration check at 1M
mkdir -p gpurun_out
L=gpurun_out/s2.log
: > $L
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -8 gpurun_out/s2_pytest.log >> $L
for o in "amg_agg=0" "amg_agg=1"; do
  echo "== bench 1M $o" >> $L
  REMO_BENCH_OPTS=$o timeout 300 python bench.py --size 1M --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/s2_1M.json 2> gpurun_out/s2_1M.err; echo "rc=$?" >> $L
  python -c "
import json
d = json.load(open('gpurun_out/s2_1M.json'))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'iters', d['config']['iterations'], 'stages', {k: round(v,2) for k,v in d['config']['stage_ms_one_context_alone'].items()}, d['config']['amg_levels'])" >> $L 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/s2_launches_5M.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --contexts 1 > gpurun_out/s2_ncu_launch.log 2>&1; echo "launch list rc=$?" >> $L
python tools/summarize_launches.py gpurun_out/s2_launches_5M.csv gpurun_out/s2_launches_5M_summary.csv >> $L 2>&1
gzip -f gpurun_out/s2_launches_5M.csv
cat $L
