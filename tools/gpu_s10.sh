#!/bin/bash
# round-2 session 10: full GPU suite, default bench line (all legs), pipeline mode with and without per-task interfaces
mkdir -p gpurun_out
L=gpurun_out/s10.log
: > $L
rm -f gpurun_out/golden_stats.json
REMO_GOLDEN_STATS=gpurun_out/golden_stats.json timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -6 gpurun_out/s10_pytest.log >> $L
echo "== default bench" >> $L
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/s10_bench_default.json 2> gpurun_out/s10_bench_default.err; echo "rc=$?" >> $L
tail -3 gpurun_out/s10_bench_default.err >> $L
python - >> $L 2>&1 <<PY
import json
d = json.load(open('gpurun_out/s10_bench_default.json'))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'frac', round(d['roofline']['frac'],3), 'spmm ms', round(d['roofline']['avg_launch_ms'],4), 'share', round(d['roofline']['spmm_share_of_step'],3))
print('stages', {k: round(v,2) for k,v in d['config']['stage_ms_one_context_alone'].items()})
print('plain', d['config'].get('value_plain_mesh')); print('parity', d.get('parity')); print('like', d.get('like_for_like')); print('cpu', d['cpu_baseline']['value'])
PY
for c in 0 1; do
  echo "== pipeline conforming=$c" >> $L
  timeout 1200 python bench.py --mode pipeline --pipeline-conforming $c > gpurun_out/s10_pipeline_c$c.json 2> gpurun_out/s10_pipeline_c$c.err; echo "rc=$?" >> $L; tail -2 gpurun_out/s10_pipeline_c$c.err >> $L; cat gpurun_out/s10_pipeline_c$c.json >> $L
done
cat $L
