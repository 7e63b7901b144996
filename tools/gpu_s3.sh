#!/bin/bash
# round-2 session 3: tests after the fixes, fused tail A/B, 1M morton check
mkdir -p gpurun_out
L=gpurun_out/s3.log
: > $L
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -8 gpurun_out/s3_pytest.log >> $L
run() { # name, size, env opts
  echo "== bench $1 $2 ($3)" >> $L
  env $3 timeout 600 python bench.py --size $2 --steps 4 --warmup 2 --no-cpu-baseline --no-companions > gpurun_out/s3_$1.json 2> gpurun_out/s3_$1.err; echo "rc=$?" >> $L
  python - >> $L 2>&1 <<PY
import json
d = json.load(open('gpurun_out/s3_$1.json'))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'frac', round(d['roofline']['frac'],3), 'spmm ms', round(d['roofline']['avg_launch_ms'],4), 'share', round(d['roofline']['spmm_share_of_step'],3))
print('stages', {k: round(v,2) for k,v in d['config']['stage_ms_one_context_alone'].items()}, 'levels', d['config'].get('amg_levels'))
PY
  tail -2 gpurun_out/s3_$1.err >> $L
}
run default 5M "A=1"
run unfused 5M "REMO_BENCH_OPTS=amg_fused_tail=0"
run tail100k 5M "REMO_BENCH_OPTS=amg_tail_rows=100000"
run morton1M 1M "REMO_BENCH_OPTS=amg_agg=0"
run pair1M 1M "A=1"
cat $L
