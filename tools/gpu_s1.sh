#!/bin/bash
# round-2 session 1: correctness of the new solver core + first A/Bs
mkdir -p gpurun_out
L=gpurun_out/s1.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -15 gpurun_out/s1_pytest.log >> $L
for c in 1 0; do
  echo "== spmm_probe 5M REMO_EBE_COLOR=$c" >> $L
  REMO_EBE_COLOR=$c REMO_PROBE_SIZE=5M timeout 300 python tools/spmm_probe.py --ks 1,5 >> $L 2>&1
done
run() { # name, env opts
  echo "== bench $1 ($2)" >> $L
  env $2 timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/s1_$1.json 2> gpurun_out/s1_$1.err; echo "rc=$?" >> $L
  python - >> $L 2>&1 <<PY
import json
d = json.load(open('gpurun_out/s1_$1.json'))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'frac', round(d['roofline']['frac'],3), 'spmm ms', round(d['roofline']['avg_launch_ms'],4), 'share', round(d['roofline']['spmm_share_of_step'],3))
print('stages', {k: round(v,2) for k,v in d['config']['stage_ms_one_context_alone'].items()}, 'levels', d['config'].get('amg_levels'))
PY
  tail -3 gpurun_out/s1_$1.err >> $L
}
run default "A=1"
run morton "REMO_BENCH_OPTS=amg_agg=0"
run pass2 "REMO_BENCH_OPTS=amg_passes=2"
run nocolor "REMO_EBE_COLOR=0"
cat $L
