#!/bin/bash
# round-2 session 6: mixed-precision V-cycle A/B, table validator cases, parity tests
mkdir -p gpurun_out
L=gpurun_out/s6.log
: > $L
bash tools/sanitize.sh > /dev/null 2>&1; cat gpurun_out/sanitize_summary.log >> $L
timeout 1500 python -m pytest tests -m gpu -q -x -k "not reference_logs" > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -5 gpurun_out/s6_pytest.log >> $L
run() {  # name, env...
  name=$1; shift
  echo "== bench $name ($*)" >> $L
  env "$@" timeout 900 python bench.py --no-cpu-baseline --no-companions > gpurun_out/s6_$name.json 2> gpurun_out/s6_$name.err; echo "rc=$?" >> $L
  python - $name >> $L 2>&1 <<PY
import json, sys
d = json.load(open('gpurun_out/s6_%s.json' % sys.argv[1]))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'frac', round(d['roofline']['frac'],3), 'spmm ms', round(d['roofline']['avg_launch_ms'],4), 'share', round(d['roofline']['spmm_share_of_step'],3))
print('stages', {k: round(v,2) for k,v in d['config']['stage_ms_one_context_alone'].items()}, 'parity', d.get('parity'))
PY
}
run fp32 A=1
run fp64 REMO_BENCH_OPTS=amg_fp32=0
echo "== (1M size skipped)" >> $L
cat $L
