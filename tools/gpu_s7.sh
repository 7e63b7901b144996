#!/bin/bash
# round-2 session 7: V-cycle variants (fused tail in fp32, 8 lanes per row) x contexts per GPU
mkdir -p gpurun_out
L=gpurun_out/s7.log
: > $L
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or c3" > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -3 gpurun_out/s7_pytest.log >> $L
run() {  # name, contexts, opts
  name=$1; ctx=$2; opts=$3
  echo "== bench $name (contexts $ctx, $opts)" >> $L
  REMO_BENCH_OPTS=$opts timeout 900 python bench.py --no-cpu-baseline --no-companions --contexts $ctx > gpurun_out/s7_$name.json 2> gpurun_out/s7_$name.err; echo "rc=$?" >> $L
  python - $name >> $L 2>&1 <<PY
import json, sys
d = json.load(open('gpurun_out/s7_%s.json' % sys.argv[1]))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'spmm ms', round(d['roofline']['avg_launch_ms'],4), 'solve', round(d['config']['stage_ms_one_context_alone']['solve'],1))
PY
}
run c2_base 2 amg_fp32=1
run c1_base 1 amg_fp32=1
run c1_tail 1 amg_fused_tail=1
run c1_tail_l8 1 amg_fused_tail=1,amg_lanes8=1
run c2_tail_l8 2 amg_fused_tail=1,amg_lanes8=1
run c1_l8 1 amg_lanes8=1
cat $L
