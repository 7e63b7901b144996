#!/bin/bash
# round-2 session 9: contexts per GPU with the 8-lane smoother (default on), fp64 cycle for comparison
mkdir -p gpurun_out
L=gpurun_out/s9.log
: > $L
run() {  # name, contexts, opts
  name=$1; ctx=$2; opts=$3
  echo "== bench $name (contexts $ctx, $opts)" >> $L
  REMO_BENCH_OPTS=$opts timeout 900 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-companions --contexts $ctx > gpurun_out/s9_$name.json 2> gpurun_out/s9_$name.err; echo "rc=$?" >> $L
  python - $name >> $L 2>&1 <<PY
import json, sys
d = json.load(open('gpurun_out/s9_%s.json' % sys.argv[1]))
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'iters', d['config']['iterations'],
      'spmm ms', round(d['roofline']['avg_launch_ms'],4), 'solve', round(d['config']['stage_ms_one_context_alone']['solve'],1))
PY
}
run c2 2 amg_lanes8=1
run c3 3 amg_lanes8=1
run c4 4 amg_lanes8=1
run c2_fp64 2 amg_fp32=0
run c3_fp64 3 amg_fp32=0
run c2_l4 2 amg_lanes8=0
cat $L
