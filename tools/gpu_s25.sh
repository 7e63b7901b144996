#!/bin/bash
# round-2 session 25: last A/B of the cycle's free parameters on the final code (12 steps each)
mkdir -p gpurun_out
L=gpurun_out/s25.log
: > $L
for o in amg_fp32=1 amg_passes=2 amg_sweeps=2 amg_alpha=1.3 amg_alpha=1.8 amg_rounds=2; do
  REMO_BENCH_OPTS=$o timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/s25_$o.json 2> /dev/null
  python - $o >> $L <<PY
import json, sys
d = json.load(open('gpurun_out/s25_%s.json' % sys.argv[1]))
print('%-16s value %.2f ms/step %.1f iters %s setup %.1f levels %s' % (sys.argv[1], d['value'], d['ms_per_step'], d['config']['iterations'], d['config']['stage_ms_one_context_alone']['precond_setup'], [l[0] for l in d['config']['amg_levels']]))
PY
done
cat $L
