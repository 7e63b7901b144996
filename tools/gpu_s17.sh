#!/bin/bash
# round-2 session 17: 3D golden tests (orders 2 and 3), C5 size class (20 M dofs, plain Delaunay mesh) with the round-2 solver
mkdir -p gpurun_out
L=gpurun_out/s17.log
: > $L
timeout 900 python -m pytest tests/test_gpu_golden_example01.py -q -s > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?" >> $L; grep -E "reference|passed|failed" gpurun_out/s17_pytest.log | tail -12 >> $L
echo "== 20M" >> $L
REMO_BENCH_MESH_IMPROVE=0 timeout 1500 python bench.py --size 20M --steps 3 --warmup 1 --no-cpu-baseline --no-companions > gpurun_out/s17_bench_20M.json 2> gpurun_out/s17_bench_20M.err; echo "rc=$?" >> $L
tail -3 gpurun_out/s17_bench_20M.err >> $L
python - >> $L 2>&1 <<PY
import json
d = json.load(open('gpurun_out/s17_bench_20M.json'))
print('value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), 'iters', d['config']['iterations'], 'ndof', d['config']['ndof'],
      'frac', round(d['roofline']['frac'],3), 'spmm ms', round(d['roofline']['avg_launch_ms'],4))
print('stages', {k: round(v,2) for k,v in d['config']['stage_ms_one_context_alone'].items()})
PY
cat $L
