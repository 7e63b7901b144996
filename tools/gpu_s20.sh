#!/bin/bash
# round-2 session 20: element-wise product on order-3 triangles (the reference's default 2D configuration)
mkdir -p gpurun_out
L=gpurun_out/s20.log
: > $L
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/s20_pytest.log 2>&1; echo "pytest parity rc=$?" >> $L; tail -4 gpurun_out/s20_pytest.log >> $L
rm -f gpurun_out/golden_stats.json
REMO_GOLDEN_STATS=gpurun_out/golden_stats.json timeout 1200 python -m pytest tests/test_gpu_reference_logs.py tests/test_gpu_golden_example01.py -q > gpurun_out/s20_pytest2.log 2>&1; echo "pytest reference logs rc=$?" >> $L; tail -3 gpurun_out/s20_pytest2.log >> $L
python tests/sanitize_cases.py >> $L 2>&1
for ebe in 1 0; do
  echo "== example01, REMO_SPMM_EBE=$ebe" >> $L
  REMO_SPMM_EBE=$ebe timeout 600 python bench.py --mode example01 > gpurun_out/s20_example01_ebe$ebe.json 2> gpurun_out/s20_example01_ebe$ebe.err; echo "rc=$?" >> $L; cut -c1-130 gpurun_out/s20_example01_ebe$ebe.json >> $L
  python - $ebe >> $L <<PY
import json, sys
d = json.load(open('gpurun_out/s20_example01_ebe%s.json' % sys.argv[1]))
print('  busy', round(d['config']['gpu_busy_fraction'], 3), 'iters', d['config']['iterations_median'], 'parity', d['parity']['max_rel_err_ra'])
PY
done
echo "== product time on a 2D mesh (order 3)" >> $L
python - >> $L 2>&1 <<PY
import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import helpers
from remo3d_b200 import _cabi
m2, s2, f2, _ = helpers.disc_case(h_electrode=0.01, h_axis=0.06, h_borehole=0.1, grading=0.4)
for ebe in (1, 0):
    ctx = _cabi.Context(0); ctx.set_option("spmm_ebe", ebe)
    ctx.mesh_set(2, m2.points, m2.elems, m2.mat, m2.bfacets, m2.dirichlet_flags([2]), m2.axis_vertices())
    ndof, _ = ctx.space_build(3); ctx.assemble(np.asarray(s2, float)); ctx.precond_setup("multigrid")
    ctx.rhs_point_sources(f2["src_ptr"], f2["src_z"], f2["src_fac"])
    it, rr = ctx.solve(rtol=1e-10, maxit=5000)
    k = f2["src_ptr"].shape[0] - 1
    ms = ctx.kernel_time(0, k, 50)
    print('ebe', ebe, 'ndof', ndof, 'kind', ctx.spmm_kind(), 'k', k, 'product ms', round(ms, 4), 'iters', it.tolist(), 'solve ms', round(ctx.stage_times()['solve'], 2))
    ctx.close()
PY
cat $L
