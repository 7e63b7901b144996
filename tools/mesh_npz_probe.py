"""GPU probe on a pre-generated bench mesh (npz of bench.make_mesh): PCG iterations and solve time, ctypes only."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remo3d_b200 import _cabi, planner, tools as tl

TOOLS = ["A2.0M0.5N", "N0.5M2.0A", "B5.7A0.4M", "M4.0A0.5B"]
SIGMA = [1 / 1.0, 1 / 10.0, 1 / 100.0, 1 / 10.0, 1 / 2.0]
z = np.load(sys.argv[1])
m = {k: z[k] for k in z.files}
params, sec = tl.set_tools_parameters(TOOLS)
_, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.arange(0, 100, 0.1), 5)
flat = planner.flatten_task(tasks[len(tasks) // 2], params, three_d=True)
ctx = _cabi.Context(0)
for rep in range(2):
    ctx.mesh_set(3, m["points"], m["elems"], m["mat"], m["bfacets"], m["bdir"], m["axis"])
    ndof, _ = ctx.space_build(2); nnz = ctx.nnz
    ctx.assemble(SIGMA)
    ctx.precond_setup("multigrid")
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
    iters, relres = ctx.solve(rtol=1e-10, maxit=20000)
    ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
    print("ndof %d iterations %s relres %.1e Ra %s stages %s" % (ndof, iters.tolist(), relres.max(), np.round(ra[:3], 3).tolist(),
          {k: round(v, 1) for k, v in ctx.stage_times().items()}), flush=True)
