"""GPU probe: PCG iterations ("multigrid") on the bench mesh of one size class with and without the sliver pass of
meshgen.half_ball_mesh(improve=N).  ctypes only (no torch import): python tools/mesh_quality_probe.py 200k 3"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remo3d_b200 import _cabi, meshgen, planner, tools as tl
from remo3d_b200.mesh import Mesh

TOOLS = ["A2.0M0.5N", "N0.5M2.0A", "B5.7A0.4M", "M4.0A0.5B"]
SIGMA = [1 / 1.0, 1 / 10.0, 1 / 100.0, 1 / 10.0, 1 / 2.0]
SIZES = {"1M": (0.01, 0.04, 0.19, 5.0), "200k": (0.03, 0.1, 0.33, 6.0), "60k": (0.06, 0.25, 0.5, 6.0)}
size, rounds = sys.argv[1], [int(x) for x in sys.argv[2].split(",")]
params, sec = tl.set_tools_parameters(TOOLS)
_, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.arange(0, 100, 0.1), 5)
task = tasks[len(tasks) // 2]
flat = planner.flatten_task(task, params, three_d=True)
he, ha, g, hm = SIZES[size]
material = meshgen.layered_material([-1.0, 1.5], dip_rad=np.deg2rad(30.0), borehole_radius=0.1, inclusion=((3.0, 2.0, 1.0), 1.5))
ctx = None
for imp in rounds:
    t0 = time.time()
    m = meshgen.half_ball_mesh(50.0, task[1][0], material=material, h_electrode=he, h_axis=ha, grading=g, h_max=hm, seed=0, improve=imp)
    mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
    q = meshgen._quality(m["points"], m["elems"])
    t1 = time.time()
    if ctx is None:
        ctx = _cabi.Context(0)
    ctx.mesh_set(3, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), mesh.axis_vertices())
    ndof, _ = ctx.space_build(2); nnz = ctx.nnz
    ctx.assemble(SIGMA)
    ctx.precond_setup("multigrid")
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
    iters, relres = ctx.solve(rtol=1e-10, maxit=20000)
    ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
    print("improve=%d: mesh %.1fs ndof %d  tets<0.1: %d  min quality %.3g  iterations %s  relres %.1e  Ra %s  solve-stage ms %.1f"
          % (imp, t1 - t0, ndof, int((q < 0.1).sum()), q.min(), iters.tolist(), relres.max(), np.round(ra[:4], 4).tolist(), ctx.stage_times()["solve"]), flush=True)
