"""Config C3 of BASELINE.json: `Examples/Benchmark models/Benchmark model 2` (10 / 100 ohm-m layers with 0.2 / 0.35 / 0.5 m
invasion of 5 ohm-m, 200 mm borehole, Rm = 1), the full normal / lateral tool set, all sources that share a mesh solved as
ONE multi-right-hand-side block: nrhs = batch_size, swept over 1 / 5 / 10 / 20 / 32.  GPU path (order 3, 2D
axisymmetric, both preconditioners) against the oracle's direct solve on the same mesh: Ra <= 1e-6 relative."""
import os

import numpy as np
import pytest

from oracle import fem_oracle as fo
from remo3d_b200 import _cabi, model_io, model_mesh, planner, tools as tl
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch_size", [1, 5, 10, 20, 32])
def test_benchmark_model_2_block_sizes(golden_dir, batch_size):
    d = os.path.join(golden_dir, "bm2")
    formation = model_io.load_formation_parameters(os.path.join(d, "Formation_BM2.txt"))
    borehole = model_io.load_borehole_parameters(os.path.join(d, "Borehole_BM2.txt"))
    params, sec = tl.set_tools_parameters(helpers.SIX_TOOLS)
    depths = np.arange(12.0, 18.0, 0.1)  # across the 15 m interface between an invaded 100 ohm-m bed and a 10 ohm-m bed
    centres, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, depths, batch_size)
    task = tasks[len(tasks) // 2]
    mud = float(np.interp(centres[task[0]], borehole[:, 0], borehole[:, 2]))
    mesh, sigma = model_mesh.build_task_mesh(formation, borehole[:, :2], 0.0, centres[task[0]], task[1][0], mud, 50.0,
                                             {"h_electrode": 0.02, "h_axis": 0.1, "h_borehole": 0.15, "grading": 0.5})
    flat = planner.flatten_task(task, params, three_d=False)
    nrhs = flat["src_ptr"].shape[0] - 1
    assert nrhs == batch_size
    flags = mesh.dirichlet_flags("dirichlet_boundary")
    ref = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, flags, 3, flat, dim=2, solver="direct")
    ctx = _cabi.Context(0)
    try:
        ctx.mesh_set(2, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, flags, mesh.axis_vertices())
        ctx.space_build(3)
        ctx.assemble(sigma)
        for pre in ("multigrid", "local"):
            ctx.precond_setup(pre)
            ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
            iters, relres = ctx.solve(rtol=1e-10, maxit=50000)
            assert (relres <= 1e-10).all()
            ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
            np.testing.assert_allclose(ra, ref["ra"], rtol=1e-6)
            print("batch", batch_size, pre, "ndof", ctx.ndof, "log points", ra.shape[0], "iterations", int(iters.max()), "spmm kind", ctx.spmm_kind())
    finally:
        ctx.close()
