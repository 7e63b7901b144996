"""Small cases for compute-sanitizer (tools/sanitize.sh; kept under tests/ because the oracle is the checker): every SpMM kind (element-wise product, SELL streaming / generic, CSR
one-column), both preconditioners, orders 1-3, 3D and 2D, 1 / 5 / 8 right-hand sides -- each checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import helpers  # noqa: E402
from oracle import fem_oracle as fo  # noqa: E402
from remo3d_b200 import _cabi  # noqa: E402


def run(name, mesh, sigma, flat, order, precond, opts=()):
    ctx = _cabi.Context(0)
    ctx.set_option("ebe_check", 1)
    for k, v in opts:
        ctx.set_option(k, v)
    ctx.mesh_set(mesh.dim, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), mesh.axis_vertices())
    ndof, _ = ctx.space_build(order)
    ctx.assemble(np.asarray(sigma, float))
    ctx.precond_setup(precond)
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
    it, rr = ctx.solve(rtol=1e-10, maxit=5000)
    ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
    ref = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary").astype(bool), order, flat,
                        dim=mesh.dim, solver="direct" if ndof < 30000 else "multigrid", rtol=1e-12)["ra"]
    err = float(np.max(np.abs(ra - ref) / np.abs(ref)))
    print("case %-34s ndof %6d kind %d iters %s max rel err %.2e" % (name, ndof, ctx.spmm_kind(), it.tolist(), err), flush=True)
    assert err < 1e-6, (name, err)
    ctx.close()


def main():
    mesh, sigma, flat, _ = helpers.ball_case(h_electrode=0.25, h_axis=0.8, grading=0.7)
    for precond in ("local", "multigrid"):
        run("ball order 2 %s (EBE)" % precond, mesh, sigma, flat, 2, precond)
        run("ball order 2 %s (SELL)" % precond, mesh, sigma, flat, 2, precond, (("spmm_ebe", 0),))
    run("ball order 1 multigrid", mesh, sigma, flat, 1, "multigrid")
    run("ball order 3 multigrid", mesh, sigma, flat, 3, "multigrid")
    mesh1, sigma1, flat1, _ = helpers.ball_case(h_electrode=0.25, h_axis=0.8, grading=0.7, tools=("A2.0M0.5N",), depths=(10.0,), batch_size=1)
    run("ball order 2 one column (EBE)", mesh1, sigma1, flat1, 2, "multigrid")
    run("ball order 2 one column (CSR)", mesh1, sigma1, flat1, 2, "multigrid", (("spmm_ebe", 0),))
    mesh8, sigma8, flat8, _ = helpers.ball_case(h_electrode=0.25, h_axis=0.8, grading=0.7, depths=tuple(10.0 + 0.1 * i for i in range(8)), batch_size=8)
    run("ball order 2 wide block", mesh8, sigma8, flat8, 2, "multigrid")
    m2, s2, f2, _ = helpers.disc_case(h_electrode=0.08, h_axis=0.3, h_borehole=0.3, grading=0.8)
    run("disc order 3 multigrid", m2, s2, f2, 3, "multigrid")
    run("disc order 2 local", m2, s2, f2, 2, "local")
    print("all cases ok")


if __name__ == "__main__":
    main()
