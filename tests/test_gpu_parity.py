"""GPU parity: the CUDA path through the C ABI against the CPU oracle on identical meshes.

Bars (BASELINE.json north_star): mesh/DOF numbering and CSR pattern bit-exact; assembled matrix <= 1e-12 relative
Frobenius; electrode potentials and apparent resistivity <= 1e-6 relative at CG relative residual 1e-10."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fem_oracle as fo
from remo3d_b200 import _cabi
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _cabi.Context(0)
    yield c
    c.close()


def _setup(ctx, mesh, order, dirichlet="dirichlet_boundary"):
    flags = mesh.dirichlet_flags(dirichlet)
    ctx.mesh_set(mesh.dim, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, flags, mesh.axis_vertices())
    ctx.space_build(order)
    return flags


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("case", ["box", "ball"])
def test_numbering_pattern_matrix(ctx, order, case):
    if case == "box":
        mesh, sigma = helpers.box_case(3)
    else:
        mesh, sigma, _, _ = helpers.ball_case()
    flags = _setup(ctx, mesh, order)
    space = fo.Space(mesh.nv, mesh.elems, order, 3)
    assert ctx.ndof == space.ndof and ctx.ne == space.ne
    edges, faces, ee, ef = ctx.topology()
    np.testing.assert_array_equal(edges, space.edges)
    np.testing.assert_array_equal(ee, space.elem_edges)
    if order == 3:
        assert ctx.nf == space.nf
        np.testing.assert_array_equal(faces, space.faces)
        np.testing.assert_array_equal(ef, space.elem_faces)
    np.testing.assert_array_equal(ctx.dirichlet(), space.dirichlet_dofs(mesh.bfacets, flags))

    ctx.assemble(sigma)
    rowptr, col, val = ctx.matrix()
    A = fo.assemble(mesh.points, space, sigma, mesh.mat)
    assert ctx.nnz == A.nnz
    np.testing.assert_array_equal(rowptr, A.indptr)
    np.testing.assert_array_equal(col, A.indices)
    err = np.linalg.norm(val - A.data) / np.linalg.norm(A.data)
    assert err <= 1e-12, err
    # structural properties of a stiffness matrix: symmetric, zero row sums on the vertex block
    G = sp.csr_matrix((val, col, rowptr), shape=(ctx.ndof, ctx.ndof))
    assert abs(G - G.T).max() <= 1e-12 * abs(G).max()
    ones = np.zeros(ctx.ndof)
    ones[: mesh.nv] = 1.0  # the constant function in the hierarchical basis
    assert np.abs(G @ ones).max() <= 1e-10 * abs(G).max()


def test_assembly_is_bit_reproducible(ctx):
    mesh, sigma, _, _ = helpers.ball_case()
    _setup(ctx, mesh, 2)
    ctx.assemble(sigma)
    v1 = ctx.matrix()[2].copy()
    ctx.assemble(sigma)
    v2 = ctx.matrix()[2]
    assert np.array_equal(v1, v2)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_solve_matches_oracle(ctx, order):
    mesh, sigma, flat, _ = helpers.ball_case()
    ref = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), order, flat)
    _setup(ctx, mesh, order)
    ctx.assemble(sigma)
    ctx.precond_setup("local")
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
    nrhs = flat["src_ptr"].shape[0] - 1
    axis = ref["axis"]
    for r in range(nrhs):
        lo, hi = flat["src_ptr"][r], flat["src_ptr"][r + 1]
        f = fo.point_source_rhs(axis, ctx.ndof, flat["src_z"][lo:hi], flat["src_fac"][lo:hi])
        np.testing.assert_allclose(ctx.rhs(r), f, rtol=0, atol=1e-15)
    iters, relres = ctx.solve(rtol=1e-10, maxit=20000)
    assert (relres <= 1e-10).all() and (iters > 0).all()
    # potentials at every electrode of the task
    zs = np.unique(np.concatenate([flat["pt_z0"], flat["pt_z1"][~np.isnan(flat["pt_z1"])]]))
    for r in range(nrhs):
        u_gpu = ctx.sample_axis(zs, np.full(zs.shape, r))
        u_ref = np.array([fo.sample_axis(axis, ref["U"][:, r], z) for z in zs])
        np.testing.assert_allclose(u_gpu, u_ref, rtol=1e-6)
        # the whole solution vector, relative to its norm
        u = ctx.solution(r)
        assert np.linalg.norm(u - ref["U"][:, r]) <= 1e-6 * np.linalg.norm(ref["U"][:, r])
    ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
    np.testing.assert_allclose(ra, ref["ra"], rtol=1e-6)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_multigrid_preconditioner_same_solution(ctx, order):
    """'multigrid' (hierarchical two-level + aggregation AMG on the P1 block) must reach the same solution as the
    oracle, in far fewer iterations than 'local' (Jacobi)."""
    mesh, sigma, flat, _ = helpers.ball_case()
    ref = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), order, flat)
    _setup(ctx, mesh, order)
    ctx.assemble(sigma)
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
    ctx.precond_setup("local")
    it_jac, _ = ctx.solve(rtol=1e-10, maxit=20000)
    ctx.precond_setup("multigrid")
    it_mg, relres = ctx.solve(rtol=1e-10, maxit=2000)
    assert (relres <= 1e-10).all()
    assert it_mg.max() < it_jac.max(), (it_mg, it_jac)
    ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
    np.testing.assert_allclose(ra, ref["ra"], rtol=1e-6)
    for r in range(ctx.nrhs):
        u = ctx.solution(r)
        assert np.linalg.norm(u - ref["U"][:, r]) <= 1e-6 * np.linalg.norm(ref["U"][:, r])
    print("order", order, "iterations jacobi", it_jac.tolist(), "multigrid", it_mg.tolist())


@pytest.mark.parametrize("dim,order", [(3, 1), (3, 2), (3, 3), (2, 2), (2, 3)])
def test_preconditioner_pieces_come_from_the_elements(ctx, dim, order):
    """remo_precond_setup never reads the assembled matrix: diag(A) and the vertex block of the V-cycle are gathered
    from the element metrics.  Both must equal the corresponding entries of the oracle's matrix (the vertex block
    bit-identically to the library's own CSR export, which is built afterwards, on demand)."""
    if dim == 3:
        mesh, sigma = helpers.ball_case()[:2]
        flags = mesh.dirichlet_flags("dirichlet_boundary")
    else:
        mesh, sigma = helpers.disc_case()[:2]
        flags = mesh.dirichlet_flags([2])
    ctx.mesh_set(dim, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, flags, mesh.axis_vertices())
    ndof, nnz0 = ctx.space_build(order)
    assert nnz0 == 0  # lazy: no CSR pattern yet
    ctx.assemble(sigma)
    ctx.precond_setup("multigrid")
    dinv, (rp, cl, vl), levels = ctx.precond_get(vertex_block=True)
    space = fo.Space(mesh.nv, mesh.elems, order, dim)
    A = fo.assemble(mesh.points, space, sigma, mesh.mat)
    con = space.dirichlet_dofs(mesh.bfacets, flags)
    want = np.where(con, 0.0, 1.0 / A.diagonal())
    np.testing.assert_allclose(dinv, want, rtol=1e-11, atol=0)  # vs the oracle; bit-identical to the library's own CSR below
    Avv = A[: mesh.nv][:, : mesh.nv].tocsr()
    Avv.sort_indices()
    np.testing.assert_array_equal(rp, Avv.indptr)
    np.testing.assert_array_equal(cl, Avv.indices)
    assert np.linalg.norm(vl - Avv.data) <= 1e-12 * np.linalg.norm(Avv.data)
    # ... and bit-identical to the same entries of the library's own matrix
    rowptr, col, val = ctx.matrix()
    G = sp.csr_matrix((val, col, rowptr), shape=(ndof, ndof))[: mesh.nv][:, : mesh.nv].tocsr()
    G.sort_indices()
    assert np.array_equal(G.data, vl)
    assert np.array_equal(np.where(con, 0.0, 1.0 / sp.csr_matrix((val, col, rowptr), shape=(ndof, ndof)).diagonal()), dinv)
    assert levels[0] == (mesh.nv, Avv.nnz) and all(a[0] > b[0] for a, b in zip(levels, levels[1:])) and levels[-1][0] <= 512
    print("levels", levels)


def test_eager_matrix_option_gives_the_same_matrix(ctx):
    mesh, sigma = helpers.ball_case()[:2]
    _setup(ctx, mesh, 2)
    ctx.assemble(sigma)
    lazy = ctx.matrix()
    ctx.set_option("lazy_matrix", 0)
    try:
        ctx.mesh_set(3, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), mesh.axis_vertices())
        ndof, nnz = ctx.space_build(2)
        assert nnz == lazy[1].shape[0]
        ctx.assemble(sigma)
        eager = ctx.matrix()
    finally:
        ctx.set_option("lazy_matrix", 1)
    for a, b in zip(lazy, eager):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("opts", [{"amg_agg": 0}, {"amg_agg": 1, "amg_passes": 3}, {"amg_agg": 1, "amg_passes": 2}, {"amg_agg": 1, "amg_passes": 2, "amg_rounds": 2}])
def test_aggregation_variants_same_solution(ctx, opts):
    """The V-cycle hierarchy (Morton-rank aggregates of round 1, strength-based pairwise aggregation with 1-3 passes) only
    changes the preconditioner: same solution, and the pairwise aggregates never need more iterations than 1.15 x Morton's."""
    mesh, sigma, flat, _ = helpers.ball_case(h_electrode=0.05, h_axis=0.2, grading=0.45)
    ref = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), 2, flat)
    _setup(ctx, mesh, 2)
    ctx.assemble(sigma)
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
    its = {}
    try:
        for name, o in (("morton", {"amg_agg": 0}), ("this", opts)):
            for k, v in o.items():
                ctx.set_option(k, v)
            ctx.precond_setup("multigrid")
            it, relres = ctx.solve(rtol=1e-10, maxit=3000)
            assert (relres <= 1e-10).all()
            its[name] = int(it.max())
            ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
            np.testing.assert_allclose(ra, ref["ra"], rtol=1e-6)
            print(name, o, "iterations", it.tolist(), "levels", ctx.precond_get()[2])
    finally:
        for k, v in (("amg_agg", 1), ("amg_passes", 3), ("amg_rounds", 4)):
            ctx.set_option(k, v)
    assert its["this"] <= 1.15 * its["morton"] + 2, its


@pytest.mark.parametrize("fp32", [0, 1])
def test_fused_tail_of_the_vcycle_is_bit_identical(ctx, fp32):
    """The small levels of the V-cycle can run as one cluster kernel (amg.cu k_vcycle_tail<T>, both precisions of the cycle);
    with the per-level launches of the same lane layout (4 lanes per row) the PCG must take the same iterations and give the
    same bits (SELL product: deterministic summation order)."""
    mesh, sigma, flat, _ = helpers.ball_case(h_electrode=0.05, h_axis=0.2, grading=0.45)
    _setup(ctx, mesh, 2)
    ctx.assemble(sigma)
    out = {}
    try:
        ctx.set_option("spmm_ebe", 0)
        ctx.set_option("amg_fp32", fp32)
        ctx.set_option("amg_lanes8", 0)
        for fused in (1, 0):
            ctx.set_option("amg_fused_tail", fused)
            ctx.set_option("amg_tail_rows", 4000)
            ctx.precond_setup("multigrid")
            ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
            it, relres = ctx.solve(rtol=1e-10, maxit=3000)
            assert (relres <= 1e-10).all()
            out[fused] = (it.copy(), np.stack([ctx.solution(r) for r in range(ctx.nrhs)]))
            print("fp32", fp32, "fused", fused, "iterations", it.tolist(), "levels", ctx.precond_get()[2])
    finally:
        ctx.set_option("spmm_ebe", 1)
        ctx.set_option("amg_fused_tail", 0)
        ctx.set_option("amg_tail_rows", 20000)
        ctx.set_option("amg_fp32", 1)
        ctx.set_option("amg_lanes8", 1)
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])


def test_mixed_precision_vcycle_against_the_fp64_cycle(ctx):
    """`amg_fp32` (default): the V-cycle in fp32 inside the fp64 PCG.  Same solution to 1e-8, iteration counts within 5 % of the
    fp64 cycle (measured: 235-248 on this mesh; at C4 158 vs 160), both lane layouts of the sweeps."""
    mesh, sigma, flat, _ = helpers.ball_case(h_electrode=0.05, h_axis=0.2, grading=0.45)
    _setup(ctx, mesh, 2)
    ctx.assemble(sigma)
    res = {}
    try:
        for fp32, lanes8 in ((0, 0), (1, 0), (1, 1), (0, 1)):
            ctx.set_option("amg_fp32", fp32)
            ctx.set_option("amg_lanes8", lanes8)
            ctx.precond_setup("multigrid")
            ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
            it, relres = ctx.solve(rtol=1e-10, maxit=3000)
            assert (relres <= 1e-10).all()
            res[(fp32, lanes8)] = (it.copy(), np.stack([ctx.solution(r) for r in range(ctx.nrhs)]))
    finally:
        ctx.set_option("amg_fp32", 1)
        ctx.set_option("amg_lanes8", 1)
    it0, u0 = res[(0, 0)]
    for key, (it, u) in res.items():
        assert np.abs(it.astype(int) - it0.astype(int)).max() <= 0.05 * it0.max(), (key, it, it0)
        assert np.abs(u - u0).max() <= 1e-8 * np.abs(u0).max(), key


def test_mesh_after_the_sliver_pass_matches_oracle(ctx):
    """bench.py's meshes go through meshgen.half_ball_mesh(improve=N) (sliver pass): the path on such a mesh against the
    oracle -- numbering, matrix, Ra (both preconditioners)."""
    from remo3d_b200 import meshgen, planner, tools as tl
    from remo3d_b200.mesh import Mesh

    params, sec = tl.set_tools_parameters(["A2.0M0.5N", "N0.5M2.0A", "B5.7A0.4M", "M4.0A0.5B"])
    _, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.arange(0, 100, 0.1), 5)
    task = tasks[len(tasks) // 2]
    flat = planner.flatten_task(task, params, three_d=True)
    material = meshgen.layered_material([-1.0, 1.5], dip_rad=np.deg2rad(30.0), borehole_radius=0.1, inclusion=((3.0, 2.0, 1.0), 1.5))
    m = meshgen.half_ball_mesh(50.0, task[1][0], material=material, h_electrode=0.08, h_axis=0.3, grading=0.55, h_max=6.0, seed=0, improve=3)
    mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
    sigma = [1 / 1.0, 1 / 10.0, 1 / 100.0, 1 / 10.0, 1 / 2.0]
    flags = mesh.dirichlet_flags("dirichlet_boundary")
    ref = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, flags, 2, flat)
    _setup(ctx, mesh, 2)
    ctx.assemble(sigma)
    rowptr, col, val = ctx.matrix()
    np.testing.assert_array_equal(rowptr, ref["A"].indptr)
    np.testing.assert_array_equal(col, ref["A"].indices)
    assert np.linalg.norm(val - ref["A"].data) <= 1e-12 * np.linalg.norm(ref["A"].data)
    for pre in ("multigrid", "local"):
        ctx.precond_setup(pre)
        ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
        it, relres = ctx.solve(rtol=1e-10, maxit=20000)
        assert (relres <= 1e-10).all()
        ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
        np.testing.assert_allclose(ra, ref["ra"], rtol=1e-6)


def test_homogeneous_ball_gives_rho(ctx):
    """Known answer: homogeneous medium -> Ra == rho for every tool, up to discretisation error."""
    mesh, sigma, flat, _ = helpers.ball_case(h_electrode=0.03, h_axis=0.12, grading=0.4, layered=False)
    _setup(ctx, mesh, 2)
    ctx.assemble(sigma)
    ctx.precond_setup("local")
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
    ctx.solve(rtol=1e-10, maxit=20000)
    ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
    np.testing.assert_allclose(ra, 10.0, rtol=5e-3)


def test_two_current_electrodes_and_edge_sources(ctx):
    """+1/-1 source pairs (force_single_electrode_configuration=False) and electrodes that fall inside axis edges."""
    mesh, sigma, flat, _ = helpers.ball_case(tools=("A0.2B3.0M", "N1.0A0.5B"), fsec=False, depths=(10.0, 10.1), batch_size=4)
    # shift every electrode by 1 mm: none is a mesh vertex any more -> edge shape functions are exercised
    for key in ("src_z", "pt_z0", "pt_z1"):
        flat[key] = flat[key] + 1e-3
    for order in (2, 3):
        ref = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), order, flat)
        _setup(ctx, mesh, order)
        ctx.assemble(sigma)
        ctx.precond_setup("local")
        ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
        ctx.solve(rtol=1e-10, maxit=20000)
        ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
        np.testing.assert_allclose(ra, ref["ra"], rtol=1e-6)
        assert np.isnan(flat["pt_z1"]).all()  # single potential electrode branch of worker.py:120-131


def test_error_paths(ctx):
    mesh, sigma = helpers.box_case(2)
    with pytest.raises(_cabi.RemoError):
        _cabi.Context(99)
    _setup(ctx, mesh, 1)
    with pytest.raises(_cabi.RemoError) as e:
        ctx.assemble([1.0])  # material index 2 outside the sigma list
    assert e.value.code == _cabi.ERR_ARG
    ctx.assemble(sigma)
    with pytest.raises(_cabi.RemoError) as e:
        ctx.solve()
    assert e.value.code == _cabi.ERR_STATE
    ctx.precond_setup("local")
    with pytest.raises(_cabi.RemoError) as e:
        ctx.rhs_point_sources([0, 1], [5.0], [1.0])  # outside the axis
    assert e.value.code == _cabi.ERR_MESH
    with pytest.raises(_cabi.RemoError):
        ctx.space_build(4)
    ctx.rhs_point_sources([0, 1], [0.0], [1.0])
    ctx.precond_setup("local") if False else None
    with pytest.raises(_cabi.RemoError) as e:
        ctx.solve(rtol=1e-14, maxit=2)
    assert e.value.code == _cabi.ERR_NOCONV


# ---------------------------------------------------------------------------------------------- 2D axisymmetric path
@pytest.mark.parametrize("order", [1, 2, 3])
def test_2d_axisymmetric_matches_oracle(ctx, order):
    """ngsolve_functions.py:34: 2 pi r sigma grad u . grad v on triangles (cell bubble at order 3), Dirichlet = bc [2]."""
    mesh, sigma, flat, _ = helpers.disc_case()
    flags = mesh.dirichlet_flags([2])
    ctx.mesh_set(2, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, flags, mesh.axis_vertices())
    ctx.space_build(order)
    space = fo.Space(mesh.nv, mesh.elems, order, 2)
    assert ctx.ndof == space.ndof and ctx.ne == space.ne
    edges, _, ee, _ = ctx.topology()
    np.testing.assert_array_equal(edges, space.edges)
    np.testing.assert_array_equal(ee, space.elem_edges)
    np.testing.assert_array_equal(ctx.dirichlet(), space.dirichlet_dofs(mesh.bfacets, flags))
    ctx.assemble(sigma)
    rowptr, col, val = ctx.matrix()
    A = fo.assemble(mesh.points, space, sigma, mesh.mat)
    np.testing.assert_array_equal(rowptr, A.indptr)
    np.testing.assert_array_equal(col, A.indices)
    assert np.linalg.norm(val - A.data) <= 1e-12 * np.linalg.norm(A.data)
    ref = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, flags, order, flat, dim=2, solver="direct")
    for pre in ("local", "multigrid"):
        ctx.precond_setup(pre)
        ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
        iters, relres = ctx.solve(rtol=1e-10, maxit=20000)
        assert (relres <= 1e-10).all()
        ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
        np.testing.assert_allclose(ra, ref["ra"], rtol=1e-6)
        for r in range(ctx.nrhs):
            u = ctx.solution(r)
            assert np.linalg.norm(u - ref["U"][:, r]) <= 1e-6 * np.linalg.norm(ref["U"][:, r])


# ---------------------------------------------------------------------------------------------- block sizes (config C3)
@pytest.mark.parametrize("nrhs", [1, 2, 3, 4, 7, 12, 17, 32])
def test_block_sizes_match_oracle(ctx, nrhs):
    """Every multi-right-hand-side width (all SpMM / vector-kernel template instances, odd widths padded to an even
    stride) against the oracle's direct solve; sources alternate single electrodes and +1/-1 pairs."""
    mesh, sigma, _, _ = helpers.ball_case(h_electrode=0.1, h_axis=0.4, grading=0.6)
    flags = mesh.dirichlet_flags("dirichlet_boundary")
    space = fo.Space(mesh.nv, mesh.elems, 2, 3)
    A = fo.assemble(mesh.points, space, sigma, mesh.mat)
    con = space.dirichlet_dofs(mesh.bfacets, flags)
    axis = fo.Axis(mesh.points, space)
    zs = axis.z[(axis.z > -8) & (axis.z < 8)]
    rng = np.random.default_rng(nrhs)
    ptr, sz, sf = [0], [], []
    for r in range(nrhs):
        if r % 3 == 2:
            a, b = rng.choice(zs, 2, replace=False)
            sz += [a + 1e-3, b]  # one of the pair inside an axis edge
            sf += [1.0, -1.0]
        else:
            sz += [float(rng.choice(zs))]
            sf += [1.0]
        ptr.append(len(sz))
    F = np.stack([fo.point_source_rhs(axis, space.ndof, sz[ptr[r]:ptr[r + 1]], sf[ptr[r]:ptr[r + 1]]) for r in range(nrhs)], axis=1)
    U = fo.solve_direct(A, F, con)
    ctx.mesh_set(3, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, flags, mesh.axis_vertices())
    ctx.space_build(2)
    ctx.assemble(sigma)
    ctx.precond_setup("multigrid" if nrhs % 2 else "local")
    ctx.rhs_point_sources(ptr, sz, sf)
    iters, relres = ctx.solve(rtol=1e-10, maxit=20000)
    assert iters.shape == (nrhs,) and (relres <= 1e-10).all()
    for r in range(nrhs):
        np.testing.assert_allclose(ctx.rhs(r), F[:, r], rtol=0, atol=1e-15)
        u = ctx.solution(r)
        assert np.linalg.norm(u - U[:, r]) <= 1e-6 * np.linalg.norm(U[:, r]), r
    pts = rng.choice(zs, 6)
    rhs = rng.integers(0, nrhs, 6)
    got = ctx.sample_axis(pts, rhs)
    want = np.array([fo.sample_axis(axis, U[:, r], z) for z, r in zip(pts, rhs)])
    np.testing.assert_allclose(got, want, rtol=1e-6)
    with pytest.raises(_cabi.RemoError):
        ctx.solution(nrhs)  # the padding column of an odd width is not addressable


def _check_spmm(ctx, ks, seed=5):
    rowptr, col, val = ctx.matrix()
    A = sp.csr_matrix((val, col, rowptr), shape=(ctx.ndof, ctx.ndof))
    con = ctx.dirichlet().astype(bool)
    absA = abs(A)
    rng = np.random.default_rng(seed)
    for k in ks:
        P = rng.standard_normal((ctx.ndof, k))
        Q, pq = ctx.spmm_apply(P)
        ref = A @ P
        ref[con] = 0.0
        scale = absA @ abs(P) + 1e-300
        assert np.max(abs(Q - ref) / scale) <= 1e-14 * 8, (k, np.max(abs(Q - ref) / scale))
        dots = (P * ref).sum(0)
        assert np.max(abs(pq - dots) / (abs(P * ref).sum(0))) <= 1e-13, k


@pytest.mark.parametrize("dim,order,ebe", [(3, 1, 1), (3, 2, 1), (3, 2, 0), (3, 3, 1), (3, 3, 0), (2, 2, 1), (2, 3, 1), (2, 3, 0)])
def test_spmm_kernels_against_scipy(ctx, dim, order, ebe):
    """Every SpMM kernel the PCG can pick (CSR one-column, SELL generic for strides 2/4/16/32, SELL streaming for 5..8
    right-hand sides, and the element-wise product of ebe.cu, which order-2 / order-3 tets and order-3 triangles take for up to
    6 right-hand sides unless switched off) against the exact product with the oracle-checked CSR matrix: Q = A P on free rows, 0 on constrained
    rows, and the fused per-column dots p.q."""
    mesh, sigma = (helpers.ball_case() if dim == 3 else helpers.disc_case())[:2]
    ctx.set_option("spmm_ebe", ebe)
    try:
        _setup(ctx, mesh, order)
        ctx.assemble(sigma)
        ctx.set_option("ebe_check", 1)
        _check_spmm(ctx, (1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 17, 32))
        if (dim == 3 and order in (2, 3)) or (dim == 2 and order == 3):  # tets of order 2 / 3, axisymmetric triangles of order 3
            ctx.spmm_apply(np.ones((ctx.ndof, 5)))
            assert ctx.spmm_kind() == (2 if ebe else 1), ctx.spmm_kind()  # 2 = element-wise product, 1 = SELL copy
    finally:
        ctx.set_option("ebe_check", 0)
        ctx.set_option("spmm_ebe", 1)


def test_elementwise_spmm_on_degenerate_batches(ctx):
    """ebe.cu on the star mesh: one vertex belongs to every tet, so inside a batch of 256 tets one dof has 256 entries
    (the longest jagged diagonal) next to edges with 2-3; and on a mesh smaller than one batch."""
    for order in (2, 3):
        mesh, sigma = helpers.star_case(nsphere=260 if order == 2 else 120)  # order 3: the row of the centre stays below the assembly limit
        _setup(ctx, mesh, order)
        ctx.assemble(sigma)
        _check_spmm(ctx, (1, 5, 6), seed=7)
        mesh, sigma = helpers.box_case(2)
        _setup(ctx, mesh, order)
        ctx.assemble(sigma)
        _check_spmm(ctx, (2, 5), seed=8)


@pytest.mark.parametrize("order", [2, 3])
def test_pattern_of_a_high_valence_vertex(ctx, order):
    """A vertex shared by ~500 tets: its row has > 2048 candidate columns, which takes the whole-CTA branch of the
    pattern builder; numbering, pattern and values must still match the oracle exactly."""
    mesh, sigma = helpers.star_case()
    deg = np.bincount(mesh.elems.ravel()).max()
    assert deg * (10 if order == 2 else 20) > 2048
    _setup(ctx, mesh, order)
    space = fo.Space(mesh.nv, mesh.elems, order, 3)
    A = fo.assemble(mesh.points, space, sigma, mesh.mat)
    rowptr, col, _ = ctx.matrix(values=False)
    np.testing.assert_array_equal(rowptr, A.indptr)
    np.testing.assert_array_equal(col, A.indices)
    if np.diff(A.indptr).max() > 2600:
        # documented limit of the numeric phase: the shared-memory row buffer of k_assemble_rows holds 8 rows at a time;
        # a 3600-entry row (order 3, 516 tets around one vertex) must fail loudly, not silently
        with pytest.raises(_cabi.RemoError, match="row buffer"):
            ctx.assemble(sigma)
            ctx.matrix()  # the CSR values are gathered on demand: the limit shows when somebody asks for them
        return
    ctx.assemble(sigma)
    val = ctx.matrix()[2]
    assert np.linalg.norm(val - A.data) / np.linalg.norm(A.data) <= 1e-12
