"""The CPU oracle pinned by independent checks: reference tensors by numerical quadrature and against the generated
CUDA header, polynomial exactness of the element matrices, the homogeneous-ball known answer, reciprocity."""
import os
import re
from fractions import Fraction

import numpy as np
import pytest

from oracle import fem_oracle as fo
from remo3d_b200 import meshgen, planner, tools as tl
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _eval_poly(p, lam):
    return sum(float(c) * np.prod([lam[i] ** e for i, e in enumerate(exps)], axis=0) for exps, c in p.items()) if p else 0.0 * lam[0]


def _duffy_rule(dim, n=8):
    """Gauss-Legendre on the Duffy-collapsed cube: exact for the polynomial degrees used here; weights sum to 1."""
    x, w = np.polynomial.legendre.leggauss(n)
    x, w = 0.5 * (x + 1), 0.5 * w
    if dim == 2:
        a, b = np.meshgrid(x, x, indexing="ij")
        wa, wb = np.meshgrid(w, w, indexing="ij")
        l1, l2 = a, b * (1 - a)
        wt = wa * wb * (1 - a) * 2.0
        lam = [1 - l1 - l2, l1, l2]
    else:
        a, b, c = np.meshgrid(x, x, x, indexing="ij")
        wa, wb, wc = np.meshgrid(w, w, w, indexing="ij")
        l1, l2, l3 = a, b * (1 - a), c * (1 - a) * (1 - b)
        wt = wa * wb * wc * (1 - a) ** 2 * (1 - b) * 6.0
        lam = [1 - l1 - l2 - l3, l1, l2, l3]
    return [l.ravel() for l in lam], wt.ravel()


@pytest.mark.parametrize("dim,order", [(3, 1), (3, 2), (3, 3), (2, 1), (2, 2), (2, 3)])
def test_reference_tensors_by_quadrature(dim, order):
    lam, wt = _duffy_rule(dim)
    assert abs(wt.sum() - 1.0) < 1e-13
    basis = fo.local_basis(dim, order)
    n = dim + 1
    dvals = [[_eval_poly(fo._pdiff(b, i), lam) for i in range(n)] for b in basis]
    pairs = fo.metric_pairs(dim)
    if dim == 3:
        T, _ = fo.reference_tensors(3, order)
        weights = [np.ones_like(wt)]
        T = T[None]
    else:
        T, _ = fo.reference_tensors(2, order, weighted=True)
        weights = lam
    for k, wk in enumerate(weights):
        for m, (i, j) in enumerate(pairs):
            for a in range(len(basis)):
                for b in range(len(basis)):
                    f = dvals[a][i] * dvals[b][j]
                    if i != j:
                        f = f + dvals[a][j] * dvals[b][i]
                    assert abs(np.sum(wt * wk * f) - T[k, m, a, b]) < 1e-12


def test_generated_cuda_header_matches_oracle_tensors():
    text = open(os.path.join(ROOT, "remo3d_b200", "csrc", "ref_tensors.inc")).read()
    for dim, tag in ((3, "T3"), (2, "T2")):
        for p in (1, 2, 3):
            body = re.search(r"REF_%s_P%d\[\d+\] = \{(.*?)\};" % (tag, p), text, flags=re.S).group(1)
            vals = [float(Fraction(int(float(a)), int(float(b)))) if b else float(a)
                    for a, b in re.findall(r"(-?\d+\.0)(?:/(\d+\.0))?", body)]
            ref = fo.reference_tensors(dim, p, weighted=(dim == 2))[0].ravel()
            np.testing.assert_allclose(np.array(vals), ref, rtol=0, atol=1e-15)


def _quadratic_interpolant(space, pts, u):
    """Coefficients of a quadratic function in the hierarchical basis (vertex values + edge bubbles)."""
    x = np.zeros(space.ndof)
    x[: space.nv] = u(pts)
    if space.order >= 2:
        a, b = space.edges[:, 0], space.edges[:, 1]
        mid = u(0.5 * (pts[a] + pts[b]))
        pe = space.order - 1
        x[space.edge_base + pe * np.arange(space.ne)] = 4 * mid - 2 * (x[a] + x[b])
    return x


def test_p2_tet_closed_form_matches_tensors():
    """The closed form remo3d_b200/csrc/ebe.cu (p2_apply) uses for y = K_e x on an order-2 tet -- written from
    grad u = sum_j c_j(l) grad l_j with c_j linear in the barycentrics -- against K_e = sum_m S_m T[m] with the exact
    reference tensors, for random metric numbers and vectors."""
    T = fo.reference_tensors(3, 2)[0]
    pairs = fo.metric_pairs(3)
    le = fo.local_edges(3)
    rng = np.random.default_rng(11)
    for _ in range(20):
        g = rng.standard_normal((4, 3))
        g[0] = -(g[1] + g[2] + g[3])
        S = g @ g.T
        K = np.einsum("m,mab->ab", np.array([S[i, j] for i, j in pairs]), T)
        x = rng.standard_normal(10)
        xe = np.zeros((4, 4))
        for e, (a, b) in enumerate(le):
            xe[a, b] = xe[b, a] = x[4 + e]
        sj = xe.sum(axis=1)
        d0 = x[:4] + 0.25 * sj          # mean of c_j
        bs = 0.25 * x[:4] + 0.05 * sj   # mean of l_a c_j without the x_{ja} term
        y = np.empty(10)
        y[:4] = S @ d0
        B = S @ bs
        V = S @ xe                      # V[b][a] = sum_j S_bj xe[j][a]
        for e, (a, b) in enumerate(le):
            y[4 + e] = B[a] + B[b] + 0.05 * (V[b, a] + V[a, b])
        assert np.max(np.abs(y - K @ x)) <= 1e-14 * np.max(np.abs(K)) * np.max(np.abs(x)) * 10


@pytest.mark.parametrize("order", [1, 2, 3])
def test_energy_of_polynomials_is_exact(order):
    pts, elems, bf, bc = meshgen.box_mesh(3)
    space = fo.Space(pts.shape[0], elems, order, 3)
    A = fo.assemble(pts, space, [2.0], np.zeros(elems.shape[0], int))
    assert abs(A - A.T).max() < 1e-12
    const = np.zeros(space.ndof)
    const[: space.nv] = 1.0
    assert np.abs(A @ const).max() < 1e-11
    lin = _quadratic_interpolant(space, pts, lambda p: 1 + 2 * p[:, 0] - 3 * p[:, 1] + 0.5 * p[:, 2]) if order > 1 else None
    x = np.zeros(space.ndof)
    x[: space.nv] = 1 + 2 * pts[:, 0] - 3 * pts[:, 1] + 0.5 * pts[:, 2]
    vol = 2.0
    assert abs(x @ A @ x - 2.0 * vol * (4 + 9 + 0.25)) < 1e-10
    if order >= 2:
        np.testing.assert_allclose(lin, x, atol=1e-12)
        q = _quadratic_interpolant(space, pts, lambda p: p[:, 0] ** 2 + p[:, 1] * p[:, 2])
        # int over [0,1]^2 x [-1,1] of sigma (4x^2 + z^2 + y^2) = 2 * (8/3 + 2/3 + 2/3)
        assert abs(q @ A @ q - 2.0 * (8 / 3 + 2 / 3 + 2 / 3)) < 1e-10


def test_homogeneous_half_ball_known_answer():
    """u = rho/(4 pi r) scaled by 2 on the half-ball -> Ra == rho for every tool (worker.py:129-131), up to discretisation."""
    mesh, sigma, flat, _ = helpers.ball_case(h_electrode=0.05, h_axis=0.2, grading=0.45, layered=False)
    res = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), 2, flat)
    np.testing.assert_allclose(res["ra"], 10.0, rtol=1e-2)
    res1 = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), 1, flat)
    np.testing.assert_allclose(res1["ra"], 10.0, rtol=6e-2)
    assert np.abs(res["ra"] - 10).max() < np.abs(res1["ra"] - 10).max()


def test_reciprocity_of_two_electrode_tool():
    """remo3d.py:211-214 swaps A,B,M -> M,N,A (one current electrode instead of two).  On one mesh the discrete
    reciprocity u_{A,-B}(M) == u_M(A) - u_M(B) holds to solver accuracy (symmetric stiffness matrix)."""
    mesh, sigma, flat, _ = helpers.ball_case(h_electrode=0.1, h_axis=0.4, grading=0.6)
    space = fo.Space(mesh.nv, mesh.elems, 2, 3)
    A = fo.assemble(mesh.points, space, sigma, mesh.mat)
    con = space.dirichlet_dofs(mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"))
    axis = fo.Axis(mesh.points, space)
    zs = np.unique(np.concatenate([flat["src_z"], flat["pt_z0"]]))
    za, zb, zm = zs[0], zs[2], zs[-1]
    f1 = fo.point_source_rhs(axis, space.ndof, [za, zb], [1.0, -1.0])
    f2 = fo.point_source_rhs(axis, space.ndof, [zm], [1.0])
    U = fo.solve_direct(A, np.stack([f1, f2], axis=1), con)
    lhs = fo.sample_axis(axis, U[:, 0], zm)
    rhs = fo.sample_axis(axis, U[:, 1], za) - fo.sample_axis(axis, U[:, 1], zb)
    assert abs(lhs - rhs) <= 1e-9 * abs(lhs)


def test_jacobi_pcg_matches_direct():
    mesh, sigma, flat, _ = helpers.ball_case(h_electrode=0.12, h_axis=0.5, grading=0.7)
    a = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), 1, flat, solver="direct")
    b = fo.solve_task(mesh.points, mesh.elems, mesh.mat, sigma, mesh.bfacets, mesh.dirichlet_flags("dirichlet_boundary"), 1, flat, solver="pcg")
    np.testing.assert_allclose(a["ra"], b["ra"], rtol=1e-9)


def test_p3_tet_pair_list_matches_tensors():
    """csrc/ebe_p3_apply.inc (order-3 element-wise product) is generated from the symmetric pair list of the exact reference
    tensors: its NumPy emulation equals K x, and the committed file is what the generator emits."""
    import importlib.util
    import os
    import tempfile

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ebe_apply", os.path.join(root, "tools", "gen_ebe_apply.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert gen.check(T=fo.reference_tensors(3, 3)[0]) < 1e-14  # against the oracle's independent derivation of the tensors
    Tw = fo.reference_tensors(2, 3, weighted=True)[0]              # [k, m, a, b] -> the generator's m = 6 k + pair
    assert gen.check(T=Tw.reshape(18, Tw.shape[2], Tw.shape[3]), dim=2) < 1e-14
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "p3.inc")
        gen.emit(path)
        gen.emit(path, dim=2, name="p3tri_apply", mode="a")
        assert open(path).read() == open(os.path.join(root, "remo3d_b200", "csrc", "ebe_p3_apply.inc")).read()
