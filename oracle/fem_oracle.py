"""CPU ORACLE (test infrastructure, NOT product code) for the ReMo3D forward-solve hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product package `remo3d_b200` never does.

PARITY UNPINNED at the NGSolve boundary: the arithmetic of this path lives in Netgen/NGSolve
(un-vendored, un-pinned C++ dependency: `/root/reference/remo3d/ngsolve_functions.py:4`, `setup.py:12`
does not even list it), which is neither under /root/reference nor installable here, and the
reference has no tests or golden vectors at that boundary.  This file therefore
*restates the published algorithm* (H1-conforming hierarchical finite elements of order 1..3 on
simplices, Galerkin stiffness matrix, Dirichlet elimination, point-source right-hand side, point
evaluation) following the reference's own call sites, and is pinned by
  * known-answer tests (homogeneous ball: Ra == rho; patch tests; reciprocity) in tests/, and
  * the reference's committed full-pipeline logs (`Examples/Example_01/Output/.../Results_1.txt`)
    at the reference's own mesh-noise level (tests/test_oracle_golden.py).
Conventions that NGSolve fixes internally and that cannot be confirmed here (scaling of the
high-order basis functions, edge/face numbering) are DEFINED here; vertex values of the solution,
hence potentials at electrodes and apparent resistivities, are independent of them.

What is restated, with the reference line each function follows:
  topology / dof numbering      ngsolve_functions.py:27      fes = H1(mesh, order=3, dirichlet=...)
  element_matrices / assemble   ngsolve_functions.py:31-36,47 a += grad(u)*grad(v)*sigma*dx (3D),
                                                              2*pi*grad(u)*grad(v)*x*sigma*dx (2D)
  point_source_rhs              ngsolve_functions.py:10-21,39-44  AddPointSource
  solve                         ngsolve_functions.py:46-56   Preconditioner + CGSolver (+ condensation)
  sample_axis                   workers/worker.py:122-131    gfu(mesh(0, z)) / gfu(mesh(0, 0, z))
  apparent_resistivity          workers/worker.py:113-134

Discrete space.  Each simplex has its vertices sorted by global number, local
vertices 0<1<..<d.  With barycentric coordinates l_i the local basis is
    vertex i            : l_i
    edge (i<j), k=0     : l_i l_j                       (order >= 2)
    edge (i<j), k=1     : l_i l_j (l_j - l_i)           (order 3)
    face (i<j<k)        : l_i l_j l_k                   (order 3; in 2D this is the cell bubble)
local order  [vertices | edges 01,02,(03),12,(13),(23) x (p-1) | faces 012,(013,023,123)],
global order [vertices | edge e: nV+(p-1)e+k | faces: nV+(p-1)nE+f], edges / faces numbered
lexicographically by their sorted global vertex tuples.  Because vertices are sorted, every edge
function is oriented from the smaller to the larger global vertex on every element, so a single set
of reference tensors serves all elements.
"""
from fractions import Fraction
from functools import lru_cache
from itertools import combinations
from math import factorial

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SNAP_TOL = 1e-9  # |z - z_vertex| below which an axis point is treated as the vertex itself


# ----------------------------------------------------------------------------------------------
# reference element: exact polynomial algebra in barycentric coordinates
# ----------------------------------------------------------------------------------------------
def _pmul(p, q):
    out = {}
    for ea, ca in p.items():
        for eb, cb in q.items():
            e = tuple(x + y for x, y in zip(ea, eb))
            out[e] = out.get(e, 0) + ca * cb
    return {e: c for e, c in out.items() if c != 0}


def _padd(p, q, s=1):
    out = dict(p)
    for e, c in q.items():
        out[e] = out.get(e, 0) + s * c
    return {e: c for e, c in out.items() if c != 0}


def _pdiff(p, i):
    out = {}
    for e, c in p.items():
        if e[i] > 0:
            e2 = list(e)
            e2[i] -= 1
            out[tuple(e2)] = out.get(tuple(e2), 0) + c * e[i]
    return out


def _pint(p, dim):
    """(1/|K|) * integral over the simplex of a polynomial in barycentrics: d! prod(a_i!) / (|a|+d)!"""
    tot = Fraction(0)
    for e, c in p.items():
        num = factorial(dim)
        for a in e:
            num *= factorial(a)
        tot += Fraction(c) * Fraction(num, factorial(sum(e) + dim))
    return tot


def local_edges(dim):
    return list(combinations(range(dim + 1), 2))


def local_faces(dim):
    return list(combinations(range(dim + 1), 3))


def n_local_dofs(dim, order):
    nv = dim + 1
    ne = len(local_edges(dim)) * (order - 1)
    nf = len(local_faces(dim)) if order == 3 else 0
    return nv + ne + nf


def local_basis(dim, order):
    """List of basis polynomials (dict exponent-tuple -> Fraction) in the local dof order."""
    n = dim + 1

    def lam(i):
        return {tuple(1 if k == i else 0 for k in range(n)): Fraction(1)}

    basis = [lam(i) for i in range(n)]
    if order >= 2:
        for (i, j) in local_edges(dim):
            b0 = _pmul(lam(i), lam(j))
            basis.append(b0)
            if order == 3:
                basis.append(_pmul(b0, _padd(lam(j), lam(i), -1)))
    if order == 3:
        for (i, j, k) in local_faces(dim):
            basis.append(_pmul(_pmul(lam(i), lam(j)), lam(k)))
    assert len(basis) == n_local_dofs(dim, order)
    return basis


def metric_pairs(dim):
    """The (i<=j) index pairs of G_ij = grad l_i . grad l_j, in the order used by the tensors."""
    n = dim + 1
    return [(i, j) for i in range(n) for j in range(i, n)]


@lru_cache(maxsize=None)
def reference_tensors(dim, order, weighted=False):
    """Exact reference tensors.

    unweighted:  T[m, a, b]    with K_ab = sigma |K| sum_m G_m T[m,a,b]
    weighted  :  T[k, m, a, b] with K_ab = 2 pi sigma |K| sum_k r_k sum_m G_m T[k,m,a,b]   (2D axisymmetric)
    where m runs over metric_pairs(dim); for i<j the (i,j) and (j,i) terms are already summed.
    Returned as float64 (the rationals are exactly representable to 1 ulp) plus the Fractions.
    """
    n = dim + 1
    basis = local_basis(dim, order)
    nd = len(basis)
    grads = [[_pdiff(b, i) for i in range(n)] for b in basis]
    pairs = metric_pairs(dim)
    weights = [None] if not weighted else [{tuple(1 if q == k else 0 for q in range(n)): Fraction(1)} for k in range(n)]
    frac = np.empty((len(weights), len(pairs), nd, nd), dtype=object)
    for w, wt in enumerate(weights):
        for m, (i, j) in enumerate(pairs):
            for a in range(nd):
                for b in range(nd):
                    integrand = _pmul(grads[a][i], grads[b][j])
                    if i != j:
                        integrand = _padd(integrand, _pmul(grads[a][j], grads[b][i]))
                    if wt is not None:
                        integrand = _pmul(integrand, wt)
                    frac[w, m, a, b] = _pint(integrand, dim)
    val = np.array([[[[float(x) for x in row] for row in mat] for mat in blk] for blk in frac], dtype=np.float64)
    if not weighted:
        return val[0], frac[0]
    return val, frac


# ----------------------------------------------------------------------------------------------
# topology and dof numbering  (ngsolve_functions.py:27)
# ----------------------------------------------------------------------------------------------
def _unique_rows(keys):
    """Lexicographically sorted unique rows + inverse map."""
    uniq, inv = np.unique(keys, axis=0, return_inverse=True)
    return uniq, inv.reshape(-1)


class Space:
    """Topology + dof tables of the order-p H1 space on a simplicial mesh."""

    def __init__(self, nv, elems, order, dim):
        assert order in (1, 2, 3) and dim in (2, 3)
        self.nv, self.order, self.dim = int(nv), order, dim
        self.sorted_elems = np.sort(np.asarray(elems, dtype=np.int64), axis=1)
        se = self.sorted_elems
        nt = se.shape[0]
        le, lf = local_edges(dim), local_faces(dim)
        ekeys = np.stack([se[:, [i, j]] for (i, j) in le], axis=1).reshape(-1, 2)
        self.edges, inv = _unique_rows(ekeys)
        self.elem_edges = inv.reshape(nt, len(le))
        if dim == 3:
            fkeys = np.stack([se[:, [i, j, k]] for (i, j, k) in lf], axis=1).reshape(-1, 3)
            self.faces, inv = _unique_rows(fkeys)
            self.elem_faces = inv.reshape(nt, len(lf))
        else:  # 2D: the order-3 "face" function is the cell bubble, numbered by element
            self.faces = se.copy()
            self.elem_faces = np.arange(nt).reshape(nt, 1)
        self.ne, self.nf, self.nt = self.edges.shape[0], self.faces.shape[0], nt
        p = order
        self.edge_base = self.nv
        self.face_base = self.nv + (p - 1) * self.ne
        self.ndof = self.face_base + (self.nf if p == 3 else 0)
        self.nld = n_local_dofs(dim, order)

    def elem_dofs(self):
        p = self.order
        cols = [self.sorted_elems]
        if p >= 2:
            e = self.edge_base + (p - 1) * self.elem_edges
            cols.append(np.stack([e + k for k in range(p - 1)], axis=2).reshape(self.nt, -1))
        if p == 3:
            cols.append(self.face_base + self.elem_faces)
        return np.concatenate(cols, axis=1)

    def edge_id(self, a, b):
        """Global edge number of vertex pairs (a<b), -1 if absent."""
        a = np.atleast_1d(a).astype(np.int64)
        b = np.atleast_1d(b).astype(np.int64)
        key = self.edges[:, 0] * self.nv + self.edges[:, 1]
        q = a * self.nv + b
        pos = np.searchsorted(key, q)
        pos = np.clip(pos, 0, key.shape[0] - 1)
        return np.where(key[pos] == q, pos, -1)

    def face_id(self, a, b, c):
        key = (self.faces[:, 0] * self.nv + self.faces[:, 1]) * self.nv + self.faces[:, 2]
        q = (np.atleast_1d(a).astype(np.int64) * self.nv + np.atleast_1d(b)) * self.nv + np.atleast_1d(c)
        pos = np.clip(np.searchsorted(key, q), 0, key.shape[0] - 1)
        return np.where(key[pos] == q, pos, -1)

    def dirichlet_dofs(self, bfacets, bflag):
        """Constrained dofs: every vertex / edge / face dof of a boundary facet flagged Dirichlet
        (`dirichlet=` argument of ngs.H1, ngsolve_functions.py:27)."""
        mask = np.zeros(self.ndof, dtype=bool)
        bf = np.sort(np.asarray(bfacets, dtype=np.int64)[np.asarray(bflag, dtype=bool)], axis=1)
        if bf.shape[0] == 0:
            return mask
        mask[np.unique(bf)] = True
        p = self.order
        if p >= 2:
            if self.dim == 3:
                pairs = np.concatenate([bf[:, [0, 1]], bf[:, [0, 2]], bf[:, [1, 2]]])
            else:
                pairs = bf
            eid = self.edge_id(pairs[:, 0], pairs[:, 1])
            assert (eid >= 0).all(), "boundary facet edge not in mesh"
            for k in range(p - 1):
                mask[self.edge_base + (p - 1) * eid + k] = True
        if p == 3 and self.dim == 3:
            fid = self.face_id(bf[:, 0], bf[:, 1], bf[:, 2])
            assert (fid >= 0).all(), "boundary facet not a mesh face"
            mask[self.face_base + fid] = True
        return mask


# ----------------------------------------------------------------------------------------------
# element matrices and assembly  (ngsolve_functions.py:31-36, 47)
# ----------------------------------------------------------------------------------------------
def geometry(points, sorted_elems, dim):
    """|K| and the metric G_m = grad l_i . grad l_j for (i<=j)."""
    x = points[sorted_elems]  # nt x (d+1) x d
    J = x[:, 1:, :] - x[:, :1, :]  # rows = edge vectors from vertex 0
    det = np.linalg.det(J)
    vol = np.abs(det) / factorial(dim)
    Jinv = np.linalg.inv(J)  # columns = grad l_1..l_d
    g = np.empty((x.shape[0], dim + 1, dim))
    g[:, 1:, :] = np.swapaxes(Jinv, 1, 2)
    g[:, 0, :] = -g[:, 1:, :].sum(axis=1)
    pairs = metric_pairs(dim)
    G = np.stack([np.einsum("td,td->t", g[:, i, :], g[:, j, :]) for (i, j) in pairs], axis=1)
    return vol, G


def element_matrices(points, space, sigma_elem, chunk=None):
    """K_e for every element, shape nt x nld x nld."""
    dim, order = space.dim, space.order
    vol, G = geometry(points, space.sorted_elems, dim)
    if dim == 3:
        T, _ = reference_tensors(3, order)
        return np.einsum("t,tm,mab->tab", sigma_elem * vol, G, T, optimize=True)
    # 2D axisymmetric: weight 2 pi r, r = x coordinate (ngsolve_functions.py:34), integrated exactly
    Tw, _ = reference_tensors(2, order, weighted=True)
    r = points[space.sorted_elems][:, :, 0]
    return np.einsum("t,tk,tm,kmab->tab", 2 * np.pi * sigma_elem * vol, r, G, Tw, optimize=True)


def assemble(points, space, sigma, mat, chunk=200000):
    """Global stiffness matrix, CSR with sorted columns, all rows kept (Dirichlet rows included)."""
    sigma_elem = np.asarray(sigma, dtype=np.float64)[np.asarray(mat)]
    dofs = space.elem_dofs()
    n = space.ndof
    nld = space.nld
    parts = []
    for lo in range(0, space.nt, chunk):
        hi = min(space.nt, lo + chunk)
        sub = _SubSpace(space, lo, hi)
        Ke = element_matrices(points, sub, sigma_elem[lo:hi])
        d = dofs[lo:hi]
        rows = np.repeat(d, nld, axis=1).reshape(-1)
        cols = np.tile(d, (1, nld)).reshape(-1)
        # coo -> csr sums duplicates and KEEPS entries that are (or cancel to) exactly zero: the pattern is
        # structural, as in a finite element code (sparse "+" would silently drop them)
        parts.append(sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(n, n)).tocsr().tocoo())
    A = sp.coo_matrix((np.concatenate([q.data for q in parts]),
                       (np.concatenate([q.row for q in parts]), np.concatenate([q.col for q in parts]))), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


class _SubSpace:
    def __init__(self, space, lo, hi):
        self.dim, self.order = space.dim, space.order
        self.sorted_elems = space.sorted_elems[lo:hi]


# ----------------------------------------------------------------------------------------------
# axis evaluation: point sources and sampling  (ngsolve_functions.py:10-21, worker.py:122-131)
# ----------------------------------------------------------------------------------------------
class Axis:
    """The chain of mesh edges on x=0 (2D) / x=y=0 (3D), sorted by z (last coordinate)."""

    def __init__(self, points, space, tol=1e-9):
        dim = space.dim
        on_axis = np.all(np.abs(points[:, : dim - 1]) <= tol, axis=1)
        idx = np.nonzero(on_axis)[0]
        order = np.argsort(points[idx, dim - 1], kind="stable")
        self.vertices = idx[order]
        self.z = points[self.vertices, dim - 1]
        self.space = space

    def shape(self, z):
        """Non-zero basis functions at axis point z -> (dofs, values)."""
        sp_ = self.space
        i = int(np.searchsorted(self.z, z))
        # snap to a vertex
        for j in (i - 1, i):
            if 0 <= j < self.z.shape[0] and abs(self.z[j] - z) <= SNAP_TOL:
                return np.array([self.vertices[j]]), np.array([1.0])
        if i == 0 or i == self.z.shape[0]:
            raise ValueError("axis point z=%g outside the mesh" % z)
        v0, v1 = self.vertices[i - 1], self.vertices[i]
        t = (z - self.z[i - 1]) / (self.z[i] - self.z[i - 1])
        # orient from smaller to larger global vertex number
        if v0 < v1:
            a, b, la, lb = v0, v1, 1.0 - t, t
        else:
            a, b, la, lb = v1, v0, t, 1.0 - t
        dofs, vals = [a, b], [la, lb]
        p = sp_.order
        if p >= 2:
            e = int(sp_.edge_id(a, b)[0])
            if e < 0:
                raise ValueError("consecutive axis vertices %d,%d are not joined by a mesh edge" % (a, b))
            dofs.append(sp_.edge_base + (p - 1) * e)
            vals.append(la * lb)
            if p == 3:
                dofs.append(sp_.edge_base + 2 * e + 1)
                vals.append(la * lb * (lb - la))
        return np.array(dofs), np.array(vals)


def point_source_rhs(axis, ndof, positions, facs):
    """AddPointSource for every non-zero source term (ngsolve_functions.py:39-44)."""
    f = np.zeros(ndof)
    for z, fac in zip(positions, facs):
        if fac != 0.0:
            d, s = axis.shape(z)
            np.add.at(f, d, fac * s)
    return f


def sample_axis(axis, u, z):
    d, s = axis.shape(z)
    return float(np.dot(u[d], s))


# ----------------------------------------------------------------------------------------------
# solve  (ngsolve_functions.py:46-56) and Ra  (worker.py:113-134)
# ----------------------------------------------------------------------------------------------
def solve_direct(A, F, constrained):
    """Exact (sparse LU) solution of the Dirichlet problem; F may hold several right-hand sides (columns).
    Static condensation (ngsolve_functions.py:53-56) does not change the solution, so the full system is solved."""
    free = np.nonzero(~constrained)[0]
    Aff = A[free][:, free].tocsc()
    lu = spla.splu(Aff)
    F = np.asarray(F, dtype=np.float64)
    U = np.zeros_like(F)
    U[free] = lu.solve(F[free])
    return U


def jacobi_pcg(A, f, constrained, rtol=1e-10, maxit=100000):
    """Reference PCG (x0 = 0, Jacobi = NGSolve 'local' preconditioner), stop on ||r||_2 <= rtol ||b||_2.
    Returns (u, iterations, relres).  Used for the CPU baseline timing and iteration-count checks."""
    free = ~constrained
    d = A.diagonal()
    dinv = np.where(free & (d > 0), 1.0 / np.where(d != 0, d, 1.0), 0.0)
    b = np.where(free, f, 0.0)
    x = np.zeros_like(b)
    r = b.copy()
    z = dinv * r
    p = z.copy()
    rz = r @ z
    bnorm = np.sqrt(b @ b)
    if bnorm == 0:
        return x, 0, 0.0
    it = 0
    for it in range(1, maxit + 1):
        q = A @ p
        q[~free] = 0.0
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        rn = np.sqrt(r @ r)
        if rn <= rtol * bnorm:
            break
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it, float(np.sqrt(r @ r) / bnorm)


def two_level_pcg(A, F, constrained, nv, rtol=1e-10, maxit=5000):
    """PCG with the reference's DEFAULT preconditioner, `ngs.Preconditioner(a, "multigrid")`
    (`remo3d.py:82`, `ngsolve_functions.py:46`).  On a single-level mesh NGSolve's multigrid is a two-level method [NGS]:
    an exact sparse factorisation of the lowest-order (P1) problem plus a smoother on the high-order dofs.  In the
    hierarchical basis the P1 matrix is the leading nv x nv block, so the restatement is
        z_vertex = A_vv^-1 r_vertex   (sparse LU of the free vertex block, factorised once per mesh)
        z_high   = D^-1 r_high        (Jacobi on the edge / face dofs)
    applied additively (symmetric, positive definite -> plain PCG).  All columns of F advance in lock-step with their own
    alpha / beta (a converged column is frozen), zero start, stop on ||r||_2 <= rtol ||b||_2 per column.
    Returns (U, iterations per column, relres per column)."""
    free = ~np.asarray(constrained, dtype=bool)
    F = np.asarray(F, dtype=np.float64)
    if F.ndim == 1:
        F = F[:, None]
    B = np.where(free[:, None], F, 0.0)
    fv = np.nonzero(free[:nv])[0]
    lu = spla.splu(A[fv][:, fv].tocsc())
    d = A.diagonal()
    dinv = np.where(free & (d > 0), 1.0 / np.where(d != 0, d, 1.0), 0.0)

    def precond(R):
        Z = dinv[:, None] * R
        Z[fv] = lu.solve(np.ascontiguousarray(R[fv]))
        return Z

    k = B.shape[1]
    X = np.zeros_like(B)
    R = B.copy()
    Z = precond(R)
    P = Z.copy()
    rz = np.einsum("ij,ij->j", R, Z)
    bb = np.einsum("ij,ij->j", B, B)
    active = bb > 0
    iters = np.zeros(k, dtype=np.int64)
    rr = bb.copy()
    for _ in range(maxit):
        if not active.any():
            break
        Q = A @ P
        Q[~free] = 0.0
        pq = np.einsum("ij,ij->j", P, Q)
        alpha = np.where(active & (pq > 0), rz / np.where(pq != 0, pq, 1.0), 0.0)
        X += P * alpha
        R -= Q * alpha
        rr = np.einsum("ij,ij->j", R, R)
        iters += active
        active = active & (rr > rtol * rtol * bb)
        Z = precond(R)
        rzn = np.einsum("ij,ij->j", R, Z)
        beta = np.where(active, rzn / np.where(rz != 0, rz, 1.0), 0.0)
        P = Z + P * beta
        rz = rzn
    return X, iters, np.sqrt(rr / np.where(bb > 0, bb, 1.0))


def apparent_resistivity(axis, u, k, z0, z1=None, scale=1.0):
    """worker.py:117-131: |K (u(z1) - u(z0))| with two potential electrodes (ascending z),
    |K u(z0)| with one; `scale` = 0.5 on the 3D half-ball."""
    if z1 is None or (isinstance(z1, float) and z1 != z1):
        return abs(k * sample_axis(axis, u, z0)) * scale
    return abs(k * (sample_axis(axis, u, z1) - sample_axis(axis, u, z0))) * scale


def solve_task(points, elems, mat, sigma, bfacets, bflag, order, flat, dim=3, solver="auto", rtol=1e-13):
    """One mesh task end to end: assemble once, all right-hand sides, all log points.
    `flat` is the dict produced by remo3d_b200.planner.flatten_task (plain arrays)."""
    space = Space(points.shape[0], elems, order, dim)
    A = assemble(points, space, sigma, mat)
    con = space.dirichlet_dofs(bfacets, bflag)
    axis = Axis(points, space)
    nrhs = flat["src_ptr"].shape[0] - 1
    F = np.zeros((space.ndof, nrhs))
    for r in range(nrhs):
        lo, hi = flat["src_ptr"][r], flat["src_ptr"][r + 1]
        F[:, r] = point_source_rhs(axis, space.ndof, flat["src_z"][lo:hi], flat["src_fac"][lo:hi])
    if solver == "auto":  # sparse LU fill-in explodes on 3D high-order systems
        solver = "direct" if space.ndof <= 30000 else "pcg"
    iters = None
    if solver == "direct":
        U = solve_direct(A, F, con)
    elif solver == "multigrid":  # the reference's default preconditioner, all right-hand sides in one block
        U, iters, _ = two_level_pcg(A, F, con, points.shape[0], rtol=rtol)
    else:
        U = np.stack([jacobi_pcg(A, F[:, r], con, rtol=rtol)[0] for r in range(nrhs)], axis=1)
    ra = np.array([
        apparent_resistivity(axis, U[:, flat["pt_rhs"][i]], flat["pt_k"][i], flat["pt_z0"][i], flat["pt_z1"][i], flat["scale"])
        for i in range(flat["pt_rhs"].shape[0])
    ])
    return {"space": space, "A": A, "constrained": con, "U": U, "ra": ra, "axis": axis, "iters": iters}
