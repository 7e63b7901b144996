"""Additive exact corrections on the worst elements (patch = the 10 dofs of a low-quality tet, or of it and its face neighbours) on top
of the two-level preconditioner, bench mesh after the sliver pass (CPU, SciPy).  Negative result (profiles/r02_notes.md section 3):
after the sliver pass only 19 of 143 405 tets have quality < 0.1 at 202 k dofs and the patches only add over-counted corrections --
105 iterations without, 108-142 with.

    python tools/precond_study/patch_correction.py 200k
"""
import os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from oracle import fem_oracle as fo
from remo3d_b200 import meshgen
size = sys.argv[1]
task, flat = bench.make_task()
m = bench.make_mesh(size, task, print)
pts, elems, mat = m["points"], m["elems"], m["mat"]
space = fo.Space(pts.shape[0], elems, 2, 3)
A = fo.assemble(pts, space, bench.SIGMA, mat).tocsr()
con = space.dirichlet_dofs(m["bfacets"], m["bdir"].astype(bool))
axis = fo.Axis(pts, space)
F = np.zeros((space.ndof, 2))
for r in range(2):
    lo, hi = flat["src_ptr"][r], flat["src_ptr"][r + 1]
    F[:, r] = fo.point_source_rhs(axis, space.ndof, flat["src_z"][lo:hi], flat["src_fac"][lo:hi])
nv = pts.shape[0]
free = ~np.asarray(con, bool)
B = np.where(free[:, None], F, 0.0)
fv = np.nonzero(free[:nv])[0]
lu = spla.splu(A[fv][:, fv].tocsc())
d = A.diagonal(); dinv = np.where(free & (d > 0), 1.0 / np.where(d != 0, d, 1.0), 0.0)
edofs = space.elem_dofs()
qual = meshgen._quality(pts, space.sorted_elems)
print('ndof', space.ndof, 'tets', elems.shape[0], 'quality: <0.02', (qual<0.02).sum(), '<0.05', (qual<0.05).sum(), '<0.1', (qual<0.1).sum(), '<0.2', (qual<0.2).sum(), 'min', qual.min())
def pcg(M, rtol=1e-10, maxit=2000):
    X = np.zeros_like(B); R = B.copy(); Z = M(R); P = Z.copy()
    rz = (R*Z).sum(0); bb = (B*B).sum(0)
    for it in range(1, maxit+1):
        Q = A @ P; Q[~free] = 0
        a = rz / (P*Q).sum(0)
        X += P*a; R -= Q*a
        if np.all((R*R).sum(0) <= rtol**2*bb): return it
        Z = M(R); rzn = (R*Z).sum(0)
        P = Z + P*(rzn/rz); rz = rzn
    return maxit
def base(R):
    Z = dinv[:, None] * R
    Z[fv] = lu.solve(np.ascontiguousarray(R[fv]))
    return Z
print('two-level (exact P1 + Jacobi):', pcg(base), flush=True)
def with_patches(q0, mode):
    bad = np.nonzero(qual < q0)[0]
    patches = []
    if mode == 'tet':
        for t in bad:
            dofs = edofs[t][free[edofs[t]]]
            if dofs.size: patches.append(dofs)
    else:  # edge dofs of all tets sharing a vertex with the bad tet... here: all dofs of the bad tet's face neighbours too
        v2t = {}
        se = space.sorted_elems
        for t in bad:
            vs = set(se[t])
            nb = np.nonzero(np.isin(se, list(vs)).sum(axis=1) >= 3)[0]   # tets sharing a face
            dofs = np.unique(edofs[nb].ravel()); dofs = dofs[free[dofs]]
            patches.append(dofs)
    invs = [np.linalg.inv(A[p][:, p].toarray()) for p in patches]
    # only the HIGH-ORDER part of the patch correction is used on top of the exact P1 (keep it additive & SPD)
    def M(R):
        Z = base(R)
        for p, Ai in zip(patches, invs):
            Z[p] += Ai @ R[p]
        return Z
    return M, len(patches), (np.mean([len(p) for p in patches]) if patches else 0)
for q0 in (0.05, 0.1, 0.2):
    for mode in ('tet', 'face'):
        M, n, sz = with_patches(q0, mode)
        t=time.time(); it = pcg(M); print('patches q<%.2f %-4s: %5d patches of ~%.0f dofs -> %d iterations (%.0fs)' % (q0, mode, n, sz, it, time.time()-t), flush=True)
