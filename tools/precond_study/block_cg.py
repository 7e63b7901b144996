"""Block PCG against lock-step PCG on the C4 bench mesh with the two-level preconditioner (CPU, SciPy): the negative result in
profiles/r02_notes.md section 3 (96 vs 105 iterations at 202 225 dofs, 5 sources 5 cm apart).

    python tools/precond_study/block_cg.py 200k
"""
import os, sys, time
import numpy as np, scipy.sparse.linalg as spla, scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from oracle import fem_oracle as fo
size = sys.argv[1]
task, flat = bench.make_task()
m = bench.make_mesh(size, task, print)
pts, elems, mat = m["points"], m["elems"], m["mat"]
space = fo.Space(pts.shape[0], elems, 2, 3)
A = fo.assemble(pts, space, bench.SIGMA, mat)
con = space.dirichlet_dofs(m["bfacets"], m["bdir"].astype(bool))
axis = fo.Axis(pts, space)
nrhs = flat["src_ptr"].shape[0] - 1
F = np.zeros((space.ndof, nrhs))
for r in range(nrhs):
    lo, hi = flat["src_ptr"][r], flat["src_ptr"][r + 1]
    F[:, r] = fo.point_source_rhs(axis, space.ndof, flat["src_z"][lo:hi], flat["src_fac"][lo:hi])
nv = pts.shape[0]
print('ndof', space.ndof, 'nrhs', nrhs, 'src z', flat["src_z"])
free = ~np.asarray(con, bool)
B = np.where(free[:, None], F, 0.0)
fv = np.nonzero(free[:nv])[0]
lu = spla.splu(A[fv][:, fv].tocsc())
d = A.diagonal(); dinv = np.where(free & (d > 0), 1.0 / np.where(d != 0, d, 1.0), 0.0)
def precond(R):
    Z = dinv[:, None] * R
    Z[fv] = lu.solve(np.ascontiguousarray(R[fv]))
    return Z
def Amul(P):
    Q = A @ P; Q[~free] = 0.0; return Q
rtol = 1e-10
# --- standard
t=time.time(); X, it, rr = fo.two_level_pcg(A, F, con, nv, rtol=rtol); print('lock-step PCG iters', it, 'relres', rr.max(), '%.1fs'%(time.time()-t))
# --- block CG (Dubrulle-R variant, BCGrQ): R = Q C, orthonormalise residual block in the M^-1 inner product? use standard O'Leary with QR of P
def block_pcg(B, maxit=1000):
    k = B.shape[1]
    X = np.zeros_like(B); R = B.copy(); Z = precond(R); P = Z.copy()
    bb = np.sqrt(np.einsum('ij,ij->j', B, B))
    RZ = R.T @ Z
    for itn in range(1, maxit+1):
        Q = Amul(P)
        PQ = P.T @ Q
        alpha = np.linalg.solve(PQ, RZ)
        X += P @ alpha
        R -= Q @ alpha
        rel = np.sqrt(np.einsum('ij,ij->j', R, R)) / bb
        if rel.max() <= rtol: return X, itn, rel
        Z = precond(R)
        RZn = R.T @ Z
        beta = np.linalg.solve(RZ, RZn)
        P = Z + P @ beta
        RZ = RZn
        if itn % 20 == 0: print('  it', itn, 'relres max %.2e min %.2e cond(PQ) %.1e'%(rel.max(), rel.min(), np.linalg.cond(PQ)))
    return X, maxit, rel
t=time.time(); Xb, itb, relb = block_pcg(B); print('block PCG iters', itb, 'relres', relb, '%.1fs'%(time.time()-t))
print('solution diff', np.abs(Xb-X).max()/np.abs(X).max())
# --- block CG with orthonormalised search block (A-orthonormalise P each step: P <- P L^-T where P^T A P = L L^T)
def block_pcg_orth(B, maxit=1000):
    X = np.zeros_like(B); R = B.copy(); Z = precond(R); P = Z.copy()
    bb = np.sqrt(np.einsum('ij,ij->j', B, B))
    for itn in range(1, maxit+1):
        Q = Amul(P)
        PQ = P.T @ Q
        L = np.linalg.cholesky((PQ + PQ.T) / 2)
        Li = np.linalg.inv(L)
        P = P @ Li.T; Q = Q @ Li.T      # now P^T A P = I
        alpha = P.T @ R
        X += P @ alpha
        R -= Q @ alpha
        rel = np.sqrt(np.einsum('ij,ij->j', R, R)) / bb
        if rel.max() <= rtol: return X, itn, rel
        Z = precond(R)
        beta = -(Q.T @ Z)
        P = Z + P @ beta
    return X, maxit, rel
t=time.time(); Xo, ito, relo = block_pcg_orth(B); print('block PCG (A-orthonormal P) iters', ito, 'relres', relo, '%.1fs'%(time.time()-t))
print('solution diff', np.abs(Xo-X).max()/np.abs(X).max())
