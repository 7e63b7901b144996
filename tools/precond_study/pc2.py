import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
size = sys.argv[1]
A = sp.load_npz("/tmp/study/A_%s.npz" % size).tocsr()
aux = np.load("/tmp/study/aux_%s.npz" % size)
free, F, nv, pts = aux["free"], aux["F"], int(aux["nv"]), aux["pts"]
idx = np.where(free)[0]
Af = A[idx][:, idx].tocsr()
Ff = F[idx][:, :2].copy()
nvf = (idx < nv).sum()
d = Af.diagonal()
Avv = Af[:nvf][:, :nvf].tocsc()
Aee = Af[nvf:][:, nvf:].tocsc()
from pc_common import pcg
lu = spl.splu(Avv)
n = Af.shape[0]
P = sp.eye(n, nvf, format="csr")  # injection
def coarse(R):
    Z = np.zeros_like(R); Z[:nvf] = lu.solve(R[:nvf]); return Z
# symmetric GS two-level
L = sp.tril(Af, 0).tocsr(); U = sp.triu(Af, 0).tocsr()
def sgs_twolevel(R):
    x = spl.spsolve_triangular(L, R, lower=True)
    r1 = R - Af @ x
    x = x + coarse(r1)
    r2 = R - Af @ x
    x = x + spl.spsolve_triangular(U, r2, lower=False)
    return x
t=time.time(); X, it = pcg(Af, Ff, sgs_twolevel); print("multiplicative SGS + exact P1", it, time.time()-t)
# additive: SGS on full + coarse
def sgs(R):
    y = spl.spsolve_triangular(L, R, lower=True)
    return spl.spsolve_triangular(U, y*d[:,None], lower=False)
def add_sgs(R): return sgs(R) + coarse(R)
t=time.time(); X, it = pcg(Af, Ff, add_sgs); print("additive SGS(full) + exact P1", it, time.time()-t)
# exact-exact hierarchical
lue = spl.splu(Aee)
def hier_ee(R):
    Z = np.empty_like(R); Z[:nvf] = lu.solve(R[:nvf]); Z[nvf:] = lue.solve(R[nvf:]); return Z
t=time.time(); X, it = pcg(Af, Ff, hier_ee); print("hier exact/exact", it, time.time()-t)
