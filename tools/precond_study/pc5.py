import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from pc_common import pcg
import pc3lib as L3  # noqa
size = sys.argv[1]
S = L3.load(size)
Af, Ff, nvf, d, Avv, vpts = S
n = Af.shape[0]
Aee = Af[nvf:][:, nvf:].tocsr()
lu = spl.splu(Avv.tocsc())
de = Aee.diagonal()
# spectrum of D^-1 Aee
from scipy.sparse.linalg import eigsh, LinearOperator
Dm = sp.diags(1/np.sqrt(de))
B = (Dm @ Aee @ Dm).tocsr()
lmax = eigsh(B, k=3, which="LA", return_eigenvectors=False, tol=1e-3)
lmin = eigsh(B, k=5, which="SA", return_eigenvectors=False, tol=1e-3, maxiter=20000)
print("edge block Jacobi-scaled: lmax", lmax, "lmin", lmin)
# preconditioned full operator extreme eigenvalues via pcg's Lanczos? use eigsh on symmetrized M^-1/2 A M^-1/2 not available; skip
Le = sp.tril(Aee,0).tocsr(); Ue = sp.triu(Aee,0).tocsr()
def hier_sgs_e(R):
    Z = np.empty_like(R); Z[:nvf] = lu.solve(R[:nvf])
    y = spl.spsolve_triangular(Le, R[nvf:], lower=True)
    Z[nvf:] = spl.spsolve_triangular(Ue, y*de[:,None], lower=False)
    return Z
t=time.time(); X, it = pcg(Af, Ff, hier_sgs_e); print("hier exact P1 + SGS edges", it, time.time()-t, flush=True)
def cheb_e(deg, lmax, lmin):
    # Chebyshev polynomial approx of Aee^-1 with Jacobi scaling
    theta = (lmax+lmin)/2; delta = (lmax-lmin)/2
    def app(R):
        # solves B y = Dm R  ; z = Dm y
        b = R * (1/np.sqrt(de))[:,None]
        x = np.zeros_like(b); r = b.copy()
        sigma = theta/delta; rho = 1/sigma
        dd = r/theta
        for k in range(deg):
            x = x + dd
            if k == deg-1: break
            r = r - B @ dd
            rho_n = 1/(2*sigma - rho)
            dd = rho_n*rho*dd + 2*rho_n/delta*r
            rho = rho_n
        return x * (1/np.sqrt(de))[:,None]
    return app
for deg in (2,3):
    ce = cheb_e(deg, float(lmax.max())*1.05, float(lmax.max())/ (8 if deg==2 else 15))
    def M(R):
        Z = np.empty_like(R); Z[:nvf] = lu.solve(R[:nvf]); Z[nvf:] = ce(R[nvf:]); return Z
    t=time.time(); X, it = pcg(Af, Ff, M); print("hier exact P1 + cheb%d edges"%deg, it, time.time()-t, flush=True)
lue = spl.splu(Aee.tocsc())
def hier_ee(R):
    Z = np.empty_like(R); Z[:nvf] = lu.solve(R[:nvf]); Z[nvf:] = lue.solve(R[nvf:]); return Z
t=time.time(); X, it = pcg(Af, Ff, hier_ee); print("hier exact/exact", it, time.time()-t)
