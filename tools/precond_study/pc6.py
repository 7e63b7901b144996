import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from pc_common import pcg
import pc3lib as L3  # noqa
size = sys.argv[1]
which = sys.argv[2].split(",")
S = L3.load(size)
Af, Ff, nvf, d, Avv, vpts = S
n = Af.shape[0]
lu = spl.splu(Avv.tocsc())
def coarse_add(x, r):
    x[:nvf] += lu.solve(r[:nvf]); return x
l1 = 1.0/np.asarray(abs(Af).sum(1)).ravel()
def mult(smooth_pre, smooth_post):
    def M(R):
        x = smooth_pre(R)
        r1 = R - Af @ x
        x = coarse_add(x, r1)
        r2 = R - Af @ x
        return x + smooth_post(r2)
    return M
def jac_l1(nu):
    def s(R):
        x = l1[:,None]*R
        for i in range(1,nu): x = x + l1[:,None]*(R - Af@x)
        return x
    return s
Dm = 1/np.sqrt(d)
def cheb(deg, lmax, ratio):
    lmin = lmax/ratio
    theta = (lmax+lmin)/2; delta = (lmax-lmin)/2
    def app(R):
        b = R*Dm[:,None]
        x = np.zeros_like(b); r = b.copy()
        sigma = theta/delta; rho = 1/sigma
        dd = r/theta
        for k in range(deg):
            x = x + dd
            if k == deg-1: break
            r = r - Dm[:,None]*(Af @ (Dm[:,None]*dd))
            rho_n = 1/(2*sigma - rho)
            dd = rho_n*rho*dd + 2*rho_n/delta*r
            rho = rho_n
        return x*Dm[:,None]
    return app
if "sgs" in which:
    L = sp.tril(Af, 0).tocsr(); U = sp.triu(Af, 0).tocsr()
    t=time.time(); X, it = pcg(Af, Ff, mult(lambda R: spl.spsolve_triangular(L, R, lower=True), lambda R: spl.spsolve_triangular(U, R, lower=False))); print("mult SGS", it, time.time()-t, flush=True)
if "l1" in which:
  for nu in (1,2):
    t=time.time(); X, it = pcg(Af, Ff, mult(jac_l1(nu), jac_l1(nu))); print("mult l1-jacobi nu=%d"%nu, it, "spmm/it", 1+2+2*(nu-1), time.time()-t, flush=True)
if "cheb" in which:
  from scipy.sparse.linalg import eigsh
  B = (sp.diags(Dm) @ Af @ sp.diags(Dm)).tocsr()
  lmax = float(eigsh(B, k=1, which="LA", return_eigenvectors=False, tol=1e-3)[0]); print("lmax", lmax)
  for deg, ratio in ((2,6),(3,10),(4,15),(3,20)):
    c = cheb(deg, lmax*1.05, ratio)
    t=time.time(); X, it = pcg(Af, Ff, mult(c, c)); print("mult cheb deg=%d ratio=%g"%(deg,ratio), it, "spmm/it", 1+2+2*(deg-1), time.time()-t, flush=True)
