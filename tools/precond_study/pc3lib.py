import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from pc_common import pcg

def load(size):
    A = sp.load_npz("/tmp/study/A_%s.npz" % size).tocsr()
    aux = np.load("/tmp/study/aux_%s.npz" % size)
    free, F, nv, pts = aux["free"], aux["F"], int(aux["nv"]), aux["pts"]
    idx = np.where(free)[0]
    Af = A[idx][:, idx].tocsr()
    Ff = F[idx][:, :2].copy()
    nvf = int((idx < nv).sum())
    d = Af.diagonal()
    Avv = Af[:nvf][:, :nvf].tocsr()
    vpts = pts[idx[:nvf]]
    return Af, Ff, nvf, d, Avv, vpts


def morton(p):
    lo, hi = p.min(0), p.max(0)
    q = ((p - lo) / (hi - lo) * 2097151).astype(np.uint64)
    def spread(v):
        v = v & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    return spread(q[:,0]) | (spread(q[:,1]) << np.uint64(1)) | (spread(q[:,2]) << np.uint64(2))

class Level: pass
def build_plain(A0, pts0, agg=8, coarsest=256, smooth_P=False, omegaP=0.66):
    levels = []
    A = A0; first = True
    while True:
        L = Level(); L.A = A; L.n = A.shape[0]
        L.l1 = 1.0 / np.asarray(abs(A).sum(1)).ravel()
        L.dinv = 1.0 / A.diagonal()
        levels.append(L)
        if L.n <= coarsest: break
        if first:
            perm = np.argsort(morton(pts0), kind="stable"); first = False
            aggmap = np.empty(L.n, np.int64); aggmap[perm] = np.arange(L.n) // agg
        else:
            aggmap = np.arange(L.n) // agg
        nc = aggmap.max() + 1
        T = sp.csr_matrix((np.ones(L.n), (np.arange(L.n), aggmap)), shape=(L.n, nc))
        if smooth_P:
            Dinv = sp.diags(L.dinv)
            T = (T - omegaP * (Dinv @ (A @ T))).tocsr()
        L.P = T
        A = (T.T @ A @ T).tocsr()
    L.inv = np.linalg.inv(A.toarray())
    return levels

def cycle(levels, l, b, sweeps=1, alpha=1.5, smoother="l1", gamma=1, omega=1.0):
    L = levels[l]
    if l == len(levels) - 1: return L.inv @ b
    w = (L.l1 if smoother == "l1" else L.dinv * omega)[:, None]
    x = w * b
    for s in range(1, sweeps): x = x + w * (b - L.A @ x)
    r = b - L.A @ x
    bc = L.P.T @ r
    xc = cycle(levels, l + 1, bc, sweeps, alpha, smoother, gamma, omega)
    for g in range(1, gamma):
        xc = xc + cycle(levels, l+1, bc - levels[l+1].A @ xc, sweeps, alpha, smoother, gamma, omega)
    x = x + alpha * (L.P @ xc)
    for s in range(sweeps): x = x + w * (b - L.A @ x)
    return x

def make_M(S, levels, **kw):
    Af, Ff, nvf, d, Avv, vpts = S
    def M(R):
        Z = R / d[:, None]
        Z[:nvf] = cycle(levels, 0, R[:nvf], **kw)
        return Z
    return M

