import sys, time, numpy as np, scipy.sparse as sp
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from pc_common import pcg
import pc3lib as L3  # noqa
size = sys.argv[1]
S = L3.load(size)
Af, Ff, nvf, d, Avv, vpts = S
lv = L3.build_plain(Avv, vpts)
def cyc_add0(levels, b, alpha=1.5, w0=1.0):
    L = levels[0]
    xc = L3.cycle(levels, 1, L.P.T @ b, sweeps=1, alpha=alpha)
    return w0 * L.l1[:, None] * b + alpha * (L.P @ xc)
def make(fn):
    def M(R):
        Z = R / d[:, None]; Z[:nvf] = fn(R[:nvf]); return Z
    return M
t=time.time(); X, it = pcg(Af, Ff, L3.make_M(S, lv, sweeps=1, alpha=1.5)); print("V(1,1) all levels", it, time.time()-t, flush=True)
for w0, al in ((1.0,1.5),(1.0,1.0),(1.5,1.5),(2.0,2.0)):
    t=time.time(); X, it = pcg(Af, Ff, make(lambda b: cyc_add0(lv, b, alpha=al, w0=w0))); print("additive level 0 w0=%g alpha=%g" % (w0, al), it, time.time()-t, flush=True)
