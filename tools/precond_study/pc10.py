"""More work on the coarse levels only (level 0 keeps one sweep): sweeps / cycle index below level 0."""
import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from pc_common import pcg
import pc3lib as L3
size = sys.argv[1]
S = L3.load(size)
Af, Ff, nvf, d, Avv, vpts = S
lv = L3.build_plain(Avv, vpts)
def cyc(levels, l, b, sw0, swc, gamma_c, alpha):
    L = levels[l]
    if l == len(levels) - 1: return L.inv @ b
    sweeps = sw0 if l == 0 else swc
    w = L.l1[:, None]
    x = w * b
    for s in range(1, sweeps): x = x + w * (b - L.A @ x)
    r = b - L.A @ x
    bc = L.P.T @ r
    xc = cyc(levels, l + 1, bc, sw0, swc, gamma_c, alpha)
    for g in range(1, gamma_c):
        xc = xc + cyc(levels, l + 1, bc - levels[l + 1].A @ xc, sw0, swc, gamma_c, alpha)
    x = x + alpha * (L.P @ xc)
    for s in range(sweeps): x = x + w * (b - L.A @ x)
    return x
def make(**kw):
    def M(R):
        Z = R / d[:, None]; Z[:nvf] = cyc(lv, 0, R[:nvf], **kw); return Z
    return M
for kw in (dict(sw0=1, swc=1, gamma_c=1, alpha=1.5), dict(sw0=1, swc=2, gamma_c=1, alpha=1.5), dict(sw0=1, swc=1, gamma_c=2, alpha=1.5),
           dict(sw0=1, swc=2, gamma_c=2, alpha=1.5), dict(sw0=1, swc=3, gamma_c=2, alpha=1.3)):
    t = time.time(); X, it = pcg(Af, Ff, make(**kw)); print(kw, it, "%.0f s" % (time.time() - t), flush=True)
