import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import pc3lib as L3  # noqa
size = sys.argv[1]
S = L3.load(size)
Af, Ff, nvf, d, Avv, vpts = S
lu = spl.splu(Avv.tocsc())
def M(R):
    Z = R / d[:, None]; Z[:nvf] = lu.solve(R[:nvf]); return Z
B = Ff[:, :1].copy()
X = np.zeros_like(B); R = B.copy(); Z = M(R); P = Z.copy()
rz = (R*Z).sum(); bb = (B*B).sum()
al, be = [], []
for it in range(1, 400):
    Q = Af @ P
    a = rz / (P*Q).sum(); al.append(a)
    X += P*a; R -= Q*a
    if (R*R).sum() <= 1e-20*bb: break
    Z = M(R); rzn = (R*Z).sum(); b = rzn/rz; be.append(b)
    P = Z + P*b; rz = rzn
m = len(al)
T = np.zeros((m, m))
for j in range(m):
    T[j, j] = 1/al[j] + (be[j-1]/al[j-1] if j > 0 else 0)
    if j+1 < m: T[j, j+1] = T[j+1, j] = np.sqrt(be[j])/al[j]
ev = np.linalg.eigvalsh(T)
print("its", m, "ritz min", ev[:12], "max", ev[-6:], "kappa", ev[-1]/ev[0])
print("count below 0.05:", (ev<0.05).sum(), "below 0.1", (ev<0.1).sum(), "below 0.2", (ev<0.2).sum())
