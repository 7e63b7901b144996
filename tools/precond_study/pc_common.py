import numpy as np
def pcg(Aop, B, M, rtol=1e-10, maxit=5000):
    X = np.zeros_like(B); R = B.copy(); Z = M(R); P = Z.copy()
    rz = (R*Z).sum(0); bb = (B*B).sum(0)
    for it in range(1, maxit+1):
        Q = Aop @ P
        a = rz / (P*Q).sum(0)
        X += P*a; R -= Q*a
        rr = (R*R).sum(0)
        if np.all(rr <= rtol**2*bb): return X, it
        Z = M(R); rzn = (R*Z).sum(0)
        P = Z + P*(rzn/rz); rz = rzn
    return X, maxit
