"""Two-level preconditioner variants on the bench matrix (CPU, SciPy): the numbers behind DESIGN.md section 4
("Jacobi 371 iterations; hierarchical splitting with an EXACT P1 solve 136; exact P1 + exact edge-block solve 103;
multiplicative cycle 69 but 3 SpMM per iteration", 60 k-dof mesh, 2 right-hand sides, PCG to 1e-10).

    python tools/precond_study/build_matrix.py 60k          # once: assembles the matrix with the oracle into /tmp/study
    python tools/precond_study/two_level_variants.py 60k
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pc_common import pcg  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "60k"
A = sp.load_npz("/tmp/study/A_%s.npz" % size).tocsr()
aux = np.load("/tmp/study/aux_%s.npz" % size)
free, F, nv = aux["free"], aux["F"], int(aux["nv"])
idx = np.where(free)[0]
Af = A[idx][:, idx].tocsr()
Ff = F[idx][:, :2].copy()
nvf = int((idx < nv).sum())
d = Af.diagonal()
lu_v = spl.splu(Af[:nvf][:, :nvf].tocsc())
Aee = Af[nvf:][:, nvf:].tocsr()
lu_e = spl.splu(Aee.tocsc())
de = Aee.diagonal()
L, U = sp.tril(Af, 0).tocsr(), sp.triu(Af, 0).tocsr()


def report(name, M):
    t = time.time()
    _, it = pcg(Af, Ff, M)
    print("%-58s %4d iterations  %.1f s" % (name, it, time.time() - t), flush=True)


def coarse(R):
    Z = np.zeros_like(R)
    Z[:nvf] = lu_v.solve(R[:nvf])
    return Z


def split(edge):  # additive hierarchical splitting: exact P1 block + `edge` on the high-order block
    def M(R):
        Z = np.empty_like(R)
        Z[:nvf] = lu_v.solve(R[:nvf])
        Z[nvf:] = edge(R[nvf:])
        return Z
    return M


def multiplicative_sgs(R):  # forward GS, coarse correction, backward GS: 3 products with A per application
    x = spl.spsolve_triangular(L, R, lower=True)
    x = x + coarse(R - Af @ x)
    return x + spl.spsolve_triangular(U, R - Af @ x, lower=False)


report("Jacobi", lambda R: R / d[:, None])
report("exact P1 + Jacobi on the edges (the reference's default)", split(lambda R: R / de[:, None]))
report("exact P1 + exact edge block", split(lu_e.solve))
report("multiplicative symmetric GS + exact P1", multiplicative_sgs)
