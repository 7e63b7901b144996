import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools/precond_study")
import bench
from oracle import fem_oracle as fo
from remo3d_b200 import meshgen
from pc_common import pcg

def iters(m, flat, tag):
    pts, elems, mat = m["points"], m["elems"], m["mat"]
    space = fo.Space(pts.shape[0], elems, 2, 3)
    A = fo.assemble(pts, space, bench.SIGMA, mat).tocsr()
    con = np.asarray(space.dirichlet_dofs(m["bfacets"], m["bdir"]), bool)
    axis = fo.Axis(pts, space)
    nrhs = 2
    F = np.zeros((space.ndof, nrhs))
    for r in range(nrhs):
        lo, hi = flat["src_ptr"][r], flat["src_ptr"][r+1]
        F[:, r] = fo.point_source_rhs(axis, space.ndof, flat["src_z"][lo:hi], flat["src_fac"][lo:hi])
    free = np.where(~con)[0]
    Af = A[free][:, free].tocsr(); Ff = F[free]
    nv = pts.shape[0]
    isv = free < nv
    iv = np.where(isv)[0]; ie = np.where(~isv)[0]
    lu = spla.splu(Af[iv][:, iv].tocsc())
    dinv = 1.0 / Af.diagonal()
    def M(R):
        Z = np.empty_like(R)
        Z[iv] = lu.solve(R[iv]); Z[ie] = R[ie] * dinv[ie, None]
        return Z
    t0 = time.time()
    X, it = pcg(Af, Ff, M)
    def Mj(R): return R * dinv[:, None]
    q = meshgen._quality(pts, elems)
    print("%s: ndof %d nt %d  iters(exact P1 + Jacobi) %d  [%.0fs]  quality min %.3g  <0.05: %d  <0.1: %d  <0.2: %d" % (tag, space.ndof, len(elems), it, time.time()-t0, q.min(), (q<0.05).sum(), (q<0.1).sum(), (q<0.2).sum()))
    return it

if __name__ == "__main__":
    size = sys.argv[1]
    task, flat = bench.make_task()
    m = bench.make_mesh(size, task, print)
    iters(m, flat, "plain " + size)
