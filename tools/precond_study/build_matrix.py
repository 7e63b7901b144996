import sys, os, time, numpy as np, scipy.sparse as sp
sys.path.insert(0, "/root/repo")
import bench
from oracle import fem_oracle as fo
size = sys.argv[1]
task, flat = bench.make_task()
m = bench.make_mesh(size, task, print)
pts, elems, mat = m["points"], m["elems"], m["mat"]
t0=time.time()
space = fo.Space(pts.shape[0], elems, 2, 3)
A = fo.assemble(pts, space, bench.SIGMA, mat).tocsr()
con = space.dirichlet_dofs(m["bfacets"], m["bdir"])
axis = fo.Axis(pts, space)
nrhs = flat["src_ptr"].shape[0]-1
F = np.zeros((space.ndof, nrhs))
for r in range(nrhs):
    lo, hi = flat["src_ptr"][r], flat["src_ptr"][r+1]
    F[:, r] = fo.point_source_rhs(axis, space.ndof, flat["src_z"][lo:hi], flat["src_fac"][lo:hi])
print("ndof", space.ndof, "nnz", A.nnz, "nv", pts.shape[0], time.time()-t0)
free = ~np.asarray(con, bool)
sp.save_npz("/tmp/study/A_%s.npz" % size, A)
np.savez("/tmp/study/aux_%s.npz" % size, free=free, F=F, nv=pts.shape[0], pts=pts)
