"""Edge-block variants of the hierarchical split: point Jacobi vs block Jacobi with one block per first vertex."""
import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
sys.path.insert(0, "/root/repo")
from pc_common import pcg
import pc3lib as L3
import bench
from oracle import fem_oracle as fo
size = sys.argv[1]
S = L3.load(size)
Af, Ff, nvf, d, Avv, vpts = S
aux = np.load("/tmp/study/aux_%s.npz" % size)
free, nv = aux["free"], int(aux["nv"])
task, flat = bench.make_task()
m = bench.make_mesh(size, task, print)
space = fo.Space(m["points"].shape[0], m["elems"], 2, 3)
edges = space.edges
idx = np.where(free)[0]
eidx = idx[nvf:] - nv            # edge number of every free edge dof
lu = spl.splu(Avv.tocsc())
Aee = Af[nvf:][:, nvf:].tocsr()
ne = Aee.shape[0]
def block_inverse(groups):
    # groups: array of group id per free edge dof -> block-diagonal inverse as sparse matrix
    order = np.argsort(groups, kind="stable"); g = groups[order]
    starts = np.r_[0, np.flatnonzero(np.diff(g)) + 1, len(g)]
    rows, cols, vals = [], [], []
    Ac = Aee.tocsc()
    for a, b in zip(starts[:-1], starts[1:]):
        ids = order[a:b]
        blk = Aee[ids][:, ids].toarray()
        inv = np.linalg.inv(blk)
        rr, cc = np.meshgrid(ids, ids, indexing="ij")
        rows.append(rr.ravel()); cols.append(cc.ravel()); vals.append(inv.ravel())
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(ne, ne)), np.diff(starts)
def make(Binv):
    def M(R):
        Z = np.empty_like(R); Z[:nvf] = lu.solve(R[:nvf]); Z[nvf:] = Binv @ R[nvf:]; return Z
    return M
t = time.time(); X, it = pcg(Af, Ff, make(sp.diags(1 / Aee.diagonal()))); print("exact P1 + point Jacobi", it, time.time() - t, flush=True)
for name, grp in (("first vertex", edges[eidx, 0]), ("second vertex", edges[eidx, 1])):
    Binv, sizes = block_inverse(grp.astype(np.int64))
    t = time.time(); X, it = pcg(Af, Ff, make(Binv)); print("exact P1 + block Jacobi by %s (mean block %.1f, max %d)" % (name, sizes.mean(), sizes.max()), it, time.time() - t, flush=True)
