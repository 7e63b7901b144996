import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, "/root/repo")
size = sys.argv[1]
A = sp.load_npz("/tmp/study/A_%s.npz" % size).tocsr()
aux = np.load("/tmp/study/aux_%s.npz" % size)
nv, pts = int(aux["nv"]), aux["pts"]
n = A.shape[0]
# need edge midpoints: rebuild space
import bench
from oracle import fem_oracle as fo
task, flat = bench.make_task()
m = bench.make_mesh(size, task, print)
space = fo.Space(m["points"].shape[0], m["elems"], 2, 3)
edges = space.edges  # (ne,2)
loc = np.vstack([pts, 0.5*(pts[edges[:,0]] + pts[edges[:,1]])])
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
code = morton(loc)
cls = (np.arange(n) >= nv).astype(np.uint64)
for name, order in (("natural", np.arange(n)), ("class-major morton", np.lexsort((code, cls))), ("pure morton", np.argsort(code, kind="stable"))):
    Ao = A[order]
    for B in (64, 128, 256, 512, 1024):
        fp, nz = [], []
        for b0 in range(0, n, B):
            sub = Ao[b0:b0+B]
            fp.append(len(np.unique(sub.indices))); nz.append(sub.nnz)
        fp = np.array(fp); nz = np.array(nz)
        nvb = (nv + B - 1)//B if name != "pure morton" else 0
        print("%-20s B=%4d  footprint mean %6.0f max %6d  (vertex blocks mean %6.0f, edge blocks mean %6.0f)  total loads/nnz %.3f  KB mean %.0f" % (
            name, B, fp.mean(), fp.max(), fp[:nvb].mean() if nvb else 0, fp[nvb:].mean(), fp.sum()/nz.sum(), fp.mean()*64/1024))
