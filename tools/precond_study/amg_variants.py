"""Round-2 study: which V-cycle on the P1 block brings the GPU's "multigrid" iteration count down to the exact-P1-solve
count of the hierarchical split?  SciPy emulation on the sliver-passed bench meshes (matrices from build_matrix.py written
to /tmp/study2).  Usage: python amg_variants.py 200k|1M [variant ...]"""
import sys, time, os
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pc_common import pcg
import pc3lib
from pc3lib import morton, Level

STUDY = os.environ.get("STUDY_DIR", "/tmp/study2")

def load(size):
    A = sp.load_npz("%s/A_%s.npz" % (STUDY, size)).tocsr()
    aux = np.load("%s/aux_%s.npz" % (STUDY, size))
    free, F, nv, pts = aux["free"], aux["F"], int(aux["nv"]), aux["pts"]
    idx = np.where(free)[0]
    Af = A[idx][:, idx].tocsr()
    Ff = F[idx][:, :2].copy()
    nvf = int((idx < nv).sum())
    d = Af.diagonal()
    Avv = Af[:nvf][:, :nvf].tocsr()
    vpts = pts[idx[:nvf]]
    return Af, Ff, nvf, d, Avv, vpts

def strength_graph(A, theta):
    """|a_ij| >= theta * sqrt(a_ii a_jj), off-diagonal, negative couplings only count (M-matrix-like)"""
    A = A.tocoo()
    d = np.asarray(A.tocsr().diagonal())
    m = (A.row != A.col) & (-A.data >= theta * np.sqrt(d[A.row] * d[A.col]))
    S = sp.csr_matrix((-A.data[m], (A.row[m], A.col[m])), shape=A.shape)
    return S

def greedy_agg(A, theta=0.08, order=None):
    """Vanek-style greedy aggregation on the strength graph."""
    n = A.shape[0]
    S = strength_graph(A, theta)
    ip, ix, iv = S.indptr, S.indices, S.data
    agg = np.full(n, -1, np.int64)
    na = 0
    it = range(n) if order is None else order
    for i in it:
        if agg[i] >= 0: continue
        nb = ix[ip[i]:ip[i+1]]
        if nb.size and (agg[nb] >= 0).any(): continue
        agg[i] = na; agg[nb] = na; na += 1
    # pass 2: attach leftovers to the strongest neighbouring aggregate
    left = np.where(agg < 0)[0]
    agg2 = agg.copy()
    for i in left:
        nb = ix[ip[i]:ip[i+1]]; w = iv[ip[i]:ip[i+1]]
        ok = agg[nb] >= 0
        if ok.any():
            agg2[i] = agg[nb[ok][np.argmax(w[ok])]]
    agg = agg2
    for i in np.where(agg < 0)[0]:
        nb = ix[ip[i]:ip[i+1]]
        agg[i] = na; 
        for j in nb:
            if agg[j] < 0: agg[j] = na
        na += 1
    return agg, na

def pairwise_agg(A, passes=3, theta=0.0):
    """passes of strongest-neighbour pairwise matching (2^passes nodes per aggregate at most) on the Galerkin-collapsed graph."""
    n = A.shape[0]
    total = np.arange(n)
    Ac = A.tocsr()
    for p in range(passes):
        m = Ac.shape[0]
        C = Ac.tocoo()
        d = Ac.diagonal()
        off = C.row != C.col
        w = -C.data[off] / np.sqrt(np.abs(d[C.row[off]] * d[C.col[off]]))
        S = sp.csr_matrix((w, (C.row[off], C.col[off])), shape=(m, m))
        ip, ix, iv = S.indptr, S.indices, S.data
        match = np.full(m, -1, np.int64)
        # greedy: visit in order of max strength
        mx = np.array([iv[ip[i]:ip[i+1]].max() if ip[i+1] > ip[i] else -1 for i in range(m)])
        for i in np.argsort(-mx, kind="stable"):
            if match[i] >= 0: continue
            nb = ix[ip[i]:ip[i+1]]; ww = iv[ip[i]:ip[i+1]]
            ok = (match[nb] < 0) & (ww > theta)
            if ok.any():
                j = nb[ok][np.argmax(ww[ok])]
                match[i] = j; match[j] = i
            else:
                match[i] = i
        ids = np.full(m, -1, np.int64); na = 0
        for i in range(m):
            if ids[i] < 0:
                ids[i] = na; ids[match[i]] = na; na += 1
        T = sp.csr_matrix((np.ones(m), (np.arange(m), ids)), shape=(m, na))
        Ac = (T.T @ Ac @ T).tocsr()
        total = ids[total]
    return total, Ac.shape[0]

def build(A0, pts0, kind="morton", agg=8, coarsest=256, smooth_P=False, omegaP=0.66, theta=0.08, passes=3):
    levels = []
    A = A0; first = True
    while True:
        L = Level(); L.A = A; L.n = A.shape[0]
        L.l1 = 1.0 / np.asarray(abs(A).sum(1)).ravel()
        L.dinv = 1.0 / A.diagonal()
        levels.append(L)
        if L.n <= coarsest: break
        if kind == "morton":
            if first:
                perm = np.argsort(morton(pts0), kind="stable")
                aggmap = np.empty(L.n, np.int64); aggmap[perm] = np.arange(L.n) // agg
            else:
                aggmap = np.arange(L.n) // agg
            nc = aggmap.max() + 1
        elif kind == "greedy":
            aggmap, nc = greedy_agg(A, theta)
        elif kind == "pair":
            aggmap, nc = pairwise_agg(A, passes)
        first = False
        T = sp.csr_matrix((np.ones(L.n), (np.arange(L.n), aggmap)), shape=(L.n, nc))
        if smooth_P:
            # filtered-matrix smoothing of the tentative prolongator
            Dinv = sp.diags(L.dinv)
            T = (T - omegaP * (Dinv @ (A @ T))).tocsr()
        L.P = T
        A = (T.T @ A @ T).tocsr()
        if nc >= L.n: break
    L.inv = np.linalg.pinv(A.toarray())
    return levels

def cheb_cycle(levels, l, b, deg=2, alpha=1.0):
    pass

def run(size, names):
    S = load(size)
    Af, Ff, nvf, d, Avv, vpts = S
    print("size %s: n %d nv(free) %d nnz_vv %d" % (size, Af.shape[0], nvf, Avv.nnz))
    variants = {
        "exact": None,
        "cur": dict(b=dict(kind="morton"), c=dict(sweeps=1, alpha=1.5)),
        "cur_s2": dict(b=dict(kind="morton"), c=dict(sweeps=2, alpha=1.5)),
        "cur_a1": dict(b=dict(kind="morton"), c=dict(sweeps=1, alpha=1.0)),
        "cur_a2": dict(b=dict(kind="morton"), c=dict(sweeps=1, alpha=2.0)),
        "cur_W": dict(b=dict(kind="morton"), c=dict(sweeps=1, alpha=1.5, gamma=2)),
        "mort4": dict(b=dict(kind="morton", agg=4), c=dict(sweeps=1, alpha=1.5)),
        "greedy": dict(b=dict(kind="greedy", theta=0.08), c=dict(sweeps=1, alpha=1.5)),
        "greedy_a1": dict(b=dict(kind="greedy", theta=0.08), c=dict(sweeps=1, alpha=1.0)),
        "greedy_t25": dict(b=dict(kind="greedy", theta=0.25), c=dict(sweeps=1, alpha=1.5)),
        "greedy_sa": dict(b=dict(kind="greedy", theta=0.08, smooth_P=True), c=dict(sweeps=1, alpha=1.0)),
        "greedy_sa_s2": dict(b=dict(kind="greedy", theta=0.08, smooth_P=True), c=dict(sweeps=2, alpha=1.0)),
        "pair3": dict(b=dict(kind="pair", passes=3), c=dict(sweeps=1, alpha=1.5)),
        "pair2": dict(b=dict(kind="pair", passes=2), c=dict(sweeps=1, alpha=1.5)),
        "pair3_W": dict(b=dict(kind="pair", passes=3), c=dict(sweeps=1, alpha=1.5, gamma=2)),
        "pair3_W_a1": dict(b=dict(kind="pair", passes=3), c=dict(sweeps=1, alpha=1.0, gamma=2)),
        "pair3_s2": dict(b=dict(kind="pair", passes=3), c=dict(sweeps=2, alpha=1.5)),
        "pair3_sa": dict(b=dict(kind="pair", passes=3, smooth_P=True), c=dict(sweeps=1, alpha=1.0)),
        "pair2_sa": dict(b=dict(kind="pair", passes=2, smooth_P=True), c=dict(sweeps=1, alpha=1.0)),
        "mort_sa": dict(b=dict(kind="morton", smooth_P=True), c=dict(sweeps=1, alpha=1.0)),
    }
    for name in names or variants:
        v = variants[name]
        t0 = time.time()
        if v is None:
            lu = spl.splu(Avv.tocsc())
            def M(R):
                Z = R / d[:, None]; Z[:nvf] = lu.solve(R[:nvf]); return Z
            info = ""
        else:
            lv = build(Avv, vpts, **v["b"])
            M = pc3lib.make_M(S, lv, **v["c"])
            info = "levels %s nnz %s" % ([L.n for L in lv], [L.A.nnz for L in lv])
        tb = time.time() - t0
        X, it = pcg(Af, Ff, M, maxit=1500)
        print("%-12s iters %4d  (build %.0fs, solve %.0fs) %s" % (name, it, tb, time.time() - t0 - tb, info), flush=True)

if __name__ == "__main__":
    run(sys.argv[1], sys.argv[2:])


def handshake_agg(A, passes=3, rounds=6, theta=0.0):
    """What the GPU does (amg.cu): per pass, `rounds` rounds of 'every unmatched row points at its strongest unmatched
    neighbour (ties: smaller index); mutual pointers become a pair'; rows left over stay single.  Vectorised NumPy."""
    n = A.shape[0]
    total = np.arange(n)
    Ac = A.tocsr()
    for p in range(passes):
        m = Ac.shape[0]
        C = Ac.tocoo()
        d = Ac.diagonal()
        off = (C.row != C.col)
        r, c = C.row[off], C.col[off]
        w = -C.data[off] / np.sqrt(np.abs(d[r] * d[c]))
        keep = w > theta
        r, c, w = r[keep], c[keep], w[keep]
        match = np.full(m, -1, np.int64)
        for rd in range(rounds):
            ok = (match[r] < 0) & (match[c] < 0)
            rr, cc, ww = r[ok], c[ok], w[ok]
            if rr.size == 0: break
            # argmax per row: sort by (row, -w, col)
            o = np.lexsort((cc, -ww, rr))
            rr, cc = rr[o], cc[o]
            first = np.r_[True, rr[1:] != rr[:-1]]
            pick = np.full(m, -1, np.int64)
            pick[rr[first]] = cc[first]
            i = np.where(pick >= 0)[0]
            mutual = i[pick[pick[i]] == i]
            match[mutual] = pick[mutual]
        single = match < 0
        match[single] = np.where(single)[0]
        leader = np.arange(m) <= match
        ids = np.cumsum(leader) - 1
        ids = np.where(leader, ids, ids[match])
        na = int(leader.sum())
        T = sp.csr_matrix((np.ones(m), (np.arange(m), ids)), shape=(m, na))
        Ac = (T.T @ Ac @ T).tocsr()
        total = ids[total]
    return total, Ac.shape[0]
