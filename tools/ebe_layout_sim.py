"""CPU model of the shared-memory traffic of k_spmm_ebe (csrc/ebe.cu): rebuilds the per-batch tables (distinct dofs, pieces
of <= 8 entries, ranks by entry count, jagged-diagonal scratch positions) in NumPy exactly as k_ebe_batch does and counts
the wavefronts of the two random 8-byte access streams of a pass -- the gathers of staged rows (10 LDS.64 per tet) and
the scatters of element results (10 STS.64 per tet) -- under the half-warp bank model (16 lanes x 8 bytes: one
wavefront per distinct address landing in the busiest 8-byte bank pair).  Used to choose the conflict-aware layout
(`--layout color`) before spending GPU time; numbers in profiles/r02_notes.md.

    python tools/ebe_layout_sim.py 200k [nbatches]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import fem_oracle as fo  # noqa: E402

TPB, NLD, SPLIT = 256, 10, 8


def morton3(p):
    lo, hi = p.min(0), p.max(0)
    q = np.clip((p - lo) / (hi - lo) * 2097151.0, 0, 2097151).astype(np.uint64)

    def spread(v):
        v = v & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v

    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))


def batch_entries(dofs):
    """dofs: (ntet<=256, 10) -> sorted entry list as k_ebe_batch sees it: arrays (dof, tet, slot) sorted by dof, stable in
    (tet, slot) order."""
    nt = dofs.shape[0]
    tet = np.repeat(np.arange(nt), NLD)
    slot = np.tile(np.arange(NLD), nt)
    d = dofs.reshape(-1)
    o = np.argsort(d, kind="stable")
    return d[o], tet[o], slot[o]


def pieces_of(d):
    """piece id of every sorted entry (a dof with more than SPLIT entries is cut into pieces of SPLIT) + piece sizes."""
    n = d.shape[0]
    first = np.r_[True, d[1:] != d[:-1]]
    start = np.maximum.accumulate(np.where(first, np.arange(n), 0))
    within = np.arange(n) - start
    head = first | (within % SPLIT == 0)
    pid = np.cumsum(head) - 1
    cnt = np.bincount(pid)
    idx_in_piece = within % SPLIT
    return pid, cnt, idx_in_piece


def wavefronts(addr_word, half=16):
    """addr_word: (nwarp_groups, 16) word (8-byte) addresses of one half-warp request each -> wavefronts per request."""
    out = np.empty(addr_word.shape[0], np.int64)
    for i, a in enumerate(addr_word):
        u = np.unique(a)
        out[i] = np.bincount(u % half, minlength=half).max()
    return out


def layout_current(d, tet, slot, pid, cnt, iip):
    """ranks by entry count (descending, stable in dof order), jd offsets; returns (row of every piece, lpos of every entry)."""
    order = np.argsort(-cnt, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    scount = cnt[order]
    ng = np.array([(scount > i).sum() for i in range(SPLIT + 1)])
    jd = np.r_[0, np.cumsum(ng)][: SPLIT + 1]
    return rank, jd[iip] + rank[pid], jd


def layout_color(d, tet, slot, pid, cnt, iip, xst, seed=0):
    """Conflict-aware layout.  (1) rows: inside every entry-count class the pieces may take any row of the class; a row's
    bank pair is (row * xst) mod 16 for every pass (+ r, a common shift), so rows are dealt greedily: pieces in descending
    number of gather groups pick the residue that collides least with the pieces already placed in their groups (a group =
    the 16 lanes of one half-warp reading one slot), subject to the residue capacity of the class.  (2) entry order inside a
    piece: entry i of a piece of n entries lands at jd[i] + row; the n entries may take the n indices in any order, chosen
    greedily against the scatter groups (same half-warp, same slot, all addresses distinct)."""
    npieces = cnt.shape[0]
    order = np.argsort(-cnt, kind="stable")
    scount = cnt[order]
    ng = np.array([(scount > i).sum() for i in range(SPLIT + 1)])
    jd = np.r_[0, np.cumsum(ng)][: SPLIT + 1]
    group = (tet // 16) * NLD + slot  # gather / scatter group of every entry
    ngroups = group.max() + 1
    # entries of each piece
    ent_of = [[] for _ in range(npieces)]
    for e, p in enumerate(pid):
        ent_of[p].append(e)
    occ = np.zeros((ngroups, 16), np.int32)  # gather: distinct rows per (group, residue)
    row = np.full(npieces, -1, np.int64)
    # classes = contiguous rank ranges of equal count
    pos = 0
    for c in range(SPLIT, 0, -1):
        members = np.nonzero(cnt == c)[0]
        if members.size == 0:
            continue
        rows = np.arange(pos, pos + members.size)
        pos += members.size
        res = (rows * xst) % 16
        free = {r: list(rows[res == r]) for r in range(16)}
        # most constrained first: pieces with many distinct groups
        key = [-len(set(group[ent_of[p]])) for p in members]
        for p in members[np.argsort(key, kind="stable")]:
            gs = np.unique(group[ent_of[p]])
            cost = occ[gs].sum(axis=0).astype(float)  # rows already on that residue in my groups
            cost = np.array([cost[r] if free[r] else np.inf for r in range(16)])
            r = int(np.argmin(cost))
            row[p] = free[r].pop()
            occ[gs, r] += 1
    # scatter: entry order inside the pieces
    socc = np.zeros((ngroups, 16), np.int32)
    lpos = np.empty(d.shape[0], np.int64)
    for p in np.argsort(-cnt, kind="stable"):
        ents = ent_of[p]
        n = len(ents)
        avail = list(range(n))
        for e in ents:
            g = group[e]
            best = min(avail, key=lambda i: socc[g, (jd[i] + row[p]) % 16])
            avail.remove(best)
            lpos[e] = jd[best] + row[p]
            socc[g, lpos[e] % 16] += 1
    return row, lpos, jd


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "200k"
    nb_max = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    nr = 5
    xst = nr | 1
    task, flat = bench.make_task()
    m = bench.make_mesh(size, task, print)
    pts, elems = m["points"], m["elems"]
    space = fo.Space(pts.shape[0], elems, 2, 3)
    dofs = space.elem_dofs()
    cen = pts[space.sorted_elems].mean(axis=1)
    perm = np.argsort(morton3(cen), kind="stable")
    dofs = dofs[perm]
    nb = (dofs.shape[0] + TPB - 1) // TPB
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(nb - 1, size=min(nb_max, nb - 1), replace=False))
    tot = {"cur": [0, 0, 0], "color": [0, 0, 0]}
    t0 = time.time()
    U = []
    for b in pick:
        bd = dofs[b * TPB:(b + 1) * TPB]
        d, tet, slot = batch_entries(bd)
        pid, cnt, iip = pieces_of(d)
        U.append(cnt.shape[0])
        for name in ("cur", "color"):
            if name == "cur":
                row, lpos, jd = layout_current(d, tet, slot, pid, cnt, iip)
                row = row[pid]
            else:
                rowp, lpos, jd = layout_color(d, tet, slot, pid, cnt, iip, xst)
                row = rowp[pid]
            # requests: per (half-warp, slot): 16 lanes
            hw = tet // 16
            g = hw * NLD + slot
            o = np.lexsort((tet, g))
            gat = (row[o] * xst).reshape(-1, 16)
            sca = lpos[o].reshape(-1, 16)
            wg = wavefronts(gat).sum()
            ws = wavefronts(sca).sum()
            tot[name][0] += wg
            tot[name][1] += ws
            tot[name][2] += gat.shape[0]
    for name, (wg, ws, n) in tot.items():
        print("%-6s gathers %.2f wavefronts per half-warp request, scatters %.2f  (%d requests, %d batches, U mean %.0f max %d)  [%.0fs]"
              % (name, wg / n, ws / n, n, len(pick), np.mean(U), np.max(U), time.time() - t0))


if __name__ == "__main__":
    main()
