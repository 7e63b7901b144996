"""Synthetic 2D axisymmetric mesher (host side): half-disc in (r, z), conforming to the model interfaces.

Stands in for `ConstructNetgen2dModel` / `ConstructGmsh2dModel` (`/root/reference/remo3d/netgen_functions.py:120-334`,
`gmsh_functions.py:384-542`), which need Netgen / Gmsh.  Preserved contract:

  * coordinates (x, y) = (r, z), axis r = 0 is a chain of mesh edges carrying every electrode as a vertex
    (`netgen_functions.py:135-138, 226-229`), natural condition on the axis;
  * outer arc r^2 + z^2 = R^2 is the Dirichlet boundary: bc number 2 (`netgen_functions.py:310`, `worker.py:97`)
    named 'dirichlet_boundary' (`gmsh_functions.py:507-519`); the axis is bc 1;
  * material 0 = borehole, then per layer top -> bottom flushed zone (if any), undisturbed zone
    (`netgen_functions.py:99-116`); the borehole wall (caliper polyline), the bed boundaries and the invasion fronts
    are mesh edges, so every triangle lies in one material.

Point cloud = nested jittered centred-square lattices driven by a size field + points marched along every interface;
lattice points too close to an interface are dropped, which makes every interface segment a Delaunay (Gabriel) edge.
"""
import numpy as np

from .meshgen import boundary_facets


class SizeField2D:
    """h = min(h_e + g dist(electrodes), h_a + g dist(tool segment), h_b + g max(0, r - r_strip), h_max)."""

    def __init__(self, electrodes_z, h_electrode, h_axis, h_borehole, r_strip, h_max, grading):
        self.ez = np.asarray(sorted(electrodes_z), dtype=float)
        self.h_e, self.h_a, self.h_b, self.r_strip, self.h_max, self.g = h_electrode, h_axis, h_borehole, r_strip, h_max, grading
        self._ezl = [float(v) for v in self.ez]

    def __call__(self, p):
        p = np.atleast_2d(p)
        r, z = p[:, 0], p[:, 1]
        dz = np.min(np.abs(z[:, None] - self.ez[None, :]), axis=1)
        d_e = np.hypot(r, dz)
        zc = np.clip(z, self.ez[0], self.ez[-1])
        d_a = np.hypot(r, z - zc)
        h = np.minimum(self.h_e + self.g * d_e, self.h_a + self.g * d_a)
        h = np.minimum(h, self.h_b + self.g * np.maximum(0.0, r - self.r_strip))
        return np.minimum(h, self.h_max)

    def scalar(self, r, z):
        """The same size at ONE point in plain Python: the bisections of _march evaluate thousands of single points, where the
        NumPy call overhead of __call__ (not the arithmetic) was a quarter of the mesh time."""
        import bisect
        from math import hypot

        ez = self._ezl
        j = bisect.bisect_left(ez, z)
        dz = min(abs(z - ez[j - 1]) if j > 0 else float("inf"), abs(ez[j] - z) if j < len(ez) else float("inf"))
        zc = min(max(z, ez[0]), ez[-1])
        h = min(self.h_e + self.g * hypot(r, dz), self.h_a + self.g * hypot(r, z - zc), self.h_b + self.g * max(0.0, r - self.r_strip))
        return min(h, self.h_max)


def _march(p0, p1, size, must=()):
    """Points along the segment p0 -> p1 with spacing <= size(x), including both ends and the `must` parameters."""
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    L = np.linalg.norm(p1 - p0)
    if L == 0:
        return p0[None, :]
    ts = sorted(set([0.0, 1.0] + [float(t) for t in must if 0 < t < 1]))
    out = []
    stack = list(zip(ts[:-1], ts[1:]))
    keep = set(ts)
    while stack:
        a, b = stack.pop()
        mid = 0.5 * (a + b)
        if (b - a) * L > 0.9 * size.scalar(p0[0] + mid * (p1[0] - p0[0]), p0[1] + mid * (p1[1] - p0[1])):
            keep.add(mid)
            stack += [(a, mid), (mid, b)]
    for t in sorted(keep):
        out.append(p0 + t * (p1 - p0))
    return np.array(out)


def _seg_dist(p, a, b):
    """Distance of points p (n x 2) to the segment a-b."""
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(ab @ ab, 1e-300), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[:, None] * ab), axis=1)


def half_disc_mesh(radius, electrodes_z, wall, layer_tops, invasion, h_electrode=0.005, h_axis=0.03, h_borehole=0.05,
                   h_max=None, grading=0.3, seed=0, jitter=0.1):
    """Conforming triangle mesh of the half-disc.

    wall      : (z, r) arrays, the borehole wall polyline relative to the mesh centre (extended to +-R at constant radius)
    layer_tops: ascending interface depths between consecutive layers (relative)
    invasion  : per layer (len(layer_tops)+1) flushed-zone radius or None
    Returns dict(points (r,z), elems, mat, bfacets, bc, bc_names)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    R = float(radius)
    # far-field size: with R / 8 (6 m at R = 50) the long-spacing tools of the reference's examples are 2 % off (7 % at R = 25)
    # whatever the near-field sizes; from R / 20 on the logs are converged to < 1e-3 (profiles/r02_notes.md, reference-log study)
    h_max = h_max or min(R / 25.0, 2.0)
    wz, wr = np.asarray(wall[0], float), np.asarray(wall[1], float)
    sel = (wz > -R) & (wz < R)
    wz = np.concatenate([[-R], wz[sel], [R]])
    wr = np.concatenate([[np.interp(-R, wall[0], wall[1])], wr[sel], [np.interp(R, wall[0], wall[1])]])
    rwall = lambda z: np.interp(z, wz, wr)
    tops = np.asarray(layer_tops, float)
    tops = tops[(tops > -R) & (tops < R)]
    nl_all = len(layer_tops) + 1
    inv = [None if (v is None or v != v) else float(v) for v in invasion]
    r_strip = max([wr.max()] + [v for v in inv if v is not None])
    size = SizeField2D(electrodes_z, h_electrode, h_axis, h_borehole, r_strip, h_max, grading)

    # ---- interface segments (each a list of marched points)
    segs = []

    def add(p0, p1, must=()):
        pts = _march(p0, p1, size, must)
        segs.append(pts)

    # axis, with electrodes
    az = np.unique(np.concatenate([[-R, R], size.ez]))
    for a, b in zip(az[:-1], az[1:]):
        add((0.0, a), (0.0, b))
    # borehole wall polyline, clipped to the disc; break points also at the bed boundaries
    zb = np.unique(np.concatenate([wz, tops]))
    zb = zb[(np.hypot(rwall(zb), zb) < R)]
    for a, b in zip(zb[:-1], zb[1:]):
        add((rwall(a), a), (rwall(b), b))
    # wall end points -> extend to the arc along the wall direction (constant radius): close the borehole region
    for zend, sgn in ((zb[0], -1.0), (zb[-1], 1.0)):
        r0 = rwall(zend)
        zarc = sgn * np.sqrt(max(R * R - r0 * r0, 0.0))
        add((r0, zend), (r0, zarc))
    # bed boundaries: from the wall to the arc, with a break at the invasion fronts of the two adjacent layers
    layer_index_of_top = {float(t): i for i, t in enumerate(np.asarray(layer_tops, float))}
    for t in tops:
        i = layer_index_of_top[float(t)]  # interface between layer i (above) and i+1 (below)
        r0, r1 = rwall(t), np.sqrt(R * R - t * t)
        must = [(v - r0) / (r1 - r0) for v in (inv[i], inv[i + 1]) if v is not None and r0 < v < r1]
        add((r0, t), (r1, t), must)
    # invasion fronts: vertical segments r = r_fz between the layer's boundaries (clipped to the disc)
    bounds = np.concatenate([[-np.inf], np.asarray(layer_tops, float), [np.inf]])
    for i in range(nl_all):
        if inv[i] is None:
            continue
        zmax = np.sqrt(max(R * R - inv[i] ** 2, 0.0))
        a, b = max(bounds[i], -zmax), min(bounds[i + 1], zmax)
        if b > a:
            add((inv[i], a), (inv[i], b))
    # arc
    m = max(16, int(np.pi * R / h_max))
    ang = np.linspace(-np.pi / 2, np.pi / 2, m + 1)
    arc = R * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    arc[:, 0] = np.maximum(arc[:, 0], 0.0)
    feat = np.concatenate(segs)
    # snap arc points near feature end points on the arc
    ends_on_arc = feat[np.abs(np.hypot(feat[:, 0], feat[:, 1]) - R) < 1e-9]
    arc = np.concatenate([arc, ends_on_arc])
    ang_all = np.arctan2(arc[:, 1], arc[:, 0])
    arc = arc[np.argsort(ang_all)]
    keep = np.ones(arc.shape[0], bool)
    for i in range(1, arc.shape[0]):
        if np.linalg.norm(arc[i] - arc[i - 1]) < 0.3 * h_max and keep[i - 1]:
            # drop the generic arc point, keep feature end points
            is_end_i = np.any(np.all(np.abs(ends_on_arc - arc[i]) < 1e-9, axis=1)) if ends_on_arc.size else False
            if is_end_i:
                is_end_prev = np.any(np.all(np.abs(ends_on_arc - arc[i - 1]) < 1e-9, axis=1))
                if not is_end_prev:
                    keep[i - 1] = False
            else:
                keep[i] = False
    arc = arc[keep]

    # ---- lattice points by level
    nlev = int(np.ceil(np.log2(h_max / min(h_electrode, h_axis, h_borehole)))) + 1
    cloud = []
    for level in range(nlev):
        s = h_max / 2 ** level
        u = s / 2.0
        if level == 0:
            rmax, zlo, zhi = R, -R, R
        else:
            # region where h < 2s: near the tool / electrodes, or within the borehole strip
            reach = max((2 * s - size.h_e) / size.g, (2 * s - size.h_a) / size.g, 0.0) + 2 * s
            reach_b = max((2 * s - size.h_b) / size.g, 0.0) + r_strip + 2 * s if 2 * s > size.h_b else 0.0
            rmax = min(R, max(reach, reach_b))
            if reach_b > 0:
                zlo, zhi = -R, R
            else:
                zlo, zhi = max(-R, size.ez[0] - reach), min(R, size.ez[-1] + reach)
            if rmax <= 0:
                continue
        ir = np.arange(0, int(np.ceil(rmax / u)) + 1)
        iz = np.arange(int(np.floor(zlo / u)), int(np.ceil(zhi / u)) + 1)
        for parity in (0, 1):
            A, C = np.meshgrid(ir[(ir & 1) == parity], iz[(iz & 1) == parity], indexing="ij")
            A, C = A.ravel(), C.ravel()
            if parity == 0 and level > 0:
                k = ~((((A & 3) | (C & 3)) == 0) | (((A & 3) == 2) & ((C & 3) == 2)))
                A, C = A[k], C[k]
            p = np.stack([A * u, C * u], axis=1)
            if level > 0 and p.shape[0]:
                p = p[size(p) < 2 * s]
            if p.shape[0]:
                p = p + rng.uniform(-jitter, jitter, size=p.shape) * s
                cloud.append(p)
    cloud = np.concatenate(cloud)
    h = size(cloud)
    ok = (cloud[:, 0] > 0.55 * h) & (np.hypot(cloud[:, 0], cloud[:, 1]) < R - 0.55 * h)
    cloud, h = cloud[ok], h[ok]
    # drop lattice points within 0.55 h of any interface segment (keeps every interface edge Gabriel)
    ok = np.ones(cloud.shape[0], bool)
    from scipy.spatial import cKDTree

    tree = cKDTree(cloud)
    c055 = 0.55 / (1.0 - 0.55 * size.g)
    for pts in segs[len(az) - 1:]:  # all but the axis pieces (already excluded by r > 0.55 h)
        a, b = pts[0], pts[-1]
        # a point at distance d is dropped if d < 0.55 h(point) <= 0.55 (h_seg + g d) (h is Lipschitz with the grading g):
        # only points within 0.55 h_seg / (1 - 0.55 g) of the segment can qualify
        half = 0.5 * float(np.hypot(*(b - a)))
        h_seg = max(size.scalar(a[0], a[1]), size.scalar(b[0], b[1])) + size.g * half
        near = np.asarray(tree.query_ball_point(0.5 * (a + b), half + c055 * h_seg + 1e-12), dtype=np.int64)
        near = near[ok[near]] if near.size else near
        if near.size:
            d = _seg_dist(cloud[near], a, b)
            ok[near[d < 0.55 * h[near]]] = False
    cloud = cloud[ok]
    # arc points: pull inside by <= 1e-7 R (cocircular points are degenerate for Qhull), except feature end points
    arc_in = arc * (1.0 - 1e-7 * rng.uniform(0.2, 1.0, size=(arc.shape[0], 1)))
    pts = np.concatenate([feat, arc_in, cloud])
    # merge duplicates (interface crossing points appear in several segments)
    key = np.round(pts / 1e-9).astype(np.int64)
    _, first = np.unique(key, axis=0, return_index=True)
    pts = pts[np.sort(first)]
    # Morton order for locality
    q = np.clip(((pts + [0.0, R]) / (2 * R) * (2 ** 21 - 1)).astype(np.uint64), 0, 2 ** 21 - 1)

    def spread(v):
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v

    pts = pts[np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)), kind="stable")]
    lifted = pts.copy()
    on_axis = lifted[:, 0] == 0.0
    lifted[on_axis, 0] -= rng.uniform(0.0, 2e-11 * R, size=int(on_axis.sum()))  # Qhull: avoid many collinear hull points
    tri = Delaunay(lifted)
    elems = tri.simplices.astype(np.int32)
    x = pts[elems]
    area2 = (x[:, 1, 0] - x[:, 0, 0]) * (x[:, 2, 1] - x[:, 0, 1]) - (x[:, 1, 1] - x[:, 0, 1]) * (x[:, 2, 0] - x[:, 0, 0])
    e2 = ((x[:, 1] - x[:, 0]) ** 2).sum(1) + ((x[:, 2] - x[:, 1]) ** 2).sum(1) + ((x[:, 0] - x[:, 2]) ** 2).sum(1)
    keep = np.abs(area2) > 1e-6 * e2  # flat triangles on the hull (collinear axis points, arc slivers)
    elems, area2 = elems[keep], area2[keep]
    neg = area2 < 0
    elems[neg] = elems[neg][:, [0, 2, 1]]
    used = np.zeros(pts.shape[0], bool)
    used[elems.ravel()] = True
    if not used.all():
        remap = np.cumsum(used) - 1
        pts, elems = pts[used], remap[elems].astype(np.int32)
    # materials from centroids (exact: interfaces are mesh edges)
    cen = pts[elems].mean(axis=1)
    ids, nxt = [], 1
    for i in range(nl_all):
        fz = nxt if inv[i] is not None else -1
        nxt += 1 if inv[i] is not None else 0
        ids.append((fz, nxt))
        nxt += 1
    layer = np.searchsorted(np.asarray(layer_tops, float), cen[:, 1])
    mat = np.empty(elems.shape[0], np.int32)
    for i, (fz, uz) in enumerate(ids):
        s_ = layer == i
        mat[s_] = np.where(cen[s_, 0] < inv[i], fz, uz) if fz >= 0 else uz
    mat[cen[:, 0] < rwall(cen[:, 1])] = 0
    bf = boundary_facets(elems)
    mid = pts[bf].mean(axis=1)
    on_arc = np.hypot(mid[:, 0], mid[:, 1]) > R * (1 - 1e-3)
    bc = np.where(on_arc & (mid[:, 0] > 1e-9 * R), 2, 1).astype(np.int32)
    return {"points": pts, "elems": elems, "mat": mat, "bfacets": bf, "bc": bc, "bc_names": ["axis", "dirichlet_boundary"],
            "interfaces": segs[len(az) - 1:], "n_materials": nxt}


def interfaces_are_edges(mesh_dict):
    """Fraction of interface sub-segments that are mesh edges (1.0 = fully conforming)."""
    pts, elems = mesh_dict["points"], mesh_dict["elems"]
    key = {tuple(np.round(p / 1e-9).astype(np.int64)): i for i, p in enumerate(pts)}
    edges = set()
    for a, b in ((0, 1), (1, 2), (0, 2)):
        for u, v in zip(elems[:, a], elems[:, b]):
            edges.add((min(u, v), max(u, v)))
    tot = hit = 0
    for seg in mesh_dict["interfaces"]:
        ids = [key.get(tuple(np.round(p / 1e-9).astype(np.int64)), -1) for p in seg]
        for u, v in zip(ids[:-1], ids[1:]):
            tot += 1
            hit += (min(u, v), max(u, v)) in edges
    return hit / max(tot, 1)
