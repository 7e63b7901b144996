"""Synthetic tetrahedral meshers (host side).

The reference meshes with Gmsh (3D half-ball, `gmsh_functions.py:544-684`) or Netgen (2D); mesh
generation stays on the host cores in this design (BASELINE.json north_star) and neither mesher is
installable in this image, so tests and benchmarks use the generators below.  They reproduce the
*properties the solve depends on* of the reference's 3D meshes:

  * half-ball y >= 0 of radius `domain_radius` around the batch centre, symmetry plane y = 0 with a
    natural boundary condition, outer sphere = 'dirichlet_boundary' made of the boundary faces whose
    three nodes lie on |x| = R (`gmsh_functions.py:581, 648-658`);
  * the electrode axis x = y = 0 is a chain of mesh edges and every electrode is a vertex
    (`gmsh_functions.py:552-575`);
  * element size graded from `h_electrode` at the electrodes to `h_max` at the boundary
    (`gmsh_functions.py:630-645`);
  * material index per tet, 0 = borehole mud, then per layer top->bottom flushed zone (if any) and
    undisturbed zone (`gmsh_functions.py:592-624`), here assigned from the tet centroid.

Point cloud = nested, slightly jittered body-centred-cubic lattices (level l has spacing
h_max / 2^l and is used where the size field asks for it) + exact axis / symmetry-plane / sphere
points; tets = scipy.spatial.Delaunay (Qhull); boundary slivers are peeled.
"""
from itertools import permutations

import numpy as np

# ----------------------------------------------------------------------------------------------
# structured box (unit tests, patch tests)
# ----------------------------------------------------------------------------------------------
_KUHN = []
for _perm in permutations(range(3)):
    _v = [0, 0, 0]
    _path = [0]
    for _ax in _perm:
        _v[_ax] = 1
        _path.append(_v[0] + 2 * _v[1] + 4 * _v[2])
    _KUHN.append(_path)


def box_mesh(n, lo=(0.0, 0.0, -1.0), hi=(1.0, 1.0, 1.0), dirichlet=lambda c: np.ones(c.shape[0], bool)):
    """Kuhn-split structured tet mesh of a box; the edge x=lo[0], y=lo[1] is a chain of mesh edges.

    `n` = cells per direction (int or 3-tuple).  Returns arrays (points, elems, bfacets, bc) with
    bc 1 = natural, 2 = 'dirichlet_boundary' where `dirichlet(facet centroids)` is true."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    xs = [np.linspace(lo[d], hi[d], m + 1) for d, m in enumerate((nx, ny, nz))]
    X, Y, Z = np.meshgrid(*xs, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def vid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    corner = np.stack([vid(I + (c & 1), J + ((c >> 1) & 1), K + ((c >> 2) & 1)) for c in range(8)], axis=1)
    elems = np.concatenate([corner[:, path] for path in _KUHN], axis=0).astype(np.int32)
    bfacets = boundary_facets(elems)
    cen = pts[bfacets].mean(axis=1)
    bc = np.where(dirichlet(cen), 2, 1).astype(np.int32)
    return pts, elems, bfacets, bc


def _facet_keys(facets, nv):
    """Order-independent int64 key of each facet (vertex numbers < 2^21 for triangles)."""
    f = np.sort(facets.astype(np.int64), axis=1)
    key = f[:, 0]
    for c in range(1, f.shape[1]):
        key = key * nv + f[:, c]
    return key


def boundary_facets(elems, return_owner=False):
    """Faces that belong to exactly one tet (triangles) / edges of exactly one triangle."""
    d1 = elems.shape[1]
    nv = int(elems.max()) + 1
    faces = np.concatenate([np.delete(elems, i, axis=1) for i in range(d1)], axis=0)
    if d1 == 4 and nv >= (1 << 21):
        # three vertex numbers no longer fit one int64 key: sort the sorted triples lexicographically instead
        f = np.sort(faces.astype(np.int64), axis=1)
        key2 = f[:, 0] * nv + f[:, 1]
        order = np.lexsort((f[:, 2], key2))
        k2, k3 = key2[order], f[order, 2]
        new = np.ones(order.shape[0], dtype=bool)
        new[1:] = (k2[1:] != k2[:-1]) | (k3[1:] != k3[:-1])
        starts = np.flatnonzero(new)
        cnt = np.diff(np.r_[starts, order.shape[0]])
        sel = np.sort(order[starts[cnt == 1]])
    else:
        _, idx, cnt = np.unique(_facet_keys(faces, nv), return_index=True, return_counts=True)
        sel = idx[cnt == 1]
    if return_owner:
        return faces[sel].astype(np.int32), sel % elems.shape[0]
    return faces[sel].astype(np.int32)


# ----------------------------------------------------------------------------------------------
# graded half-ball
# ----------------------------------------------------------------------------------------------
class SizeField:
    """h(x) = min( h_electrode + g*dist(electrodes),  h_axis + g*dist(tool segment),  h_max
                   [, h_borehole + g*max(0, rho - r_strip)  -- the borehole column resolved along the WHOLE axis, as a mesher
                      that fragments the domain by the borehole cylinder does (`gmsh_functions.py:576-591`)] )."""

    def __init__(self, electrodes_z, h_electrode, h_axis, h_max, grading, h_borehole=None, r_strip=0.0, g_borehole=0.6,
                 borehole_window=10.0, g_window=0.08):
        self.ez = np.asarray(sorted(electrodes_z), dtype=float)
        self.h_e, self.h_a, self.h_max, self.g = h_electrode, h_axis, h_max, grading
        # borehole term: h_b + g_b * max(0, rho - r_strip) + g_w * max(0, axial distance to the tool - window)
        self.h_b, self.r_strip, self.g_b, self.win, self.g_w = h_borehole, float(r_strip), float(g_borehole), float(borehole_window), float(g_window)
        self.z_lo, self.z_hi = self.ez[0], self.ez[-1]

    def __call__(self, p):
        p = np.atleast_2d(p)
        rho2 = p[:, 0] ** 2 + p[:, 1] ** 2
        z = p[:, 2]
        j = np.clip(np.searchsorted(self.ez, z), 1, self.ez.shape[0] - 1) if self.ez.shape[0] > 1 else np.zeros(z.shape, int)
        if self.ez.shape[0] > 1:
            dz = np.minimum(np.abs(z - self.ez[j - 1]), np.abs(z - self.ez[j]))
        else:
            dz = np.abs(z - self.ez[0])
        d_e = np.sqrt(rho2 + dz ** 2)
        zc = np.clip(z, self.z_lo, self.z_hi)
        d_a = np.sqrt(rho2 + (z - zc) ** 2)
        h = np.minimum(np.minimum(self.h_e + self.g * d_e, self.h_a + self.g * d_a), self.h_max)
        if self.h_b is not None:
            far = np.maximum(0.0, np.abs(z - zc) - self.win)
            h = np.minimum(h, self.h_b + self.g_b * np.maximum(0.0, np.sqrt(rho2) - self.r_strip) + self.g_w * far)
        return h

    def reach_b(self, s):
        """Radius around the axis inside which the borehole term asks for h < 2 s (<= 0: nowhere)."""
        return -1.0 if self.h_b is None else (2 * s - self.h_b) / self.g_b + self.r_strip

    def zreach_b(self, s):
        """Axial distance beyond the tool segment up to which the borehole term asks for h < 2 s."""
        return self.win + (2 * s - self.h_b) / self.g_w


def _axis_points(size, radius):
    """Graded 1-D subdivision of [-R, R] containing every electrode exactly."""
    must = np.unique(np.concatenate([[-radius, radius], size.ez]))
    zs = list(must)
    stack = list(zip(must[:-1], must[1:]))
    while stack:
        a, b = stack.pop()
        mid = 0.5 * (a + b)
        if (b - a) > size(np.array([[0.0, 0.0, mid]]))[0]:
            zs.append(mid)
            stack.append((a, mid))
            stack.append((mid, b))
    return np.unique(np.asarray(zs))


def _bcc_level_points(level, s, size, radius, rng, jitter, half):
    """New BCC lattice points of this level (spacing s) wherever the size field is finer than 2s."""
    u = s / 2.0  # integer unit
    # region where h < 2s  ->  inside a capsule around the tool segment (or everywhere at level 0)
    if level == 0:
        reach_e = reach_a = np.inf
    else:
        reach_e = (2 * s - size.h_e) / size.g
        reach_a = (2 * s - size.h_a) / size.g
    reach_b = size.reach_b(s) if level > 0 else -1.0
    if reach_e <= 0 and reach_a <= 0 and reach_b <= 0:
        return np.zeros((0, 3))
    reach = min(max(reach_e, reach_a, reach_b, 0.0) + 2 * s, radius)
    zlo = max(size.z_lo - reach, -radius)
    zhi = min(size.z_hi + reach, radius)
    if reach_b > 0:  # the borehole column is resolved well beyond the tool
        zlo, zhi = max(min(zlo, size.z_lo - size.zreach_b(s)), -radius), min(max(zhi, size.z_hi + size.zreach_b(s)), radius)
    nxy = int(np.ceil(reach / u))
    ix = np.arange(-nxy, nxy + 1)
    iy = np.arange(0 if half else -nxy, nxy + 1)
    iz = np.arange(int(np.floor(zlo / u)), int(np.ceil(zhi / u)) + 1)
    out = []
    for parity in (0, 1):
        ax, ay, az = ix[(ix & 1) == parity], iy[(iy & 1) == parity], iz[(iz & 1) == parity]
        if ax.size == 0 or ay.size == 0 or az.size == 0:
            continue
        # chunk along z to bound memory
        step = max(1, int(4e6 // max(1, ax.size * ay.size)))
        for k0 in range(0, az.size, step):
            A, B, C = np.meshgrid(ax, ay, az[k0:k0 + step], indexing="ij")
            A, B, C = A.ravel(), B.ravel(), C.ravel()
            if parity == 0 and level > 0:
                # drop points already present in the coarser lattice: all = 0 mod 4 or all = 2 mod 4
                m = (A & 3) | (B & 3) | (C & 3)
                m2 = ((A & 3) == 2) & ((B & 3) == 2) & ((C & 3) == 2)
                keep = ~((m == 0) | m2)
                A, B, C = A[keep], B[keep], C[keep]
            p = np.stack([A, B, C], axis=1) * u
            h = size(p)
            keep = h < 2 * s if level > 0 else np.ones(p.shape[0], bool)
            out.append(p[keep])
    if not out:
        return np.zeros((0, 3))
    p = np.concatenate(out)
    p = p + rng.uniform(-jitter, jitter, size=p.shape) * s
    return p


def _plane_level_points(level, s, size, radius, rng, jitter):
    """Nested centred-square lattice on the symmetry plane y = 0 (same nesting rule as the BCC lattice)."""
    u = s / 2.0
    if level == 0:
        reach = radius
    else:
        r = max((2 * s - size.h_e) / size.g, (2 * s - size.h_a) / size.g, size.reach_b(s))
        if r <= 0:
            return np.zeros((0, 3))
        reach = min(r + 2 * s, radius)
    zlo = max(size.z_lo - reach, -radius)
    zhi = min(size.z_hi + reach, radius)
    if level > 0 and size.reach_b(s) > 0:
        zlo, zhi = max(min(zlo, size.z_lo - size.zreach_b(s)), -radius), min(max(zhi, size.z_hi + size.zreach_b(s)), radius)
    nx = int(np.ceil(reach / u))
    ix = np.arange(-nx, nx + 1)
    iz = np.arange(int(np.floor(zlo / u)), int(np.ceil(zhi / u)) + 1)
    out = []
    for parity in (0, 1):
        A, C = np.meshgrid(ix[(ix & 1) == parity], iz[(iz & 1) == parity], indexing="ij")
        A, C = A.ravel(), C.ravel()
        if parity == 0 and level > 0:
            keep = ~((((A & 3) | (C & 3)) == 0) | (((A & 3) == 2) & ((C & 3) == 2)))
            A, C = A[keep], C[keep]
        p = np.stack([A * u, np.zeros(A.shape[0]), C * u], axis=1)
        if level > 0:
            p = p[size(p) < 2 * s]
        out.append(p)
    p = np.concatenate(out)
    jit = rng.uniform(-jitter, jitter, size=p.shape) * s
    jit[:, 1] = 0.0
    return p + jit


def _sphere_points(radius, h, half):
    """Fibonacci points on the (hemi)sphere + the rim circle on y = 0."""
    n = max(16, int(4 * np.pi * radius ** 2 / (0.866 * h * h)))
    k = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * k / n)
    theta = np.pi * (1 + 5 ** 0.5) * k
    p = radius * np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)
    if not half:
        return p
    p = p[p[:, 1] > 0.4 * h]
    m = max(8, int(2 * np.pi * radius / h))
    a = 2 * np.pi * (np.arange(m) + 0.5) / m
    rim = radius * np.stack([np.cos(a), np.zeros(m), np.sin(a)], axis=1)
    rim = rim[np.abs(rim[:, 0]) > 0.4 * h]  # the poles (0,0,+-R) come from the axis
    return np.concatenate([p, rim])


class Interfaces:
    """Material interfaces of a layered formation around a borehole, in the mesh frame (`gmsh_functions.py:576-624`: the
    reference fragments the half-ball by the borehole cylinder, the layer planes and the invasion cylinders, so its
    interfaces are mesh faces).  Here the point cloud is made to carry them: every lattice point closer than `frac` of the
    local size to an interface is PROJECTED onto it (radially onto a cylinder, vertically onto a dipping plane -- both keep
    a symmetry-plane point on y = 0), near-duplicates among the projected points are dropped, and the Delaunay tets then have
    the surfaces as (almost all) faces; the per-tet material still comes from the centroid.

      wall      : (z, r) arrays -- borehole radius along the axis (caliper), or a scalar
      tops      : ascending interface depths on the axis between consecutive layers
      dip_rad   : layer planes z = top_i + tan(dip) x
      invasion  : per layer (len(tops) + 1) flushed-zone radius or None"""

    def __init__(self, wall, tops, dip_rad=0.0, invasion=None):
        self.wall = wall
        self.tops = np.asarray(tops, dtype=float)
        self.tan = float(np.tan(dip_rad))
        self.cos = float(np.cos(dip_rad))
        nl = self.tops.shape[0] + 1
        inv = [None] * nl if invasion is None else list(invasion)
        self.inv = np.array([np.nan if (v is None or v != v) else float(v) for v in inv])

    def wall_radius(self, z):
        return np.full(z.shape, float(self.wall)) if np.isscalar(self.wall) else np.interp(z, self.wall[0], self.wall[1])

    def snap(self, pts, h, movable, frac=0.5):
        """-> (points, snapped mask).  `movable`: points that may move at all (not the axis, not the sphere)."""
        rho = np.hypot(pts[:, 0], pts[:, 1])
        z = pts[:, 2]
        best = np.full(pts.shape[0], np.inf)
        kind = np.zeros(pts.shape[0], np.int8)  # 1 = radial to radius `target`, 2 = vertical to depth `target`
        target = np.zeros(pts.shape[0])
        # borehole wall
        rb = self.wall_radius(z)
        d = np.abs(rho - rb)
        best, kind, target = d.copy(), np.ones(pts.shape[0], np.int8), rb.copy()
        # layer planes (only outside the borehole: inside it is all mud)
        zeff = z - self.tan * pts[:, 0]
        if self.tops.size:
            j = np.clip(np.searchsorted(self.tops, zeff), 0, self.tops.size)
            for cand in (np.clip(j - 1, 0, self.tops.size - 1), np.clip(j, 0, self.tops.size - 1)):
                dz = np.abs(zeff - self.tops[cand]) * self.cos
                use = (dz < best) & (rho > rb - frac * h)
                best = np.where(use, dz, best)
                kind = np.where(use, 2, kind).astype(np.int8)
                target = np.where(use, self.tops[cand] + self.tan * pts[:, 0], target)
        # invasion cylinders of the layer the point is in
        layer = np.searchsorted(self.tops, zeff) if self.tops.size else np.zeros(pts.shape[0], int)
        ri = self.inv[layer]
        di = np.abs(rho - ri)
        use = np.isfinite(ri) & (di < best)
        best = np.where(use, di, best)
        kind = np.where(use, 1, kind).astype(np.int8)
        target = np.where(use, ri, target)
        go = movable & (best < frac * h) & (rho > 1e-12)
        out = pts.copy()
        rad = go & (kind == 1)
        scale = np.where(rad, target / np.where(rho > 0, rho, 1.0), 1.0)
        out[:, 0] *= scale
        out[:, 1] *= scale
        ver = go & (kind == 2)
        out[ver, 2] = target[ver]
        return out, go


def half_ball_points(radius, electrodes_z, h_electrode=0.02, h_axis=0.1, h_max=None, grading=0.35, seed=0, half=True,
                     jitter=0.08, interfaces=None, snap_frac=0.5, h_borehole=None, r_strip=0.0, g_borehole=0.6, borehole_window=10.0,
                     g_window=0.08):
    """Graded point cloud; returns (points, n_axis) with the axis points first (sorted by z)."""
    rng = np.random.default_rng(seed)
    h_max = h_max or radius / 8.0
    size = SizeField(electrodes_z, h_electrode, h_axis, h_max, grading, h_borehole, r_strip, g_borehole, borehole_window, g_window)
    axis_z = _axis_points(size, radius)
    axis = np.stack([np.zeros_like(axis_z), np.zeros_like(axis_z), axis_z], axis=1)
    nlev = int(np.ceil(np.log2(h_max / min(h_electrode, h_axis)))) + 1
    vol, pla = [], []  # (2D mesher: same structure, meshgen2d.half_disc_mesh)
    for level in range(nlev):
        s = h_max / 2 ** level
        p = _bcc_level_points(level, s, size, radius, rng, jitter, half)
        if p.shape[0]:
            h = size(p)
            rho = np.hypot(p[:, 0], p[:, 1])
            keep = (np.linalg.norm(p, axis=1) < radius - 0.55 * h) & (rho > 0.6 * h)
            if half:
                keep &= p[:, 1] > 0.45 * h
            vol.append(p[keep])
        if half:
            q = _plane_level_points(level, s, size, radius, rng, jitter)
            if q.shape[0]:
                h = size(q)
                keep = (np.linalg.norm(q, axis=1) < radius - 0.55 * h) & (np.abs(q[:, 0]) > 0.6 * h)
                pla.append(q[keep])
    sph = _sphere_points(radius, h_max, half)
    # exactly cospherical points make one giant degenerate facet of the lifted hull (Qhull crawls):
    # pull them inside by a relative 1e-7 at random; "on the sphere" is tested with 1e-6 below
    sph *= 1.0 - 1e-7 * rng.uniform(0.0, 1.0, size=(sph.shape[0], 1))
    if interfaces is not None:
        # carry the material interfaces in the point cloud (class Interfaces): project, then thin out near-duplicates
        from scipy.spatial import cKDTree

        inner = np.concatenate(pla + vol) if (pla or vol) else np.zeros((0, 3))
        h = size(inner)
        inner, snapped = interfaces.snap(inner, h, np.ones(inner.shape[0], bool), snap_frac)
        # keep clear of the axis chain and of the sphere after the move
        ok = (np.hypot(inner[:, 0], inner[:, 1]) > 0.6 * h) | ~snapped
        ok &= (np.linalg.norm(inner, axis=1) < radius - 0.55 * h) | ~snapped
        inner, snapped, h = inner[ok], snapped[ok], h[ok]
        if snapped.any():
            tree = cKDTree(inner)
            dist, nb = tree.query(inner[snapped], k=2)
            idx = np.nonzero(snapped)[0]
            # a projected point that lands within 0.4 h of another point goes (of a pair of projected points the later one)
            close = dist[:, 1] < 0.4 * h[idx]
            other = nb[:, 1]
            drop = close & (~snapped[other] | (other < idx))
            keep = np.ones(inner.shape[0], bool)
            keep[idx[drop]] = False
            inner = inner[keep]
        parts = [axis, inner, sph]
        return np.concatenate(parts), axis.shape[0], size
    parts = [axis] + pla + vol + [sph]
    return np.concatenate(parts), axis.shape[0], size


def _quality(points, elems):
    """volume / (rms edge length)^3, normalised so the regular tet scores 1."""
    x = points[elems]
    vol = np.abs(np.linalg.det(x[:, 1:] - x[:, :1])) / 6.0
    e2 = sum(((x[:, i] - x[:, j]) ** 2).sum(axis=1) for i in range(4) for j in range(i + 1, 4)) / 6.0
    return vol / (e2 ** 1.5) * (6.0 * np.sqrt(2.0))


def _peel(pts, elems):
    """Remove boundary slivers (flat tets between nearly coplanar hull points)."""
    for _ in range(6):
        qual = _quality(pts, elems)
        _, owner = boundary_facets(elems, return_owner=True)
        drop = np.zeros(elems.shape[0], bool)
        drop[owner] = True  # tets that own a boundary face ...
        drop &= qual < 2e-2  # ... and are slivers
        if not drop.any():
            break
        elems = elems[~drop]
    return elems


def _plane_lift(pts, on_plane, radius, seed):
    """Qhull is ~10x slower with thousands of exactly coplanar hull points (the symmetry plane): it triangulates a copy
    whose plane points are lifted by <= 2e-11 R (the exact coordinates are kept in the mesh); the flat tets this creates
    in the plane are dropped.  Returns the lift of every point (0 off the plane)."""
    lift = np.zeros(pts.shape[0])
    lift[on_plane] = np.random.default_rng(seed).uniform(0.0, 2e-11 * radius, size=int(on_plane.sum()))
    return lift


def _delaunay(pts, lift, on_plane, ids=None):
    """Delaunay tets of the points `ids` (all by default), global vertex numbers, without the flat tets of the plane."""
    from scipy.spatial import Delaunay

    sub = pts if ids is None else pts[ids]
    lifted = sub.copy()
    lifted[:, 1] += lift if ids is None else lift[ids]
    simp = Delaunay(lifted).simplices
    elems = (simp if ids is None else ids[simp]).astype(np.int32)
    return elems[~on_plane[elems].all(axis=1)]


def _delaunay_peeled(pts, on_plane, radius, seed, lift=None):
    """Delaunay tets of the cloud with the flat tets of the symmetry plane and the boundary slivers removed."""
    lift = _plane_lift(pts, on_plane, radius, seed) if lift is None else lift
    return _peel(pts, _delaunay(pts, lift, on_plane))


def _retriangulate_around(pts, lift, on_plane, on_hull, elems, moved, rings=3):
    """Delaunay mesh after the vertices `moved` changed place, from the old mesh `elems`: only the tets around them are
    replaced by the Delaunay tets of the surrounding sub-cloud (`rings` vertex rings).  The Delaunay triangulation of
    points in general position is unique, so the patch fits the untouched rest exactly when the sub-cloud is wide enough;
    this is CHECKED (every face in at most two tets, every face of a single tet on the domain boundary) and None is
    returned when it does not hold, so the caller can triangulate the whole cloud instead."""
    nv = pts.shape[0]
    inner = np.zeros(nv, bool)
    inner[moved] = True
    inner[elems[inner[elems].any(axis=1)].ravel()] = True  # the moved vertices and their neighbours
    region = inner[elems].any(axis=1)                       # old tets to replace: everything touching them
    cloud = inner.copy()
    for _ in range(rings - 1):
        cloud[elems[cloud[elems].any(axis=1)].ravel()] = True
    ids = np.flatnonzero(cloud)
    if ids.size * 2 > nv:
        return None  # not local any more
    patch = _delaunay(pts, lift, on_plane, ids)
    patch = patch[inner[patch].any(axis=1)]
    new = np.concatenate([elems[~region], patch])
    # validity of the union, checked on the tets that touch the sub-cloud (a mismatch can only sit there): every face in at
    # most two tets; a face of a single tet is either a domain boundary face or on the outer rim of this sub-mesh, whose
    # faces have no vertex in the sub-cloud
    sub = new[cloud[new].any(axis=1)]
    faces = np.sort(np.concatenate([np.delete(sub, i, axis=1) for i in range(4)], axis=0).astype(np.int64), axis=1)
    order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    f = faces[order]
    first = np.ones(f.shape[0], bool)
    first[1:] = (f[1:] != f[:-1]).any(axis=1)
    starts = np.flatnonzero(first)
    cnt = np.diff(np.r_[starts, f.shape[0]])
    if cnt.max() > 2:
        return None
    single = f[starts[cnt == 1]]
    bnd = on_plane | on_hull
    if not (bnd[single].all(axis=1) | ~cloud[single].any(axis=1)).all():
        return None
    # boundary slivers can only have come back inside the patch
    if ((bnd[patch].sum(axis=1) >= 3) & (_quality(pts, patch) < 2e-2)).any():
        return _peel(pts, new)
    return new


def half_ball_mesh(radius, electrodes_z, material=None, **kw):
    """Graded half-ball (or full ball with half=False) tet mesh.

    Returns dict(points, elems, mat, bfacets, bc, bc_names, n_axis); vertices are renumbered along a
    Morton curve for memory locality with the axis vertices kept as mesh vertices.  `material(centroids)`
    -> 0-based material index per tet (default: all 0)."""
    half = kw.get("half", True)
    pts, n_axis, _ = half_ball_points(radius, electrodes_z, **{k: v for k, v in kw.items() if k not in ("improve", "improve_quality", "improve_mode", "improve_local")})
    # `interfaces=Interfaces(...)`: points projected onto the material interfaces (see class Interfaces)
    # Morton order with 21 bits per axis (locality of vertex numbers -> locality of CSR columns, at every
    # refinement level: the finest cells here are ~1e-4 of the domain)
    q = np.clip(((pts + radius) / (2 * radius) * (2 ** 21 - 1)).astype(np.uint64), 0, 2 ** 21 - 1)

    def spread(v):
        v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v

    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    pts = pts[np.argsort(code, kind="stable")]
    on_plane = pts[:, 1] == 0.0
    lift = _plane_lift(pts, on_plane, radius, kw.get("seed", 0) + 1)
    elems = _delaunay_peeled(pts, on_plane, radius, 0, lift=lift)
    # optional quality pass (off by default): Delaunay meshes of well-spaced points still hold SLIVERS (four nearly
    # coplanar, nearly cocircular points), and the slivers -- not the point spacing -- set the condition number the PCG
    # sees.  Vertices of the worst tets that are free to move (not on the axis, the symmetry plane or the sphere) are
    # perturbed by a fraction of the local edge length and the cloud is re-triangulated; the best mesh is kept.
    improve = int(kw.get("improve", 0))
    if improve > 0:
        rng = np.random.default_rng(kw.get("seed", 0) + 7)
        qmin = float(kw.get("improve_quality", 0.12))
        on_hull = np.linalg.norm(pts, axis=1) >= radius * (1 - 1e-6)
        fixed = on_plane | ((pts[:, 0] == 0.0) & (pts[:, 1] == 0.0)) | on_hull
        def rank_of(q):  # the PCG iteration count follows the WORST elements first, their number second
            return (int((q < 0.01).sum()), int((q < 0.05).sum()), int((q < qmin).sum()))

        best = (rank_of(_quality(pts, elems)), pts, elems)
        for it in range(improve):
            qual = _quality(pts, elems)
            # late rounds concentrate on the worst elements (few vertices move: the re-triangulation stays local)
            bad = np.where(qual < (qmin if it < 3 else 0.05))[0]
            if bad.size == 0:
                break
            x = pts[elems[bad]]
            h = np.sqrt(sum(((x[:, i] - x[:, j]) ** 2).sum(axis=1) for i in range(4) for j in range(i + 1, 4)) / 6.0)
            if kw.get("improve_mode", "normal") == "normal":
                # a sliver's four points are nearly coplanar: push ONE free vertex of it out of that plane
                c = x.mean(axis=1, keepdims=True)
                _, _, vt = np.linalg.svd(x - c)
                nrm = vt[:, 2, :]  # direction of least extent
                # prefer a fully free vertex; failing that one of the symmetry plane (it then moves inside the plane)
                semi = on_plane & ~on_hull & ~((pts[:, 0] == 0.0) & (pts[:, 1] == 0.0))
                score = (~fixed[elems[bad]]) * 2 + semi[elems[bad]] * 1
                has = score.max(axis=1) > 0
                pick = np.argmax(score, axis=1)
                tv = elems[bad][np.arange(bad.size), pick][has]
                d = ((x - c)[np.arange(bad.size), pick] * nrm).sum(axis=1)[has]
                sgn = np.where(d >= 0, 1.0, -1.0)
                move, first = np.unique(tv, return_index=True)
                if move.size == 0:
                    break
                dirn = sgn[first][:, None] * nrm[has][first]
                inpl = semi[move]
                if inpl.any():
                    dp = dirn[inpl].copy()
                    dp[:, 1] = 0.0
                    weak = np.linalg.norm(dp, axis=1) < 0.3  # the sliver lies (almost) in the plane: any in-plane direction
                    rnd = rng.standard_normal((int(weak.sum()), 3))
                    rnd[:, 1] = 0.0
                    dp[weak] = rnd
                    dirn[inpl] = dp / np.linalg.norm(dp, axis=1)[:, None]
                step = (0.3 * h[has][first])[:, None] * dirn
                ymin = np.where(inpl, 0.0, 0.35 * h[has][first])  # off-plane vertices keep their distance from the plane
            else:
                hv = np.full(pts.shape[0], np.inf)
                np.minimum.at(hv, elems[bad].ravel(), np.repeat(h, 4))
                move = np.where(np.isfinite(hv) & ~fixed)[0]
                if move.size == 0:
                    break
                step = rng.standard_normal((move.size, 3))
                step *= (0.18 * hv[move] / np.linalg.norm(step, axis=1))[:, None]
                ymin = 0.35 * hv[move]
            pts = pts.copy()
            pts[move] += step
            if half:
                # stay on this side of the symmetry plane, and not so close to it that a flat boundary tet appears
                pts[move, 1] = np.where(on_plane[move], 0.0, np.maximum(np.abs(pts[move, 1]), ymin))
            # around a few hundred moved vertices the patch fits almost always; with thousands it rarely does (measured)
            local = _retriangulate_around(pts, lift, on_plane, on_hull, elems, move) if (kw.get("improve_local", True) and move.size <= 300) else None
            elems = local if local is not None else _delaunay_peeled(pts, on_plane, radius, 0, lift=lift)
            sc = rank_of(_quality(pts, elems))
            if sc < best[0]:
                best = (sc, pts, elems)
        _, pts, elems = best
    # positive orientation
    x = pts[elems]
    neg = np.linalg.det(x[:, 1:] - x[:, :1]) < 0
    elems[neg] = elems[neg][:, [0, 2, 1, 3]]
    # drop unused vertices (none expected) and build boundary
    used = np.zeros(pts.shape[0], bool)
    used[elems.ravel()] = True
    if not used.all():
        remap = np.cumsum(used) - 1
        pts = pts[used]
        elems = remap[elems].astype(np.int32)
    bfacets = boundary_facets(elems)
    rr = np.linalg.norm(pts, axis=1)
    on_sphere = rr >= radius * (1 - 1e-6)
    bc = np.where(on_sphere[bfacets].all(axis=1), 2, 1).astype(np.int32)
    cen = pts[elems].mean(axis=1)
    mat = np.zeros(elems.shape[0], np.int32) if material is None else np.asarray(material(cen), dtype=np.int32)
    return {"points": pts, "elems": elems, "mat": mat, "bfacets": bfacets, "bc": bc,
            "bc_names": ["symmetry_plane", "dirichlet_boundary"], "half": half}


# ----------------------------------------------------------------------------------------------
# material predicates
# ----------------------------------------------------------------------------------------------
def layered_material(layer_tops, dip_rad=0.0, borehole_radius=0.1, invasion=None, inclusion=None):
    """Centroid -> material index following the reference's order (`gmsh_functions.py:592-624`):
    0 borehole; then per layer (top->bottom): flushed zone if the layer has one, undisturbed zone.
    `layer_tops`: ascending interface depths z_1..z_{n-1} on the axis (n layers);
    `invasion`: per-layer flushed-zone radius or None/NaN; dipping planes z = z_i + tan(dip) x;
    `inclusion`: (centre xyz, radius) gets its own material index appended at the end."""
    tops = np.asarray(layer_tops, dtype=float)
    nl = tops.shape[0] + 1
    inv = [None] * nl if invasion is None else list(invasion)
    ids, nxt = [], 1
    for i in range(nl):
        has = inv[i] is not None and inv[i] == inv[i]
        fz = nxt if has else -1
        nxt += 1 if has else 0
        ids.append((fz, nxt))
        nxt += 1
    n_materials = nxt + (1 if inclusion is not None else 0)

    def fn(c):
        rho = np.hypot(c[:, 0], c[:, 1])
        # borehole radius: constant, or a caliper profile (z, r) interpolated along the axis
        rb = borehole_radius if np.isscalar(borehole_radius) else np.interp(c[:, 2], borehole_radius[0], borehole_radius[1])
        zeff = c[:, 2] - np.tan(dip_rad) * c[:, 0]
        layer = np.searchsorted(tops, zeff)
        m = np.empty(c.shape[0], np.int32)
        for i, (fz, uz) in enumerate(ids):
            sel = layer == i
            if fz >= 0:
                m[sel] = np.where(rho[sel] < inv[i], fz, uz)
            else:
                m[sel] = uz
        if inclusion is not None:
            cen, rad = inclusion
            m[np.linalg.norm(c - np.asarray(cen), axis=1) < rad] = nxt
        m[rho < rb] = 0
        return m

    fn.n_materials = n_materials
    return fn
