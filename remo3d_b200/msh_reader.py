"""Gmsh MSH 2.2 ASCII reader / writer (host side), numbering-compatible with the reference's `ReadGmsh`.

Restates `/root/reference/remo3d/gmsh_functions.py:177-382` (itself derived from Netgen's read_gmsh.py) for the element
types this path uses (points 15, lines 1, triangles 2, tets 4):

  * vertices are numbered in the order of the `$Nodes` section (`:251-257`), whatever their node ids;
  * elements keep the file order per dimension (`:259-381`), node order unchanged for first-order simplices;
  * material index = order of first appearance of the ELEMENTARY tag (`tags[1]`) among the volume elements (3D) or the
    surface elements (2D) (`:332-340, 353-361`); boundary-condition index likewise among the surface (3D) / line (2D)
    elements, with `bcname = PhysicalNames[tags[0]]` (`:292-303, 319-330`).  Materials are 0-based here (index into the
    sigma list), bc numbers 1-based as in Netgen.

The reference parses line by line in Python (a real bottleneck beyond ~1 M tets); here the two big sections are
tokenised with one `np.fromstring` each and the element rows are recovered block-wise (Gmsh writes elements grouped by
type, so rows of equal length are contiguous).
"""
import numpy as np

from .mesh import Mesh

_NODES = {15: 1, 1: 2, 2: 3, 4: 4}
_DIM = {15: 0, 1: 1, 2: 2, 4: 3}


def _section(text, name):
    a = text.find("$" + name)
    if a < 0:
        return None
    a = text.index("\n", a) + 1
    b = text.index("$End" + name, a)
    return text[a:b]


def _first_appearance_index(tags):
    """Order-of-first-appearance numbering (0-based) of an int array."""
    uniq, first, inv = np.unique(tags, return_index=True, return_inverse=True)
    rank = np.empty(uniq.shape[0], np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(uniq.shape[0])
    return rank[inv.reshape(-1)], uniq[np.argsort(first, kind="stable")]


def parse_msh(text, dim):
    """-> dict(points nv x 3, per-dimension element blocks, names).  Raises ValueError on unsupported content."""
    fmt = _section(text, "MeshFormat")
    if fmt is None or not fmt.split()[0].startswith("2"):
        raise ValueError("only Gmsh MSH 2.x ASCII files are supported")
    names = {0: "default"}
    pn = _section(text, "PhysicalNames")
    if pn is not None:
        for line in pn.strip().split("\n")[1:]:
            f = line.split()
            names[int(f[1])] = f[2][1:-1]  # gmsh_functions.py:249 (names with blanks are cut at the first blank)
    nodes = _section(text, "Nodes")
    nl = nodes.index("\n")
    nn = int(nodes[:nl].split()[0])
    flat = np.fromstring(nodes[nl + 1:], sep=" ")
    if flat.shape[0] != 4 * nn:
        raise ValueError("malformed $Nodes section")
    flat = flat.reshape(nn, 4)
    node_ids = flat[:, 0].astype(np.int64)
    points = np.ascontiguousarray(flat[:, 1:4])
    lut = np.full(int(node_ids.max()) + 1, -1, np.int64)
    lut[node_ids] = np.arange(nn)

    elems = _section(text, "Elements")
    nl = elems.index("\n")
    ne = int(elems[:nl].split()[0])
    flat = np.fromstring(elems[nl + 1:], dtype=np.int64, sep=" ")
    blocks = {0: [], 1: [], 2: [], 3: []}
    pos, count = 0, 0
    while pos < flat.shape[0]:
        etype, ntags = int(flat[pos + 1]), int(flat[pos + 2])
        if etype not in _NODES:
            raise ValueError("element type %d not supported on this path (first-order points/lines/triangles/tets only)" % etype)
        if ntags < 2:
            raise ValueError("elements need at least two tags (physical, elementary)")
        L = 3 + ntags + _NODES[etype]
        rows = flat[pos: pos + ((flat.shape[0] - pos) // L) * L].reshape(-1, L)
        same = (rows[:, 1] == etype) & (rows[:, 2] == ntags)
        n = int(np.argmin(same)) if not same.all() else rows.shape[0]
        n = min(n, ne - count)
        rows = rows[:n]
        blocks[_DIM[etype]].append((rows[:, 3], rows[:, 4], lut[rows[:, 3 + ntags:]]))
        pos += n * L
        count += n
    if count != ne:
        raise ValueError("malformed $Elements section: %d of %d elements parsed" % (count, ne))

    def cat(d, width):
        if not blocks[d]:
            return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((0, width), np.int64)
        return (np.concatenate([b[0] for b in blocks[d]]), np.concatenate([b[1] for b in blocks[d]]),
                np.concatenate([b[2] for b in blocks[d]]))

    return {"points": points[:, :dim] if dim == 2 else points, "names": names, "vol": cat(dim, dim + 1), "bnd": cat(dim - 1, dim)}


def read_msh(filename, mesh_dimensionality):
    """`ReadGmsh(filename, mesh_dimensionality)` -> remo3d_b200.mesh.Mesh."""
    if not filename.endswith(".msh"):
        filename += ".msh"
    with open(filename, "r") as f:
        p = parse_msh(f.read(), mesh_dimensionality)
    phys, elem, nodes = p["vol"]
    if (nodes < 0).any():
        raise ValueError("element refers to an undefined node")
    mat, _ = _first_appearance_index(elem) if elem.size else (np.zeros(0, np.int64), None)
    bphys, belem, bnodes = p["bnd"]
    if belem.size:
        bidx, border = _first_appearance_index(belem)
        first_phys = {int(t): int(bphys[np.argmax(belem == t)]) for t in border}
        bc_names = [p["names"].get(first_phys[int(t)], "default") for t in border]
        bc = bidx + 1
    else:
        bc, bc_names = np.zeros(0, np.int64), []
    return Mesh(p["points"], nodes, mat, bnodes, bc, bc_names)


def write_msh(filename, points, elems, elem_tags, bfacets, bfacet_tags, physical_names, node_ids=None, extra_points=()):
    """Minimal MSH 2.2 ASCII writer (tests / interchange).  *_tags: (physical, elementary) pairs per element;
    physical_names: list of (dim, tag, name)."""
    points = np.asarray(points, float)
    dim = points.shape[1]
    node_ids = np.arange(1, points.shape[0] + 1) if node_ids is None else np.asarray(node_ids)
    et_vol, et_bnd = (4, 2) if dim == 3 else (2, 1)
    with open(filename, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$PhysicalNames\n%d\n" % len(physical_names))
        for d, tag, name in physical_names:
            f.write('%d %d "%s"\n' % (d, tag, name))
        f.write("$EndPhysicalNames\n$Nodes\n%d\n" % points.shape[0])
        for i, pnt in zip(node_ids, points):
            xyz = list(pnt) + [0.0] * (3 - dim)
            f.write("%d %.17g %.17g %.17g\n" % (i, xyz[0], xyz[1], xyz[2]))
        f.write("$EndNodes\n$Elements\n%d\n" % (len(extra_points) + len(bfacets) + len(elems)))
        k = 1
        for v in extra_points:
            f.write("%d 15 2 0 %d %d\n" % (k, k, node_ids[v]))
            k += 1
        for e, (ph, el) in zip(bfacets, bfacet_tags):
            f.write("%d %d 2 %d %d %s\n" % (k, et_bnd, ph, el, " ".join(str(node_ids[v]) for v in e)))
            k += 1
        for e, (ph, el) in zip(elems, elem_tags):
            f.write("%d %d 2 %d %d %s\n" % (k, et_vol, ph, el, " ".join(str(node_ids[v]) for v in e)))
            k += 1
        f.write("$EndElements\n")
