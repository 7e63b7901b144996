"""Netgen / NGSolve mesh -> the flat arrays of the C ABI: the glue INTEGRATION.md's `SolveBVP` stub calls.

The reference hands `ngs.Mesh(mesh)` to `SolveBVP` (`workers/worker.py:100, 110`); `mesh` comes from Netgen's 2D generator
(`netgen_functions.py:329`) or from `ReadGmsh` (`gmsh_functions.py:177-382`, which fills a `netgen.meshing.Mesh` with
`MeshPoint`, `Element1D/2D/3D`, `FaceDescriptor`, `SetBCName`, `SetMaterial`).  This module reads the same object back
through Netgen's Python API and never imports netgen itself, so it is testable with a stand-in (tests/test_netgen_bridge.py):

    ngmesh.dim                                  2 or 3
    ngmesh.Points()                             iterable of MeshPoint, `.p` = (x, y, z); PointIds are 1-based, in this order
    ngmesh.Elements3D() / Elements2D() / Elements1D()
        el.vertices                             PointIds (`.nr`, 1-based; plain ints are accepted too)
        el.index                                3D: volume element -> 1-based material number (`SetMaterial(index, ..)`),
                                                    surface element -> 1-based FaceDescriptor number
                                                2D: triangle -> 1-based material number, segment -> 1-based bc number
    ngmesh.FaceDescriptor(i).bc                 3D only: 1-based boundary-condition number of face descriptor i
    ngmesh.GetBCName(i)                         name of bc number i + 1 (`SetBCName(index - 1, name)`, gmsh_functions.py:301, 328)

Conventions of the result (mesh.py): 0-based vertices and materials, 1-based bc numbers, in 2D points = (r, z) = Netgen's
(x, y) (`ngsolve_functions.py:13, 33`: the axis is x = 0 and the weight of the axisymmetric form is ngs.x).
"""
import numpy as np

from .mesh import Mesh


def _nr(v):
    return int(getattr(v, "nr", v))


def _coords(p):
    q = getattr(p, "p", p)
    return (float(q[0]), float(q[1]), float(q[2]))


def from_netgen(mesh):
    """`netgen.meshing.Mesh` or `ngsolve.Mesh` (its `.ngmesh`) -> `remo3d_b200.mesh.Mesh`."""
    ng = getattr(mesh, "ngmesh", mesh)
    dim = int(ng.dim)
    if dim not in (2, 3):
        raise ValueError("mesh dimension must be 2 or 3")
    pts = np.array([_coords(p) for p in ng.Points()], dtype=np.float64).reshape(-1, 3)[:, :dim]
    vol = list(ng.Elements3D() if dim == 3 else ng.Elements2D())
    bnd = list(ng.Elements2D() if dim == 3 else ng.Elements1D())
    elems = np.array([[_nr(v) - 1 for v in el.vertices][: dim + 1] for el in vol], dtype=np.int32).reshape(-1, dim + 1)
    mat = np.array([int(el.index) - 1 for el in vol], dtype=np.int32)
    bf = np.array([[_nr(v) - 1 for v in el.vertices][:dim] for el in bnd], dtype=np.int32).reshape(-1, dim)
    if dim == 3:
        fd_bc = {}
        bc = np.empty(len(bnd), dtype=np.int32)
        for i, el in enumerate(bnd):
            k = int(el.index)
            if k not in fd_bc:
                fd_bc[k] = int(ng.FaceDescriptor(k).bc)
            bc[i] = fd_bc[k]
    else:
        bc = np.array([int(el.index) for el in bnd], dtype=np.int32)
    names = []
    for i in range(int(bc.max()) if bc.size else 0):
        try:
            names.append(str(ng.GetBCName(i)))
        except Exception:  # noqa: BLE001 -- unnamed bc numbers (Netgen's own 2D generator names none: worker.py:97 uses [2])
            names.append("bc%d" % (i + 1))
    return Mesh(pts, elems, mat, bf, bc, names)


def mesh_arrays(mesh, dirichlet_boundary):
    """What INTEGRATION.md's stub passes to `remo_mesh_set`: (xyz, elems, mat, bfacets, bdir, axis).  `dirichlet_boundary` is
    what the reference passes as `dirichlet=` (`worker.py:90, 97`): the name 'dirichlet_boundary' or the bc numbers [2]."""
    m = mesh if isinstance(mesh, Mesh) else from_netgen(mesh)
    return m.points, m.elems, m.mat, m.bfacets, m.dirichlet_flags(dirichlet_boundary), m.axis_vertices()
