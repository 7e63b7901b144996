"""Mesh container: the stand-in for `netgen.meshing.Mesh` / `ngsolve.Mesh` on this path.

The reference hands an NGSolve mesh to `SolveBVP` and uses exactly three things of it
(SURVEY §8b): `mesh.dim` (`ngsolve_functions.py:25`), `mesh(x, y[, z])` -> mesh point
(`ngsolve_functions.py:13-15`, `workers/worker.py:122-131`) and the boundary-condition
names/numbers that `dirichlet=` selects (`worker.py:90, 97`).  This class holds the same
information as flat arrays laid out the way the C-ABI takes them (SoA-friendly, C-contiguous):

    points   (nv, dim)   float64   vertex coordinates; z (depth, positive down) is the last column
    elems    (nt, dim+1) int32     tets (3D) / triangles (2D), 0-based vertex numbers
    mat      (nt,)       int32     0-based material index into the per-material sigma list
                                    (`worker.py:101`; order contract `gmsh_functions.py:172`)
    bfacets  (nb, dim)   int32     boundary triangles (3D) / segments (2D)
    bc       (nb,)       int32     1-based boundary-condition number of each facet
    bc_names list[str]             name of bc number i at index i-1
"""
import re

import numpy as np


class MeshPoint:
    """Result of `mesh(x, y[, z])`: only points on the electrode axis are supported on this path."""

    __slots__ = ("pnt", "z")

    def __init__(self, pnt):
        self.pnt = tuple(float(c) for c in pnt)
        self.z = self.pnt[-1]


class Mesh:
    def __init__(self, points, elems, mat, bfacets, bc, bc_names=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.elems = np.ascontiguousarray(elems, dtype=np.int32)
        self.mat = np.ascontiguousarray(mat, dtype=np.int32)
        self.bfacets = np.ascontiguousarray(bfacets, dtype=np.int32)
        self.bc = np.ascontiguousarray(bc, dtype=np.int32)
        self.dim = int(self.points.shape[1])
        if self.dim not in (2, 3):
            raise ValueError("mesh dimension must be 2 or 3")
        if self.elems.ndim != 2 or self.elems.shape[1] != self.dim + 1:
            raise ValueError("elements must have dim+1 vertices")
        if self.bfacets.size and self.bfacets.shape[1] != self.dim:
            raise ValueError("boundary facets must have dim vertices")
        if self.mat.shape[0] != self.elems.shape[0] or self.bc.shape[0] != self.bfacets.shape[0]:
            raise ValueError("material / bc arrays do not match the element arrays")
        nbc = int(self.bc.max()) if self.bc.size else 0
        self.bc_names = list(bc_names) if bc_names is not None else ["bc%d" % (i + 1) for i in range(nbc)]
        self._ctx_cache = None  # (device context, order) that currently holds this mesh; see fem.py

    # counts in NGSolve's vocabulary
    @property
    def nv(self):
        return self.points.shape[0]

    @property
    def ne(self):
        return self.elems.shape[0]

    @property
    def nmat(self):
        return int(self.mat.max()) + 1 if self.mat.size else 0

    def __call__(self, x=0.0, y=0.0, z=None):
        """`mesh(0, z)` in 2D, `mesh(0, 0, z)` in 3D (`ngsolve_functions.py:13-15`)."""
        pnt = (x, y) if self.dim == 2 else (x, y, z)
        if self.dim == 3 and z is None:
            raise TypeError("a 3D mesh point needs three coordinates")
        if any(abs(c) > 1e-9 for c in pnt[:-1]):
            raise ValueError("only points on the electrode axis (x=0[, y=0]) can be located on this path")
        return MeshPoint(pnt)

    def dirichlet_flags(self, dirichlet):
        """`dirichlet=` of ngs.H1 -> uint8 flag per boundary facet.

        Accepts what the reference passes (`worker.py:90, 97`): a boundary name (NGSolve treats it as
        a regular expression, alternatives separated by '|') or a list of 1-based bc numbers."""
        if dirichlet is None:
            return np.zeros(self.bc.shape[0], dtype=np.uint8)
        if isinstance(dirichlet, str):
            pat = re.compile(dirichlet)
            numbers = [i + 1 for i, name in enumerate(self.bc_names) if pat.fullmatch(name)]
        else:
            numbers = [int(i) for i in dirichlet]
        return np.isin(self.bc, numbers).astype(np.uint8)

    def axis_vertices(self, tol=1e-9):
        """Vertices on the electrode axis sorted by z (ties impossible on a valid mesh)."""
        on = np.all(np.abs(self.points[:, : self.dim - 1]) <= tol, axis=1)
        idx = np.nonzero(on)[0]
        return idx[np.argsort(self.points[idx, self.dim - 1], kind="stable")].astype(np.int32)
