"""Stand-ins for the NGSolve objects the reference's hot path touches, backed by the CUDA library.

`FESpace` ~ `ngs.H1(mesh, order=3, dirichlet=...)`      (ngsolve_functions.py:27)
`GridFunction` ~ `ngs.GridFunction(fes)`; callable on mesh points like `gfu(mesh(0.0, z))` (worker.py:122-131)
`SolveBVPBatch` = the batched form of `SolveBVP` (ngsolve_functions.py:23-57): one assembly, one
preconditioner setup and one multi-right-hand-side PCG for all sources that share a mesh.
"""
import threading

import numpy as np

from . import _cabi
from .mesh import Mesh, MeshPoint

DEFAULT_ORDER = 3      # ngsolve_functions.py:27 hard-codes order=3
DEFAULT_RTOL = 1e-10   # BASELINE.json north_star: CG relative residual 1e-10
DEFAULT_MAXIT = 1000   # ngsolve_functions.py:50 maxsteps=1000

_tls = threading.local()


def default_context(device=None):
    """One lazily created context per (thread, device)."""
    if device is None:
        device = getattr(_tls, "device", 0)
    ctxs = getattr(_tls, "ctxs", None)
    if ctxs is None:
        ctxs = _tls.ctxs = {}
    if device not in ctxs:
        ctxs[device] = _cabi.Context(device)
    return ctxs[device]


def set_default_device(device):
    _tls.device = int(device)


class FESpace:
    """H1 space of one mesh on one device context (dof numbering: SURVEY 10.2)."""

    def __init__(self, mesh, order=DEFAULT_ORDER, dirichlet=None, ctx=None):
        if not isinstance(mesh, Mesh):
            raise TypeError("mesh must be a remo3d_b200.mesh.Mesh")
        self.mesh, self.order = mesh, int(order)
        self.ctx = ctx or default_context()
        flags = mesh.dirichlet_flags(dirichlet)
        self.ctx.mesh_set(mesh.dim, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, flags, mesh.axis_vertices())
        self.ndof, _ = self.ctx.space_build(self.order)
        self.generation = _next_generation(self.ctx)

    @property
    def nnz(self):
        return self.ctx.nnz

    def FreeDofs(self):
        return ~self.ctx.dirichlet()


def _next_generation(ctx):
    ctx._generation = getattr(ctx, "_generation", 0) + 1
    return ctx._generation


class GridFunction:
    """Solution of one right-hand side; lives on the device, evaluated on the axis on demand."""

    def __init__(self, space, rhs_index):
        self.space, self.rhs_index = space, int(rhs_index)

    def _check(self):
        if getattr(self.space.ctx, "_generation", None) != self.space.generation:
            raise RuntimeError("this GridFunction's device data was replaced by a later solve on the same context")

    def __call__(self, mp):
        self._check()
        z = mp.z if isinstance(mp, MeshPoint) else float(mp)
        return float(self.space.ctx.sample_axis([z], [self.rhs_index])[0])

    @property
    def vec(self):
        self._check()
        return self.space.ctx.solution(self.rhs_index)


def SolveBVPBatch(mesh, sigma, sources, dirichlet_boundary, preconditioner="multigrid", condense=True, order=DEFAULT_ORDER,
                  rtol=DEFAULT_RTOL, maxit=DEFAULT_MAXIT, ctx=None):
    """Solve -div(sigma grad u) = sum_k s_k delta(r - r_k) for several source configurations on one mesh.

    sources: list of (tool_geometry, source_terms) pairs, each as in `SolveBVP` (1-D arrays of equal length;
    only non-zero source terms inject current, ngsolve_functions.py:41-44).
    `condense` is accepted for signature compatibility: static condensation (ngsolve_functions.py:31,53-56)
    changes the algebra, not the solution, and tets of order <= 3 have no interior dofs to condense.
    Returns (fes, [gfu per source])."""
    if preconditioner not in _cabi.PRECOND:
        raise ValueError("preconditioner must be 'local' or 'multigrid'")
    if len(sources) < 1:
        raise ValueError("at least one source configuration is required")
    fes = FESpace(mesh, order=order, dirichlet=dirichlet_boundary, ctx=ctx)
    c = fes.ctx
    c.assemble(np.asarray([float(s) for s in sigma], dtype=np.float64))
    c.precond_setup(preconditioner)
    out = []
    for lo in range(0, len(sources), _cabi.MAX_RHS):
        if lo > 0:
            raise ValueError("more than %d source configurations per mesh are not supported in one batch" % _cabi.MAX_RHS)
        chunk = sources[lo:lo + _cabi.MAX_RHS]
        ptr, zs, fs = [0], [], []
        for geom, terms in chunk:
            geom, terms = np.asarray(geom, dtype=float), np.asarray(terms, dtype=float)
            if geom.shape != terms.shape:
                raise ValueError("tool_geometry and source_terms must have the same length")
            nz = terms != 0.0
            zs += list(geom[nz])
            fs += list(terms[nz])
            ptr.append(len(zs))
        c.rhs_point_sources(ptr, zs, fs)
        fes.iterations, fes.relres = c.solve(rtol=rtol, maxit=maxit)
        out += [GridFunction(fes, r) for r in range(len(chunk))]
    return fes, out
