"""Drop-in for `/root/reference/remo3d/ngsolve_functions.py`: same names, same call signatures."""
import numpy as np

from .fem import DEFAULT_ORDER, SolveBVPBatch


def AddPointSource(f, position, fac, model_dimensionality):
    """ngsolve_functions.py:10-21.  `f` is a list collecting (position, factor) pairs: on this path the
    right-hand side is built on the device from those pairs (remo_rhs_point_sources) when the system is
    solved, so the call only records the source."""
    if model_dimensionality not in (2, 3):
        raise ValueError("model_dimensionality must be 2 or 3")
    f.append((float(position), float(fac)))


def SolveBVP(mesh, sigma, tool_geometry, source_terms, dirichlet_boundary, preconditioner, condense, order=DEFAULT_ORDER):
    """ngsolve_functions.py:23-57 -> (fes, gfu).  One source configuration = a batch of one."""
    f = []
    for l in range(np.shape(source_terms)[0]):
        if source_terms[l] != 0.0:
            AddPointSource(f, tool_geometry[l], source_terms[l], mesh.dim)
    geom = np.array([p for p, _ in f])
    terms = np.array([s for _, s in f])
    fes, gfus = SolveBVPBatch(mesh, sigma, [(geom, terms)], dirichlet_boundary, preconditioner, condense, order=order)
    return fes, gfus[0]
