"""Drop-in for `/root/reference/remo3d/ngsolve_functions_gpu.py` (ngscuda variant): same signature with the
trailing `solve_on` argument.  On this path everything (assembly included) runs on the GPU either way."""
from .fem import DEFAULT_ORDER
from .ngsolve_functions import AddPointSource, SolveBVP as _solve  # noqa: F401


def SolveBVP(mesh, sigma, tool_geometry, source_terms, dirichlet_boundary, preconditioner, condense, solve_on="CPU",
             order=DEFAULT_ORDER):
    """ngsolve_functions_gpu.py:15-54.  `solve_on` is accepted for compatibility (the reference moves only the
    CG to the device when it is "GPU"); there is no CPU path here."""
    if solve_on not in ("CPU", "GPU"):
        raise ValueError('solve_on must be "CPU" or "GPU"')
    return _solve(mesh, sigma, tool_geometry, source_terms, dirichlet_boundary, preconditioner, condense, order=order)
