"""Drop-in alias: `from remo3d import Model` (`/root/reference/remo3d/__init__.py:13`) resolves to remo3d_b200."""
from remo3d_b200.remo3d import Model  # noqa: F401
from remo3d_b200 import ngsolve_functions, ngsolve_functions_gpu  # noqa: F401
