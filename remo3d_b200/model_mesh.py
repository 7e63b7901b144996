"""Formation / borehole model -> mesh + per-material conductivities for one mesh task (host side).

Stands in for `SelectGmshDataRange` + `ConstructGmsh3dModel` (`/root/reference/remo3d/gmsh_functions.py:10-174,
544-684`) and their Netgen twins, which need Gmsh / Netgen (not installable here; mesh generation is out of the
GPU hot path).  What is preserved is the contract the solve depends on:

  * geometry is shifted so the batch's mean source depth is z = 0 (`gmsh_functions.py:38, 102`);
  * material order: 0 = borehole mud, then per layer top -> bottom the flushed zone (if the layer has one) and
    the undisturbed zone; sigma = [1/Rm(batch depth)] + [1/rho ...] in that order (`gmsh_functions.py:160-172`);
  * dipping beds are the planes z = z_i + tan(dip) x (`gmsh_functions.py:112-128`);
  * half-ball of radius `domain_radius`, Dirichlet on the sphere, every electrode of the batch a mesh vertex.

dip = 0 models use the conforming 2D axisymmetric mesher (meshgen2d.py).  In 3D (dip > 0) interfaces are not meshed
conformingly: the material is assigned per tet from its centroid (documented in DESIGN.md).
"""
import numpy as np

from . import meshgen, meshgen2d
from .mesh import Mesh

DEFAULT_MESH_OPTIONS = {"h_electrode": 0.02, "h_axis": 0.06, "grading": 0.3, "h_max": None, "seed": 0}
DEFAULT_MESH_OPTIONS_2D = {"h_electrode": 0.01, "h_axis": 0.06, "h_borehole": 0.1, "grading": 0.4, "h_max": None, "seed": 0}


def task_sigma(formation, mud_resistivity):
    """Per-material conductivities in the reference's order (`gmsh_functions.py:160-161, 172`)."""
    res = np.ndarray.flatten(np.asarray(formation)[:, 3:5])
    res = res[~np.isnan(res)]
    return [1.0 / mud_resistivity] + list(1.0 / res)


def build_task_mesh(formation, borehole_geometry, dip_rad, centre_depth, electrodes_z, mud_resistivity, domain_radius,
                    mesh_options=None):
    """-> (Mesh, sigma list) for the batch centred at `centre_depth` with electrodes at relative depths `electrodes_z`.
    dip == 0 -> 2D axisymmetric (r, z) half-disc, interfaces meshed conformingly (the reference's Netgen/Gmsh 2D path,
    `remo3d.py:776-784`); dip > 0 -> 3D half-ball."""
    formation = np.asarray(formation, dtype=float)
    if np.isclose(dip_rad, 0.0):
        opts = dict(DEFAULT_MESH_OPTIONS_2D)
        opts.update(mesh_options or {})
        wall = (np.asarray(borehole_geometry)[:, 0] - centre_depth, np.asarray(borehole_geometry)[:, 1])
        invasion = [None if np.isnan(r) else float(r) for r in formation[:, 2]]
        m = meshgen2d.half_disc_mesh(float(domain_radius), np.asarray(electrodes_z, dtype=float), wall, formation[1:, 0] - centre_depth,
                                     invasion, **opts)
        mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
        return mesh, task_sigma(formation, mud_resistivity)
    opts = dict(DEFAULT_MESH_OPTIONS)
    opts.update(mesh_options or {})
    tops = formation[1:, 0] - centre_depth  # interfaces between consecutive layers, relative depth
    invasion = [None if np.isnan(r) else float(r) for r in formation[:, 2]]
    caliper = (np.asarray(borehole_geometry)[:, 0] - centre_depth, np.asarray(borehole_geometry)[:, 1])
    material = meshgen.layered_material(tops, dip_rad=dip_rad, borehole_radius=caliper, invasion=invasion)
    m = meshgen.half_ball_mesh(float(domain_radius), np.asarray(electrodes_z, dtype=float), material=material, **opts)
    mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
    return mesh, task_sigma(formation, mud_resistivity)
