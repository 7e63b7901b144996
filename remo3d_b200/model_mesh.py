"""Formation / borehole model -> mesh + per-material conductivities for one mesh task (host side).

Stands in for `SelectGmshDataRange` + `ConstructGmsh3dModel` (`/root/reference/remo3d/gmsh_functions.py:10-174,
544-684`) and their Netgen twins, which need Gmsh / Netgen (not installable here; mesh generation is out of the
GPU hot path).  What is preserved is the contract the solve depends on:

  * geometry is shifted so the batch's mean source depth is z = 0 (`gmsh_functions.py:38, 102`);
  * material order: 0 = borehole mud, then per layer top -> bottom the flushed zone (if the layer has one) and
    the undisturbed zone; sigma = [1/Rm(batch depth)] + [1/rho ...] in that order (`gmsh_functions.py:160-172`);
  * dipping beds are the planes z = z_i + tan(dip) x (`gmsh_functions.py:112-128`);
  * half-ball of radius `domain_radius`, Dirichlet on the sphere, every electrode of the batch a mesh vertex.

dip = 0 models use the conforming 2D axisymmetric mesher (meshgen2d.py).  In 3D (dip > 0) `mesh_options["conforming"]`
(the `Model` default) makes the point cloud carry the interfaces -- lattice points near the borehole wall, a layer plane or an
invasion cylinder are projected onto it, and the borehole column is resolved some metres beyond the tool
(`meshgen.Interfaces`, `SizeField` h_borehole term) -- so the Delaunay tets have them as faces; without it the material is
simply taken per tet centroid on a formation-independent triangulation that all tasks of one electrode pattern share.
Measured against the reference's committed Example_01 log (order 2, depth 5.5, A2.0M0.5N): -2.3 % per centroid, -0.6 % with
the interfaces at the same sizes, 2e-5 with the near field of tests/test_gpu_golden_example01.py (profiles/r02_notes.md).
"""
import numpy as np

from . import meshgen, meshgen2d
from .mesh import Mesh

DEFAULT_MESH_OPTIONS = {"h_electrode": 0.02, "h_axis": 0.06, "grading": 0.3, "h_max": None, "seed": 0}
DEFAULT_MESH_OPTIONS_2D = {"h_electrode": 0.01, "h_axis": 0.06, "h_borehole": 0.1, "grading": 0.4, "h_max": None, "seed": 0}


# how far and how finely the borehole column is resolved beyond the tool in conforming 3D meshes (meshgen.SizeField): size
# h_borehole + g_borehole * (distance to the wall) + g_window * (axial distance to the tool beyond borehole_window metres);
# a 5 m window gives the same Ra as the whole 100 m column to 1e-3 at 1/10 of the vertices (profiles/r02_notes.md)
BOREHOLE_COLUMN = {"g_borehole": 0.8, "borehole_window": 5.0, "g_window": 0.15}
KNOWN_MESH_OPTIONS = set(DEFAULT_MESH_OPTIONS) | set(DEFAULT_MESH_OPTIONS_2D) | {"msh_path", "conforming", "g_borehole", "borehole_window", "g_window", "jitter"}


def task_sigma(formation, mud_resistivity):
    """Per-material conductivities in the reference's order (`gmsh_functions.py:160-161, 172`)."""
    res = np.ndarray.flatten(np.asarray(formation)[:, 3:5])
    res = res[~np.isnan(res)]
    return [1.0 / mud_resistivity] + list(1.0 / res)


def geometry_key(dip_rad, electrodes_z, domain_radius, mesh_options=None):
    """Hashable key of everything the 3D tet GEOMETRY of a task depends on.  The half-ball mesher places points from the
    electrode pattern, the radius and the size options only (materials are assigned per tet afterwards), so all tasks
    with the same electrode pattern -- 11 patterns for 413 tasks in config C5 (1000 depths x 4 tools) -- share one
    triangulation.  None for dip == 0: the conforming 2D meshes follow the formation and are unique per task."""
    if np.isclose(dip_rad, 0.0):
        return None
    opts = dict(DEFAULT_MESH_OPTIONS)
    opts.update({k: v for k, v in (mesh_options or {}).items() if k in DEFAULT_MESH_OPTIONS or k.startswith("improve")})
    return (float(domain_radius), tuple(np.round(np.asarray(electrodes_z, dtype=float), 6)), tuple(sorted(opts.items())))


def build_geometry_3d(electrodes_z, domain_radius, mesh_options=None):
    """The formation-independent part of a 3D task mesh: points, tets, boundary facets (+ tet centroids for the material
    predicate)."""
    opts = dict(DEFAULT_MESH_OPTIONS)
    opts.update({k: v for k, v in (mesh_options or {}).items() if k in DEFAULT_MESH_OPTIONS or k.startswith("improve")})
    m = meshgen.half_ball_mesh(float(domain_radius), np.asarray(electrodes_z, dtype=float), material=None, **opts)
    m["centroids"] = m["points"][m["elems"]].mean(axis=1)
    return m


def assign_materials_3d(geometry, formation, borehole_geometry, dip_rad, centre_depth):
    """Mesh of one task from a (shared) geometry: only the per-tet material index depends on the depth of the batch."""
    formation = np.asarray(formation, dtype=float)
    tops = formation[1:, 0] - centre_depth  # interfaces between consecutive layers, relative depth
    invasion = [None if np.isnan(r) else float(r) for r in formation[:, 2]]
    caliper = (np.asarray(borehole_geometry)[:, 0] - centre_depth, np.asarray(borehole_geometry)[:, 1])
    material = meshgen.layered_material(tops, dip_rad=dip_rad, borehole_radius=caliper, invasion=invasion)
    mat = np.asarray(material(geometry["centroids"]), dtype=np.int32)
    return Mesh(geometry["points"], geometry["elems"], mat, geometry["bfacets"], geometry["bc"], geometry["bc_names"])


def read_task_mesh(path, dim):
    """A task mesh produced by an external Gmsh run (`mesh_generator="gmsh"` with `mesh_options={"msh_path": ...}`): MSH 2.2
    ASCII, numbering contract of the reference's `ReadGmsh` (`gmsh_functions.py:177-382`, msh_reader.py)."""
    from . import msh_reader

    return msh_reader.read_msh(path, dim)


def build_task_mesh(formation, borehole_geometry, dip_rad, centre_depth, electrodes_z, mud_resistivity, domain_radius,
                    mesh_options=None, geometry=None, task_index=None):
    """-> (Mesh, sigma list) for the batch centred at `centre_depth` with electrodes at relative depths `electrodes_z`.
    dip == 0 -> 2D axisymmetric (r, z) half-disc, interfaces meshed conformingly (the reference's Netgen/Gmsh 2D path,
    `remo3d.py:776-784`); dip > 0 -> 3D half-ball (`geometry`: a shared triangulation from build_geometry_3d).
    `mesh_options["msh_path"]` (a format string taking the task index, or a callable(task_index, centre_depth,
    electrodes_z) -> path) reads the task's mesh from a Gmsh `.msh` file instead (`worker.py:82-92`)."""
    formation = np.asarray(formation, dtype=float)
    unknown = [k for k in (mesh_options or {}) if k not in KNOWN_MESH_OPTIONS and not k.startswith("improve")]
    if unknown:
        raise ValueError("unknown mesh option(s): %s" % ", ".join(sorted(unknown)))
    msh_path = (mesh_options or {}).get("msh_path")
    if msh_path is not None:
        path = msh_path(task_index, centre_depth, electrodes_z) if callable(msh_path) else str(msh_path).format(task_index)
        return read_task_mesh(path, 2 if np.isclose(dip_rad, 0.0) else 3), task_sigma(formation, mud_resistivity)
    if np.isclose(dip_rad, 0.0):
        opts = dict(DEFAULT_MESH_OPTIONS_2D)
        opts.update({k: v for k, v in (mesh_options or {}).items() if k in DEFAULT_MESH_OPTIONS_2D or k == "jitter"})
        wall = (np.asarray(borehole_geometry)[:, 0] - centre_depth, np.asarray(borehole_geometry)[:, 1])
        invasion = [None if np.isnan(r) else float(r) for r in formation[:, 2]]
        m = meshgen2d.half_disc_mesh(float(domain_radius), np.asarray(electrodes_z, dtype=float), wall, formation[1:, 0] - centre_depth,
                                     invasion, **opts)
        mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
        return mesh, task_sigma(formation, mud_resistivity)
    if geometry is None and (mesh_options or {}).get("conforming", False):
        # interfaces carried by the mesh (borehole wall, layer planes, invasion cylinders: `gmsh_functions.py:576-624`): the
        # point cloud depends on the depth of the batch, so the triangulation is built per task and cannot be shared
        tops = formation[1:, 0] - centre_depth
        invasion = [None if np.isnan(r) else float(r) for r in formation[:, 2]]
        caliper = (np.asarray(borehole_geometry)[:, 0] - centre_depth, np.asarray(borehole_geometry)[:, 1])
        opts = dict(DEFAULT_MESH_OPTIONS)
        opts.update({k: v for k, v in (mesh_options or {}).items() if k in DEFAULT_MESH_OPTIONS or k.startswith("improve")})
        r_strip = max([float(np.max(caliper[1]))] + [v for v in invasion if v is not None])
        m = meshgen.half_ball_mesh(float(domain_radius), np.asarray(electrodes_z, dtype=float), material=None,
                                   interfaces=meshgen.Interfaces(caliper, tops, dip_rad, invasion),
                                   h_borehole=float((mesh_options or {}).get("h_borehole", 0.15)), r_strip=r_strip,
                                   **{k: float((mesh_options or {}).get(k, v)) for k, v in BOREHOLE_COLUMN.items()}, **opts)
        m["centroids"] = m["points"][m["elems"]].mean(axis=1)
        geometry = m
    if geometry is None:
        geometry = build_geometry_3d(electrodes_z, domain_radius, mesh_options)
    return assign_materials_3d(geometry, formation, borehole_geometry, dip_rad, centre_depth), task_sigma(formation, mud_resistivity)
