"""Shared mesh / task fixtures for the parity tests (seeded, small enough for the oracle)."""
import numpy as np

from remo3d_b200 import meshgen, planner, tools as tl
from remo3d_b200.mesh import Mesh

SIX_TOOLS = ["B5.7A0.4M", "B4.48A1.62M", "M1.0A0.1B", "A2.0M0.5N", "N0.5M2.0A", "M4.0A0.5B"]


def box_case(n=3):
    """Unit-ish box, axis = edge x=y=0, z in [-1,1]; Dirichlet on the far faces x=1, y=1 and the ends z=+-1."""
    pts, elems, bf, bc = meshgen.box_mesh(n, dirichlet=lambda c: (c[:, 0] > 1 - 1e-9) | (c[:, 1] > 1 - 1e-9) | (np.abs(c[:, 2]) > 1 - 1e-9))
    cen = pts[elems].mean(axis=1)
    mat = (cen[:, 2] > 0).astype(np.int32) + (cen[:, 0] > 0.5).astype(np.int32)
    rng = np.random.default_rng(5)
    perm = rng.permutation(pts.shape[0])  # scramble vertex numbers: exercises the sorting of element vertices
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.shape[0])
    pts = pts[perm]
    elems = inv[elems].astype(np.int32)
    bf = inv[bf].astype(np.int32)
    mesh = Mesh(pts, elems, mat, bf, bc, ["natural", "dirichlet_boundary"])
    return mesh, [1.0, 0.25, 3.0]


def ball_case(h_electrode=0.08, h_axis=0.3, grading=0.5, tools=("A2.0M0.5N", "N0.5M2.0A", "B5.7A0.4M"), depths=(10.0, 10.1, 10.2),
              batch_size=5, layered=True, fsec=True, radius=50.0, seed=0):
    """Graded half-ball around the first task of a small plan; returns (mesh, sigma, flat task, tools params)."""
    params, sec = tl.set_tools_parameters(list(tools), fsec)
    _, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.asarray(depths, dtype=float), batch_size)
    task = tasks[0]
    ez = task[1][0]
    if layered:
        material = meshgen.layered_material([-1.0, 1.5], dip_rad=np.deg2rad(30.0), borehole_radius=0.1, invasion=[None, 0.4, None],
                                            inclusion=((3.0, 2.0, 1.0), 1.5))
        sigma = [1 / 1.0, 1 / 10.0, 1 / 5.0, 1 / 100.0, 1 / 10.0, 1 / 2.0]
    else:
        material, sigma = None, [1 / 10.0]
    m = meshgen.half_ball_mesh(radius, ez, material=material, h_electrode=h_electrode, h_axis=h_axis, grading=grading, seed=seed)
    mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
    flat = planner.flatten_task(task, params, three_d=True)
    return mesh, sigma, flat, params


def disc_case(h_electrode=0.03, h_axis=0.15, h_borehole=0.25, grading=0.6, tools=("A2.0M0.5N", "N0.5M2.0A", "B5.7A0.4M"),
              depths=(10.0, 10.1, 10.2), radius=50.0):
    """2D axisymmetric half-disc around the first task of a small plan: borehole with a caliper, three beds, one invaded."""
    from remo3d_b200 import meshgen2d

    params, sec = tl.set_tools_parameters(list(tools), True)
    _, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.asarray(depths, dtype=float), 5)
    task = tasks[0]
    wall = (np.array([-60.0, -3.0, 0.0, 2.0, 60.0]), np.array([0.1, 0.11, 0.1, 0.12, 0.1]))
    m = meshgen2d.half_disc_mesh(radius, task[1][0], wall, [-1.0, 1.5], [None, 0.4, None], h_electrode=h_electrode, h_axis=h_axis,
                                 h_borehole=h_borehole, grading=grading)
    mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
    sigma = [1 / 1.0, 1 / 10.0, 1 / 5.0, 1 / 100.0, 1 / 10.0]
    flat = planner.flatten_task(task, params, three_d=False)
    return mesh, sigma, flat, params


def star_case(nsphere=260, seed=3):
    """One central vertex surrounded by points on a sphere: the Delaunay mesh has ~2 * nsphere tets that all share the
    centre, so its matrix row has thousands of candidate columns (the whole-CTA path of the CSR pattern builder)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    v = rng.standard_normal((nsphere, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    pts = np.vstack([np.zeros((1, 3)), v])
    elems = Delaunay(pts).simplices.astype(np.int32)
    # positive orientation is not required by the path (|det| is used), vertex order is sorted on the device
    bf = meshgen.boundary_facets(elems)
    bc = np.full(bf.shape[0], 2, dtype=np.int32)
    mat = (pts[elems].mean(axis=1)[:, 2] > 0).astype(np.int32)
    return Mesh(pts, elems, mat, bf, bc, ["natural", "dirichlet_boundary"]), [1.0, 0.1]


class OracleContext:
    """Stand-in for `_cabi.Context` backed by the CPU oracle: lets the host pipeline (`worker.run_tasks`,
    `Model.simulate_logs`) run in the CPU test-suite.  Test infrastructure only."""

    def __init__(self, fail_on=()):
        self.fail_on = set(fail_on)  # task counter values at which solve() raises
        self.calls = 0
        self.ndof = 0

    def mesh_set(self, dim, points, elems, mat, bfacets, bdir, axis):
        self.m = (dim, points, elems, mat, bfacets, bdir)

    def space_build(self, order):
        self.order = order
        return 0, 0

    def assemble(self, sigma):
        self.sigma = np.asarray(sigma, dtype=float)

    def precond_setup(self, kind):
        self.kind = kind

    def rhs_point_sources(self, src_ptr, src_z, src_fac):
        self.src = (np.asarray(src_ptr), np.asarray(src_z), np.asarray(src_fac))

    def solve(self, rtol=1e-10, maxit=1000, raise_on_noconv=True):
        self.calls += 1
        if self.calls in self.fail_on:
            raise RuntimeError("injected solver failure")
        n = self.src[0].shape[0] - 1
        return np.ones(n, np.int32), np.zeros(n)

    def apparent_resistivity(self, pt_rhs, z0, z1, k, scale, out=None):
        from oracle import fem_oracle as fo

        dim, points, elems, mat, bfacets, bdir = self.m
        flat = {"src_ptr": self.src[0], "src_z": self.src[1], "src_fac": self.src[2], "pt_rhs": np.asarray(pt_rhs), "pt_z0": np.asarray(z0),
                "pt_z1": np.asarray(z1), "pt_k": np.asarray(k), "scale": scale}
        return fo.solve_task(points, elems, mat, self.sigma, bfacets, np.asarray(bdir, bool), self.order, flat, dim=dim, solver="direct")["ra"]

    def stage_times(self):
        return {}

    def close(self):
        pass
