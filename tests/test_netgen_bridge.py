"""`netgen_bridge.mesh_arrays` against the REAL reference reader: `ReadGmsh`
(`/root/reference/remo3d/gmsh_functions.py:177-382`) was run by tests/golden/make_msh_golden.py with a recording stand-in for
`netgen.meshing`; here the recorded calls are replayed into a stand-in that offers Netgen's READ API (Points / Elements*D /
FaceDescriptor / GetBCName), and the arrays the bridge extracts must equal what this repo's own `.msh` reader produces for
the same file.  Then the reference's unmodified call sequence `SolveBVP -> gfu(mesh(0, 0, z))` (`worker.py:110-131`) runs
through the shim with the C ABI replaced by the oracle (CPU)."""
import json
import os

import numpy as np
import pytest

from remo3d_b200 import msh_reader, netgen_bridge


class _Pid:
    def __init__(self, nr):
        self.nr = nr


class _El:
    def __init__(self, index, vertices):
        self.index = index
        self.vertices = [_Pid(v) for v in vertices]


class _Pt:
    def __init__(self, p):
        self.p = tuple(p)


class _Fd:
    def __init__(self, bc):
        self.bc = bc


class FakeNetgenMesh:
    """Read side of netgen.meshing.Mesh, filled from the calls the reference's ReadGmsh made."""

    def __init__(self, dim, gold):
        self.dim = dim
        self._pts = [_Pt(p) for p in gold["points"]]
        self._vol = [_El(i, v) for i, v in gold["vol"]]
        self._bnd = [_El(i, v) for i, v in gold["bnd"]]
        self._names = {int(k): v for k, v in gold["bcnames"].items()}

    def Points(self):
        return iter(self._pts)

    def Elements3D(self):
        return iter(self._vol if self.dim == 3 else [])

    def Elements2D(self):
        return iter(self._bnd if self.dim == 3 else self._vol)

    def Elements1D(self):
        return iter([] if self.dim == 3 else self._bnd)

    def FaceDescriptor(self, i):
        return _Fd(i)  # ReadGmsh: FaceDescriptor(bc=index) is added as descriptor number `index` (gmsh_functions.py:296-302)

    def GetBCName(self, i):
        return self._names[i]


@pytest.mark.parametrize("name,dim", [("box3d", 3), ("disc2d", 2)])
def test_bridge_matches_msh_reader(golden_dir, name, dim):
    gold = json.load(open(os.path.join(golden_dir, "msh_golden.json")))[name]
    ref = msh_reader.read_msh(os.path.join(golden_dir, "msh", name + ".msh"), dim)
    ng = FakeNetgenMesh(dim, gold)
    for wrapped in (ng, type("NgsMesh", (), {"ngmesh": ng})()):  # netgen mesh and ngsolve.Mesh(mesh)
        xyz, elems, mat, bf, bdir, axis = netgen_bridge.mesh_arrays(wrapped, "dirichlet_boundary")
        np.testing.assert_array_equal(xyz, ref.points)
        np.testing.assert_array_equal(elems, ref.elems)
        np.testing.assert_array_equal(mat, ref.mat)
        np.testing.assert_array_equal(bf, ref.bfacets)
        np.testing.assert_array_equal(bdir, ref.dirichlet_flags("dirichlet_boundary"))
        np.testing.assert_array_equal(axis, ref.axis_vertices())
        assert bdir.sum() > 0
    # the Netgen 2D path selects the Dirichlet boundary by number (worker.py:97)
    nums = [i + 1 for i, n in enumerate(ref.bc_names) if n == "dirichlet_boundary"]
    np.testing.assert_array_equal(netgen_bridge.mesh_arrays(ng, nums)[4], ref.dirichlet_flags("dirichlet_boundary"))


def test_reference_call_sequence_through_the_shim(monkeypatch):
    """worker.py:110-131 unmodified: `fes, gfu = SolveBVP(mesh, sigma, tool_geometry, source_terms, dirichlet, precond, condense)`
    then `gfu(mesh(0.0, 0.0, z))` -- with a Netgen-style mesh object going in through the bridge and the oracle standing in
    for the GPU context."""
    from remo3d_b200 import fem, meshgen, ngsolve_functions as ngsf
    from tests import helpers

    pts, elems, bf, bc = meshgen.box_mesh(3, dirichlet=lambda c: (c[:, 0] > 1 - 1e-9) | (c[:, 1] > 1 - 1e-9) | (np.abs(c[:, 2]) > 1 - 1e-9))
    gold = {"points": [list(p) for p in pts], "vol": [[1, list(e + 1)] for e in elems], "bnd": [[int(b), list(f + 1)] for f, b in zip(bf, bc)],
            "bcnames": {"0": "natural", "1": "dirichlet_boundary"}}
    mesh = netgen_bridge.from_netgen(FakeNetgenMesh(3, gold))
    assert mesh.dim == 3 and mesh.nv == pts.shape[0]

    class _Ctx(helpers.OracleContext):
        def solve(self, rtol=1e-10, maxit=1000, raise_on_noconv=True):
            from oracle import fem_oracle as fo

            dim, points, e, mat, bfac, bdir = self.m
            space = fo.Space(points.shape[0], e, self.order, dim)
            A = fo.assemble(points, space, self.sigma, mat)
            con = space.dirichlet_dofs(bfac, np.asarray(bdir, bool))
            self.axis = fo.Axis(points, space)
            F = fo.point_source_rhs(self.axis, space.ndof, self.src[1], self.src[2])
            self.u = fo.solve_direct(A, F[:, None], con)[:, 0]
            return np.ones(1, np.int32), np.zeros(1)

        def sample_axis(self, z, rhs):
            from oracle import fem_oracle as fo

            return np.array([fo.sample_axis(self.axis, self.u, float(zz)) for zz in np.atleast_1d(z)])

    ctx = _Ctx()
    monkeypatch.setattr(fem, "default_context", lambda *a, **k: ctx, raising=False)
    tool_geometry = np.array([-0.5, 0.0, 0.25])
    source_terms = np.array([1.0, 0.0, 0.0])
    fes, gfu = ngsf.SolveBVP(mesh, [1.0], tool_geometry, source_terms, "dirichlet_boundary", "multigrid", True)
    u0 = gfu(mesh(0.0, 0.0, 0.0))
    u1 = gfu(mesh(0.0, 0.0, 0.25))
    assert u0 > u1 > 0.0
