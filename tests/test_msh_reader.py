"""The .msh 2.2 reader against the REAL reference reader (`gmsh_functions.py:177-382`, run by tests/golden/make_msh_golden.py
with a recording stub for netgen.meshing): vertex order, element order, node order, material / bc indices and bc names."""
import json
import os

import numpy as np
import pytest

from remo3d_b200 import msh_reader


@pytest.mark.parametrize("name,dim", [("box3d", 3), ("disc2d", 2)])
def test_reader_numbering_matches_reference(golden_dir, name, dim):
    gold = json.load(open(os.path.join(golden_dir, "msh_golden.json")))[name]
    mesh = msh_reader.read_msh(os.path.join(golden_dir, "msh", name + ".msh"), dim)
    ref_pts = np.array(gold["points"])[:, :dim]
    np.testing.assert_array_equal(mesh.points, ref_pts)  # bit-exact coordinates, $Nodes order
    vol_idx = np.array([e[0] for e in gold["vol"]])
    vol_nodes = np.array([e[1] for e in gold["vol"]])
    np.testing.assert_array_equal(mesh.elems, vol_nodes - 1)  # netgen PointIds are 1-based
    np.testing.assert_array_equal(mesh.mat, vol_idx - 1)      # material numbers 1-based in netgen, 0-based here
    bnd_idx = np.array([e[0] for e in gold["bnd"]])
    bnd_nodes = np.array([e[1] for e in gold["bnd"]])
    np.testing.assert_array_equal(mesh.bfacets, bnd_nodes - 1)
    np.testing.assert_array_equal(mesh.bc, bnd_idx)
    for k, v in gold["bcnames"].items():  # SetBCName(index - 1, name)
        assert mesh.bc_names[int(k)] == v
    assert mesh.dirichlet_flags("dirichlet_boundary").sum() == (bnd_idx == 1 + [v for _, v in sorted((int(k), v) for k, v in gold["bcnames"].items())].index("dirichlet_boundary")).sum()


def test_write_read_roundtrip_large(tmp_path):
    from remo3d_b200 import meshgen

    pts, elems, bf, bc = meshgen.box_mesh(12)
    tags = [(1, 1)] * elems.shape[0]
    path = str(tmp_path / "big.msh")
    msh_reader.write_msh(path, pts, elems, tags, bf, [(2, int(b)) for b in bc], [(3, 1, "all"), (2, 2, "dirichlet_boundary")])
    mesh = msh_reader.read_msh(path, 3)
    np.testing.assert_array_equal(mesh.points, pts)
    np.testing.assert_array_equal(mesh.elems, elems)
    np.testing.assert_array_equal(mesh.bfacets, bf)
    assert mesh.nmat == 1 and mesh.bc_names == ["dirichlet_boundary"]


def test_reader_rejects_unsupported(tmp_path):
    p = tmp_path / "q.msh"
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n1\n1 5 2 1 1 1 1 1 1 1 1 1 1\n$EndElements\n")
    with pytest.raises(ValueError):
        msh_reader.read_msh(str(p), 3)
    p.write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError):
        msh_reader.read_msh(str(p), 3)
