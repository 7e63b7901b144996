"""Golden vectors for the .msh reader: run the REAL reference `ReadGmsh` (/root/reference/remo3d/gmsh_functions.py:177-382)
on sample files written by this repo, with a recording stub in place of `netgen.meshing` (Netgen is not installable here;
the reader only calls constructors and `mesh.Add/Set*`, which the stub records verbatim) and an empty stub for `gmsh`.

Run in the build container:  python tests/golden/make_msh_golden.py   -> tests/golden/msh/*.msh + msh_golden.json"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


class _Rec:
    def __init__(self, dim):
        self.dim = dim
        self.points, self.el1, self.el2, self.el3 = [], [], [], []
        self.bcnames, self.materials, self.cd2 = {}, {}, {}

    def Add(self, obj):
        kind = obj[0]
        if kind == "pt":
            self.points.append(obj[1])
            return len(self.points)  # 1-based PointId
        if kind == "fd":
            return None
        {"e1": self.el1, "e2": self.el2, "e3": self.el3}[kind].append([obj[1], list(obj[2])])

    def SetBCName(self, i, name):
        self.bcnames[i] = name

    def SetMaterial(self, i, name):
        self.materials[i] = name

    def SetCD2Name(self, i, name):
        self.cd2[i] = name


class _FD:
    def __init__(self, bc):
        self.bc = bc
        self.bcname = None

    def __getitem__(self, i):
        return ("fd",)[i]


def install_stubs():
    sys.modules["gmsh"] = types.ModuleType("gmsh")
    ng = types.ModuleType("netgen")
    m = types.ModuleType("netgen.meshing")
    m.Mesh = lambda dim: _Rec(dim)
    m.Pnt = lambda x, y, z: (x, y, z)
    m.MeshPoint = lambda p: ("pt", p)
    m.FaceDescriptor = _FD
    m.Element1D = lambda index, vertices: ("e1", index, vertices)
    m.Element2D = lambda index, vertices: ("e2", index, vertices)
    m.Element3D = lambda index, vertices: ("e3", index, vertices)
    ng.meshing = m
    sys.modules["netgen"] = ng
    sys.modules["netgen.meshing"] = m


def main():
    install_stubs()
    sys.path.insert(0, "/root/reference/remo3d")
    import gmsh_functions as ref

    from remo3d_b200 import meshgen, meshgen2d, msh_reader

    out_dir = os.path.join(HERE, "msh")
    os.makedirs(out_dir, exist_ok=True)
    gold = {}
    rng = np.random.default_rng(3)

    # 3D sample: box, 3 material regions with scrambled elementary tags, two boundary groups, non-contiguous node ids
    pts, elems, bf, bc = meshgen.box_mesh(2, dirichlet=lambda c: c[:, 0] > 1 - 1e-9)
    cen = pts[elems].mean(axis=1)
    el_tag = np.where(cen[:, 2] > 0.3, 7, np.where(cen[:, 0] > 0.5, 3, 12))
    ph_tag = np.where(el_tag == 7, 101, np.where(el_tag == 3, 102, 103))
    perm = rng.permutation(elems.shape[0])
    elems, el_tag, ph_tag = elems[perm], el_tag[perm], ph_tag[perm]
    b_el = np.where(bc == 2, 40, np.where(pts[bf].mean(axis=1)[:, 2] > 0, 41, 39))
    b_ph = np.where(bc == 2, 201, 202)
    node_ids = np.sort(rng.choice(np.arange(1, 3 * pts.shape[0]), pts.shape[0], replace=False))
    names = [(2, 201, "dirichlet_boundary"), (2, 202, "neumann"), (3, 101, "mud"), (3, 102, "layer_a"), (3, 103, "layer_b")]
    f3 = os.path.join(out_dir, "box3d.msh")
    msh_reader.write_msh(f3, pts, elems, list(zip(ph_tag, el_tag)), bf, list(zip(b_ph, b_el)), names, node_ids, extra_points=[0, 5])
    m = ref.ReadGmsh(f3, 3)
    gold["box3d"] = {"points": m.points, "vol": m.el3, "bnd": m.el2, "bcnames": {str(k): v for k, v in m.bcnames.items()},
                     "materials": {str(k): v for k, v in m.materials.items()}}

    # 2D sample: small half-disc
    d = meshgen2d.half_disc_mesh(5.0, [-0.5, 0.0, 0.8], (np.array([-6.0, 6.0]), np.array([0.1, 0.12])), [-0.3, 1.0], [None, 0.4, None],
                                 h_electrode=0.08, h_axis=0.3, h_borehole=0.3, grading=0.8, h_max=2.0)
    el_tag = np.array([11, 5, 9, 2, 30])[d["mat"]]
    ph_tag = 300 + d["mat"]
    names = [(1, 201, "axis"), (1, 202, "dirichlet_boundary")] + [(2, 300 + i, "mat%d" % i) for i in range(5)]
    f2 = os.path.join(out_dir, "disc2d.msh")
    msh_reader.write_msh(f2, d["points"], d["elems"], list(zip(ph_tag, el_tag)), d["bfacets"], list(zip(200 + d["bc"], 50 + d["bc"])), names)
    m = ref.ReadGmsh(f2, 2)
    gold["disc2d"] = {"points": m.points, "vol": m.el2, "bnd": m.el1, "bcnames": {str(k): v for k, v in m.bcnames.items()},
                      "materials": {str(k): v for k, v in m.materials.items()}}
    with open(os.path.join(HERE, "msh_golden.json"), "w") as f:
        json.dump(gold, f)
    print("wrote msh golden:", {k: (len(v["points"]), len(v["vol"]), len(v["bnd"])) for k, v in gold.items()})


if __name__ == "__main__":
    main()
