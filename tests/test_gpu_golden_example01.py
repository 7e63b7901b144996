"""End-to-end pin against the reference's own committed output: `Model.compute_synthetic_logs` on the inputs of
Examples/Example_01 vs `Examples/Example_01/Output/Results_2024_08_17__18_59_29/Results_1.txt` (tests/golden/example_01).

The reference solved 2D axisymmetric Netgen meshes at order 3.  Two runs:
  * the same 2D axisymmetric path here (conforming 2D mesher, order 3): agreement at the reference's mesh-noise level;
  * the 3D half-ball path with a tiny dip (order 2; the meshes carry the borehole wall, the layer planes and the invasion
    cylinders, `mesh_options["conforming"]` = the Model default): <= 6e-3 on the CPU oracle with the same meshes, bound 1e-2
    (round 1, materials per tet centroid: 1.8 %, bound 4e-2).  Depths inside thick beds and next to a bed boundary."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 0.01


def test_example01_2d_path_matches_reference_output(golden_dir):
    from remo3d_b200 import Model

    d = os.path.join(golden_dir, "example_01")
    gold = np.loadtxt(os.path.join(d, "Results_1.txt"), skiprows=2)
    names = open(os.path.join(d, "Results_1.txt")).readline().split()[1:]
    tools = ["A2.0M0.5N", "N0.5M2.0A", "M1.0A0.1B", "B5.7A0.4M"]
    depths = np.arange(4.0, 8.01, 0.5)  # crosses the 3.05-8.35 m invaded bed
    model = Model.compute_synthetic_logs(tools, depths, os.path.join(d, "Formation.txt"), os.path.join(d, "Borehole.txt"),
                                         cpu_workers=4, gpu_workers=1)  # defaults: dip 0 -> 2D, order 3, multigrid
    worst = 0.0
    for t in tools:
        col = names.index(t) + 1
        ref = np.array([gold[np.argmin(np.abs(gold[:, 0] - z)), col] for z in depths])
        rel = np.abs(model.logs[t][:, 1] - ref) / ref
        print(t, "rel", np.round(rel, 5))
        worst = max(worst, rel.max())
    assert all(r is not None and "error" not in r for r in model.task_records), model.task_records
    assert worst < 5e-3, worst


def test_example01_logs_match_reference_output(golden_dir):
    from remo3d_b200 import Model

    d = os.path.join(golden_dir, "example_01")
    gold = np.loadtxt(os.path.join(d, "Results_1.txt"), skiprows=2)
    names = open(os.path.join(d, "Results_1.txt")).readline().split()[1:]
    tools = ["A2.0M0.5N", "N0.5M2.0A", "M1.0A0.1B"]
    depths = np.array([5.5, 6.0, 8.3, 15.0, 15.5])
    model = Model.compute_synthetic_logs(tools, depths, os.path.join(d, "Formation.txt"), os.path.join(d, "Borehole.txt"),
                                         dip=0.01, cpu_workers=4, gpu_workers=1, order=2,
                                         mesh_options={"h_electrode": 0.012, "h_axis": 0.035, "grading": 0.28})
    worst = 0.0
    for t in tools:
        col = names.index(t) + 1
        ref = np.array([gold[np.argmin(np.abs(gold[:, 0] - z)), col] for z in depths])
        got = model.logs[t][:, 1]
        np.testing.assert_array_equal(model.logs[t][:, 0], depths)
        rel = np.abs(got - ref) / ref
        print(t, "reference", ref, "this repo", np.round(got, 4), "rel", np.round(rel, 4))
        worst = max(worst, rel.max())
    assert all(r is not None and "error" not in r for r in model.task_records), model.task_records
    assert worst < TOL, worst


def test_example01_3d_path_at_the_reference_order(golden_dir):
    """The same log through the 3D path at ORDER 3 (what the reference hard-wires, `ngsolve_functions.py:27`; the `Model`
    default): the order-3 element-wise product + the mixed-precision V-cycle end to end, on the meshes of the order-2 test
    (~29 k vertices, ~780 k dofs per mesh at order 3).  The interface representation of the mesh, not the polynomial order,
    carries the difference to the reference (a coarser near field, 17 k vertices, is 1.5 % off at depth 15.5 at either order)."""
    from remo3d_b200 import Model

    d = os.path.join(golden_dir, "example_01")
    gold = np.loadtxt(os.path.join(d, "Results_1.txt"), skiprows=2)
    names = open(os.path.join(d, "Results_1.txt")).readline().split()[1:]
    tools = ["A2.0M0.5N", "N0.5M2.0A"]
    depths = np.array([5.5, 15.5])
    model = Model.compute_synthetic_logs(tools, depths, os.path.join(d, "Formation.txt"), os.path.join(d, "Borehole.txt"),
                                         dip=0.01, cpu_workers=4, gpu_workers=1,
                                         mesh_options={"h_electrode": 0.012, "h_axis": 0.035, "grading": 0.28})
    assert all(r is not None and "error" not in r for r in model.task_records), model.task_records
    worst = 0.0
    for t in tools:
        ref = np.array([gold[np.argmin(np.abs(gold[:, 0] - z)), names.index(t) + 1] for z in depths])
        rel = np.abs(model.logs[t][:, 1] - ref) / ref
        print(t, "reference", ref, "this repo", np.round(model.logs[t][:, 1], 4), "rel", np.round(rel, 4))
        worst = max(worst, rel.max())
    assert worst < TOL, worst
