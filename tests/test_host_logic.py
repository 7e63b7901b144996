"""Host logic (tool parser, task planner, model loaders) against golden vectors produced by the
real reference code (tests/golden/make_host_golden.py imports /root/reference/remo3d/remo3d.py)."""
import json
import os

import numpy as np
import pytest

from remo3d_b200 import model_io, planner, tools as tl


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "host_golden.json")) as f:
        return json.load(f)


def _arr(x):
    return np.array([[np.nan if v == "nan" else v for v in row] for row in x], dtype=float)


def test_tool_parameters_bit_exact(gold):
    for key, g in gold["tools"].items():
        names, fsec = key.split("|")
        names = names.split(",")
        params, sec = tl.set_tools_parameters(names, fsec == "True")
        assert sec == g["sec"]
        assert list(params.keys()) == names
        for t in names:
            np.testing.assert_array_equal(params[t], np.array(g["params"][t]), err_msg=t)


def test_tool_table_survey_values():
    # SURVEY.md §3.4 worked values
    p, sec = tl.set_tools_parameters(["B5.7A0.4M", "N0.5M2.0A", "M4.0A0.5B"])
    assert sec
    np.testing.assert_allclose(p["B5.7A0.4M"], [[-6.1, -0.4, 0, 5.37929], [0, 0, 1, 0.2]], atol=1e-5)
    np.testing.assert_allclose(p["N0.5M2.0A"], [[-2.5, -2.0, 0, 125.66371], [0, 0, 1, 2.25]], atol=1e-5)
    np.testing.assert_allclose(p["M4.0A0.5B"], [[0, 4.0, 4.5, 452.38934], [1, 0, 0, -4.25]], atol=1e-5)


def test_tool_errors(gold):
    for name, msg in gold["tool_errors"].items():
        if name.startswith("'") or name.startswith("["):
            arg = eval(name)
        else:
            arg = [name]
        if msg is None:
            tl.set_tools_parameters(arg)
        else:
            with pytest.raises(ValueError) as e:
                tl.set_tools_parameters(arg)
            assert str(e.value) == msg
    with pytest.raises(ValueError):
        tl.set_tools_parameters(["A1.0M2.0N"], force_single_electrode_configuration=1)


def _cmp_nested(a, b, path="task"):
    """Compare our task nesting with the golden (lists / arrays / scalars; 'nan' strings)."""
    if isinstance(b, list):
        a_list = a.tolist() if isinstance(a, np.ndarray) else a
        assert len(a_list) == len(b), path
        for i, (x, y) in enumerate(zip(a_list, b)):
            _cmp_nested(x, y, "%s[%d]" % (path, i))
    elif b == "nan":
        assert a != a, path
    else:
        assert a == b, "%s: %r != %r" % (path, a, b)


def test_planner_bit_exact(gold):
    for g in gold["planner"]:
        params, sec = tl.set_tools_parameters(g["tools"], g["fsec"])
        centres, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.array(g["depths"]), g["batch_size"])
        _cmp_nested(centres, g["combined_depths"], "centres")
        _cmp_nested(tasks, g["tasks"])


def test_planner_example01_counts():
    # 6 tools x 251 depths, sec mode: 1506 log values <- 818 solves <- 164 meshes
    names = ["B5.7A0.4M", "B4.48A1.62M", "M1.0A0.1B", "A2.0M0.5N", "N0.5M2.0A", "M4.0A0.5B"]
    params, sec = tl.set_tools_parameters(names)
    _, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.arange(0, 25.1, 0.1), 5)
    assert len(tasks) == 164
    assert sum(len(t[2]) for t in tasks) == 818
    assert sum(len(s[2]) for t in tasks for s in t[2]) == 1506


def test_flatten_task():
    names = ["A2.0M0.5N", "N0.5M2.0A", "A0.2B3.0M"]
    params, sec = tl.set_tools_parameters(names, False)
    assert not sec
    _, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.arange(3, 4, 0.25), 4)
    seen = 0
    for task in tasks:
        f = planner.flatten_task(task, params, three_d=True)
        nrhs = len(task[2])
        assert f["src_ptr"].shape == (nrhs + 1,)
        assert f["scale"] == 0.5
        for r, st in enumerate(task[2]):
            lo, hi = f["src_ptr"][r], f["src_ptr"][r + 1]
            el = st[1]
            np.testing.assert_array_equal(f["src_z"][lo:hi], el[0, el[1] != 0])
            np.testing.assert_array_equal(f["src_fac"][lo:hi], el[1, el[1] != 0])
        for i in range(f["pt_rhs"].shape[0]):
            p = params[names[f["pt_tool"][i]]]
            n_pot = int(np.sum(p[1, :3] == 0))
            assert np.isnan(f["pt_z1"][i]) == (n_pot == 1)
            seen += 1
    assert seen == 3 * 4


def test_loaders(gold, tmp_path):
    for k, g in gold["loaders"].items():
        ff = tmp_path / (k + "_f.txt")
        bf = tmp_path / (k + "_b.txt")
        ff.write_text(g["formation_text"])
        bf.write_text(g["borehole_text"])
        f = model_io.load_formation_parameters(str(ff))
        b = model_io.load_borehole_parameters(str(bf))
        np.testing.assert_array_equal(f, _arr(g["formation"]))
        np.testing.assert_array_equal(b, _arr(g["borehole"]))
        model_io.check_model_geometry(f, b)


def test_setter_errors(gold):
    nan = np.nan
    cases = {
        "formation_gap": lambda: model_io.set_formation_parameters(np.array([[0., 1, nan, nan, 5], [1.5, 2, nan, nan, 5]])),
        "formation_neg_res": lambda: model_io.set_formation_parameters(np.array([[0., 1, nan, nan, -5]])),
        "formation_unit": lambda: model_io.set_formation_parameters(np.array([[0., 1, nan, nan, 5]]), ["M", "KM", "M"]),
        "borehole_one_row": lambda: model_io.set_borehole_parameters(np.array([[0., 0.2, 1.0]])),
        "borehole_neg": lambda: model_io.set_borehole_parameters(np.array([[0., -0.2, 1.0], [1.0, 0.2, 1.0]])),
        "borehole_type": lambda: model_io.set_borehole_parameters(np.array([[0., 0.2, 1.0], [1.0, 0.2, 1.0]]), "circumference"),
        "borehole_rm": lambda: model_io.set_borehole_parameters(np.array([[0., 0.2, 0.0], [1.0, 0.2, 1.0]])),
        "borehole_unit": lambda: model_io.set_borehole_parameters(np.array([[0., 0.2, 1.0], [1.0, 0.2, 1.0]]), units=["M", "YD"]),
        "dip_90": lambda: model_io.set_dip(90),
        "dip_neg": lambda: model_io.set_dip(-1),
    }
    for k, fn in cases.items():
        msg = gold["setter_errors"][k]
        assert msg is not None
        with pytest.raises(ValueError) as e:
            fn()
        assert str(e.value) == msg, k
    assert list(model_io.set_dip(30)) == gold["dip_30"]


def test_densify_borehole():
    b = np.array([[0.0, 0.1, 1.0], [0.1, 0.1, 1.0], [1.1, 0.2, 2.0]])
    d = model_io.densify_borehole(b)
    assert d.shape[0] > 3 and d[0, 0] == 0.0 and d[-1, 0] == 1.1
    assert np.all(np.diff(d[:, 0]) > 0)
    np.testing.assert_allclose(np.interp(0.6, d[:, 0], d[:, 1]), 0.15)
    same = model_io.densify_borehole(b[:2])
    assert same is b[:2] or np.array_equal(same, b[:2])


def test_save_results_is_byte_compatible_with_the_reference_output(tmp_path):
    """`Results_<n>.txt` (remo3d.py:957-990): feeding the reference's own committed logs back through save_results must
    reproduce its file byte for byte (header, tab separation, %.4f), and logs on a different depth grid go to a second
    file."""
    import glob
    import os

    from remo3d_b200 import Model

    golden = os.path.join(os.path.dirname(__file__), "golden", "example_01", "Results_1.txt")
    lines = open(golden).read().splitlines()
    names = lines[0].split("\t")[1:]
    table = np.loadtxt(golden, skiprows=2)
    model = Model(names)
    model.logs = {n: np.column_stack([table[:, 0], table[:, 1 + i]]) for i, n in enumerate(names)}
    model.logs["EXTRA"] = np.column_stack([table[::2, 0], table[::2, 1]])
    sub = model.save_results(str(tmp_path))
    files = sorted(glob.glob(os.path.join(sub, "Results_*.txt")))
    assert [os.path.basename(f) for f in files] == ["Results_1.txt", "Results_2.txt"]
    assert open(files[0], "rb").read() == open(golden, "rb").read()
    second = open(files[1]).read().splitlines()
    assert second[0] == "DEPTH\tEXTRA" and second[1] == "M\tOHMM" and len(second) == 2 + table[::2].shape[0]
    assert model.save_results(None) is None


def test_sliver_pass_of_the_half_ball_mesher():
    """meshgen.half_ball_mesh(improve=N): the optional quality pass perturbs the free vertices of the worst tets and
    re-triangulates.  The mesh must stay a valid input of the path (positive volumes, the same axis and boundary
    vertices, plane vertices still in the plane, every vertex used) with far fewer slivers; improve=0 must reproduce the default mesh exactly."""
    import numpy as np

    from remo3d_b200 import meshgen
    from remo3d_b200.mesh import Mesh

    ez = np.array([-1.0, 0.0, 0.5])
    kw = dict(h_electrode=0.06, h_axis=0.25, grading=0.5, h_max=6.0, seed=0)
    m0 = meshgen.half_ball_mesh(50.0, ez, **kw)
    m00 = meshgen.half_ball_mesh(50.0, ez, improve=0, **kw)
    np.testing.assert_array_equal(m0["elems"], m00["elems"])
    np.testing.assert_array_equal(m0["points"], m00["points"])
    m1 = meshgen.half_ball_mesh(50.0, ez, improve=3, **kw)
    q0, q1 = meshgen._quality(m0["points"], m0["elems"]), meshgen._quality(m1["points"], m1["elems"])
    assert (q1 < 0.1).sum() * 3 <= (q0 < 0.1).sum()
    x = m1["points"][m1["elems"]]
    assert np.all(np.linalg.det(x[:, 1:] - x[:, :1]) > 0)
    assert m1["points"].shape == m0["points"].shape and np.unique(m1["elems"]).size == m1["points"].shape[0]
    p0, p1 = m0["points"], m1["points"]
    fixed = ((p0[:, 0] == 0.0) & (p0[:, 1] == 0.0)) | (np.linalg.norm(p0, axis=1) >= 50.0 * (1 - 1e-6))
    np.testing.assert_array_equal(p1[fixed], p0[fixed])                 # axis and sphere vertices stay put
    np.testing.assert_array_equal(p1[:, 1] == 0.0, p0[:, 1] == 0.0)     # symmetry-plane vertices stay in the plane
    assert np.all(p1[:, 1] >= 0.0)
    a0 = Mesh(m0["points"], m0["elems"], m0["mat"], m0["bfacets"], m0["bc"], m0["bc_names"]).axis_vertices()
    a1 = Mesh(m1["points"], m1["elems"], m1["mat"], m1["bfacets"], m1["bc"], m1["bc_names"]).axis_vertices()
    np.testing.assert_array_equal(a0, a1)
    for z in ez:  # the electrodes are still mesh vertices
        assert np.any((m1["points"][:, 0] == 0) & (m1["points"][:, 1] == 0) & (np.abs(m1["points"][:, 2] - z) < 1e-12))


def test_3d_mesh_carries_the_material_interfaces():
    """`meshgen.Interfaces` (3D meshes of `Model`, `mesh_options["conforming"]`): lattice points near the borehole wall, a dipping
    layer plane or an invasion cylinder are projected onto it, so almost no tet straddles an interface, the mud column has its
    volume inside the resolved window, and the mesh stays a valid triangulation of the half-ball."""
    from remo3d_b200 import meshgen

    ez = np.array([-1.0, -0.5, 0.0, 0.4])
    rb, rinv, dip = 0.11, 0.4, np.deg2rad(20.0)
    tops = [-0.7, 0.9]
    itf = meshgen.Interfaces(rb, tops, dip, [None, rinv, None])
    material = meshgen.layered_material(tops, dip_rad=dip, borehole_radius=rb, invasion=[None, rinv, None])
    kw = dict(h_electrode=0.03, h_axis=0.08, grading=0.35, seed=0)
    plain = meshgen.half_ball_mesh(20.0, ez, material=material, **kw)
    conf = meshgen.half_ball_mesh(20.0, ez, material=material, interfaces=itf, h_borehole=0.15, r_strip=rinv, g_borehole=0.8,
                                  borehole_window=3.0, g_window=0.15, **kw)

    def stats(m):
        p, e = m["points"], m["elems"]
        x = p[e]
        vol = np.linalg.det(x[:, 1:] - x[:, :1]) / 6.0
        assert (vol > 0).all()
        assert abs(vol.sum() - 2.0 / 3.0 * np.pi * 20.0 ** 3) < 2e-2 * 2.0 / 3.0 * np.pi * 20.0 ** 3
        rho = np.hypot(p[:, 0], p[:, 1])
        near = np.abs(x[:, :, 2]).max(axis=1) < 2.0  # tets inside the resolved part of the column
        side = np.sign(np.round(rho[e] - rb, 9))      # -1 inside the borehole, 0 on the wall, +1 outside
        straddle = near & (side.min(axis=1) < 0) & (side.max(axis=1) > 0)
        touch = near & (np.abs(rho[e] - rb).min(axis=1) < 0.05)
        mud = vol[(m["mat"] == 0) & near].sum()
        return straddle.sum() / max(1, touch.sum()), mud, int((np.abs(rho - rb) < 1e-9).sum())

    f_plain, mud_plain, on_plain = stats(plain)
    f_conf, mud_conf, on_conf = stats(conf)
    exact = 0.5 * np.pi * rb ** 2 * 4.0  # half column, |z| < 2
    print("tets straddling the wall / tets near it: plain %.3f, with interfaces %.3f; mud volume %.4f / %.4f (exact %.4f); wall vertices %d / %d"
          % (f_plain, f_conf, mud_plain, mud_conf, exact, on_plain, on_conf))
    assert on_plain == 0 and on_conf > 200                       # the wall is sampled by mesh vertices
    assert f_conf < 0.25 * f_plain and f_conf < 0.08               # (almost) no tet crosses it any more
    assert abs(mud_conf - exact) < 0.15 * exact                    # and the column has its volume
    # every electrode is still a vertex
    axis = conf["points"][(conf["points"][:, 0] == 0.0) & (conf["points"][:, 1] == 0.0), 2]
    for z in ez:
        assert np.min(np.abs(axis - z)) < 1e-12
