"""The product (`Model.compute_synthetic_logs` on the B200 path) against EVERY full-pipeline output the reference commits:

  * Examples/Example_01/Output/.../Results_1.txt            251 depths x 6 tools, R = 50, batch 5   (tests/golden/example_01)
  * Examples/Example_02/Output/.../Results_1.txt            same model, R = 25, batch 10
  * Examples/Benchmark models/Thin-bedded model/Logs/Logs {1-4}/Results_1.txt   81 depths x 4 tools each (tests/golden/thin_bedded)

The reference produced them with 2D axisymmetric Netgen meshes, order-3 H1 and NGSolve's PCG; Netgen meshes are not
reproducible (the reference's own Example_01 and Example_02 logs differ by up to 3.1e-4), and this repo meshes with its own
conforming 2D mesher, so agreement is at discretisation level, stated per file below and measured in
profiles/r02_reference_logs.json (written by these tests when REMO_GOLDEN_STATS is set).

Thin-bedded set: which formation file belongs to which log set is not recorded (the README's "first / second" is the
opposite of the file numbering); Formation_model_1 reproduces Logs 1 / 3 and Formation_model_2 Logs 2 / 4 (the other pairing
is off by 4-15 %).  The domain radius of those runs is not recorded either: the long lateral tool A8.0M1.0N changes by 10 %
between R = 25 and R = 100 and is converged from R = 100 on, where the other three tools agree with the reference to
< 0.5 %; R = 100 is used.  A8.0M1.0N (K = 905: a 0.5 % difference in the potentials 8-9 m from the source is a 5 %
difference in Ra) stays 2-5 % above the reference however fine the mesh (CPU oracle study, same numbers): its bound is its own."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIX = ["B5.7A0.4M", "B4.48A1.62M", "M1.0A0.1B", "A2.0M0.5N", "N0.5M2.0A", "M4.0A0.5B"]  # Example_01.py / Example_02.py
THIN = ["A0.4M6.0N", "A1.62M6.0N", "A4.0M0.5N", "A8.0M1.0N"]


def _gold(path):
    names = open(path).readline().split()[1:]
    return names, np.loadtxt(path, skiprows=2)


def _stats(name, rel_by_tool, extra=None):
    out = {t: {"max": float(np.nanmax(r)), "p95": float(np.nanpercentile(r, 95)), "median": float(np.nanmedian(r)), "n": int(r.shape[0])}
           for t, r in rel_by_tool.items()}
    print(name, json.dumps(out))
    dst = os.environ.get("REMO_GOLDEN_STATS")
    if dst:
        try:
            allst = json.load(open(dst))
        except Exception:
            allst = {}
        allst[name] = dict(out, **(extra or {}))
        with open(dst, "w") as f:
            json.dump(allst, f, indent=1, sort_keys=True)
    return out


def _run(tools, depths, formation, borehole, **kw):
    from remo3d_b200 import Model

    model = Model.compute_synthetic_logs(tools, depths, formation, borehole, cpu_workers=os.cpu_count() or 4, gpu_workers=1, **kw)
    bad = [r for r in model.task_records if r is None or "error" in r]
    assert not bad, bad[:2]
    return model


@pytest.mark.parametrize("which", ["example_01", "example_02"])
def test_examples_full_table(golden_dir, which):
    d = os.path.join(golden_dir, "example_01")
    names, gold = _gold(os.path.join(d, "Results_1.txt" if which == "example_01" else "Results_1_example02_R25_batch10.txt"))
    assert names == SIX and gold.shape == (251, 7)
    depths = np.arange(0, 25.1, 0.1)
    kw = {} if which == "example_01" else {"mesh_generator": "netgen", "domain_radius": 25, "batch_size": 10}  # Example_02.py:20-21
    model = _run(SIX, depths, os.path.join(d, "Formation.txt"), os.path.join(d, "Borehole.txt"), **kw)
    rel = {t: np.abs(model.logs[t][:, 1] - gold[:, k + 1]) / gold[:, k + 1] for k, t in enumerate(SIX)}
    for t in SIX:
        np.testing.assert_allclose(model.logs[t][:, 0], gold[:, 0], atol=1e-9)
    st = _stats(which, rel, {"iters_max": int(max(max(r["iters"]) for r in model.task_records)), "tasks": len(model.task_records)})
    assert max(s["max"] for s in st.values()) < 1.5e-2, st      # worst single point (next to a bed boundary)
    assert max(s["p95"] for s in st.values()) < 4e-3, st        # 95 % of the 1506 log points
    assert max(s["median"] for s in st.values()) < 1.5e-3, st


@pytest.mark.parametrize("logs", [1, 2, 3, 4])
def test_thin_bedded_benchmark_logs(golden_dir, logs):
    d = os.path.join(golden_dir, "thin_bedded")
    names, gold = _gold(os.path.join(d, "Logs_%d_Results_1.txt" % logs))
    assert names[:4] == THIN and gold.shape[0] == 81
    shifts = np.loadtxt(os.path.join(d, "Logs_depth_shifts.txt"), skiprows=2)
    np.testing.assert_allclose(shifts[:, 0], gold[:, 0], atol=1e-9)
    depths = shifts[:, 0] if logs in (1, 2) else shifts[:, 1]  # Logs 3 / 4: measured at the shifted depths, filed under DEPT
    formation = os.path.join(d, "Formation_model_%d.txt" % (1 if logs in (1, 3) else 2))
    model = _run(THIN, depths, formation, os.path.join(d, "Borehole_model_correct_rm.txt"), domain_radius=100)
    rel = {t: np.abs(model.logs[t][:, 1] - gold[:, k + 1]) / gold[:, k + 1] for k, t in enumerate(THIN)}
    st = _stats("thin_bedded_logs_%d" % logs, rel, {"tasks": len(model.task_records)})
    for t in THIN[:3]:
        assert st[t]["max"] < 1.5e-2 and st[t]["median"] < 3e-3, (t, st[t])
    assert st[THIN[3]]["max"] < 8e-2, st[THIN[3]]
