"""The product (`Model.compute_synthetic_logs` on the B200 path) against EVERY full-pipeline output the reference commits:

  * Examples/Example_01/Output/.../Results_1.txt            251 depths x 6 tools, R = 50, batch 5   (tests/golden/example_01)
  * Examples/Example_02/Output/.../Results_1.txt            same model, batch 10 (script: R = 25, see below)
  * Examples/Benchmark models/Thin-bedded model/Logs/Logs {1-4}/Results_1.txt   81 depths x 4 tools each (tests/golden/thin_bedded)

The reference produced them with 2D axisymmetric Netgen meshes, order-3 H1 and NGSolve's PCG; Netgen meshes are not
reproducible, and this repo meshes with its own conforming 2D mesher, so agreement is at discretisation level.  Measured
(CPU oracle study and B200, profiles/r02_reference_logs.json, written by these tests when REMO_GOLDEN_STATS is set):
Example_01 <= 1e-3 on every one of the 1506 values once the far field of the mesh is fine enough (`meshgen2d` h_max =
min(R / 25, 2 m); with the former R / 8 the long-spacing tool M4.0A0.5B was 2 % off at the top of the log).

Example_02: the committed output equals Example_01's to the 4 printed digits (median relative difference 2e-5, max
3.1e-4), which a Dirichlet sphere at R = 25 cannot give: here R = 25 moves the long-spacing tools by 0.3-0.7 % at any
mesh resolution (and R = 50 / 100 agree to 1e-4).  The file is therefore compared at R = 50, batch 10 with the tight
bound, and at the script's R = 25 with a bound that covers the truncation effect.

Thin-bedded set: which formation file belongs to which log set is not recorded (the README's "first / second" is the
opposite of the file numbering); Formation_model_1 reproduces Logs 1 / 3 and Formation_model_2 Logs 2 / 4 (the other pairing
is off by 4-15 %).  The generating script is not in the repository; the reference's default R = 50 is used (R = 100 gives
the same values to 1e-3).  The difference to these logs is SYSTEMATIC and grows with the spacing of the tool, whatever the
mesh sizes and the radius (the CPU oracle with a direct solver gives the same numbers): medians 5e-4 (A0.4M6.0N), 2e-3
(A1.62M6.0N), 3.5e-3 (A4.0M0.5N) and +2.4 % for A8.0M1.0N (K = 905: a 0.5 % difference in the potentials 8-9 m from the
source is a 5 % difference in Ra) -- while the same pipeline reproduces Example_01 / Example_02 to 1e-3 for every tool,
including a 4 m dipole: a property of the unknown generating configuration of these logs (code version, mesher, radius),
bounded per tool below."""
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


@pytest.mark.parametrize("which", ["example_01", "example_02", "example_02_r25"])
def test_examples_full_table(golden_dir, which):
    d = os.path.join(golden_dir, "example_01")
    names, gold = _gold(os.path.join(d, "Results_1.txt" if which == "example_01" else "Results_1_example02_R25_batch10.txt"))
    assert names == SIX and gold.shape == (251, 7)
    depths = np.arange(0, 25.1, 0.1)
    kw = {"example_01": {}, "example_02": {"mesh_generator": "netgen", "batch_size": 10},
          "example_02_r25": {"mesh_generator": "netgen", "domain_radius": 25, "batch_size": 10}}[which]  # Example_02.py:20-21
    model = _run(SIX, depths, os.path.join(d, "Formation.txt"), os.path.join(d, "Borehole.txt"), **kw)
    rel = {t: np.abs(model.logs[t][:, 1] - gold[:, k + 1]) / gold[:, k + 1] for k, t in enumerate(SIX)}
    for t in SIX:
        np.testing.assert_allclose(model.logs[t][:, 0], gold[:, 0], atol=1e-9)
    st = _stats(which, rel, {"iters_max": int(max(max(r["iters"]) for r in model.task_records)), "tasks": len(model.task_records),
                              "noconv_tasks": sum(1 for r in model.task_records if "noconv" in r)})
    tight = which != "example_02_r25"
    assert max(s["max"] for s in st.values()) < (2.5e-3 if tight else 1.2e-2), st      # worst of the 1506 log points
    assert max(s["p95"] for s in st.values()) < (1.5e-3 if tight else 8e-3), st
    assert max(s["median"] for s in st.values()) < (5e-4 if tight else 3e-3), st


@pytest.mark.parametrize("logs", [1, 2, 3, 4])
def test_thin_bedded_benchmark_logs(golden_dir, logs):
    d = os.path.join(golden_dir, "thin_bedded")
    names, gold = _gold(os.path.join(d, "Logs_%d_Results_1.txt" % logs))
    assert names[:4] == THIN and gold.shape[0] == 81
    shifts = np.loadtxt(os.path.join(d, "Logs_depth_shifts.txt"), skiprows=2)
    np.testing.assert_allclose(shifts[:, 0], gold[:, 0], atol=1e-9)
    depths = shifts[:, 0] if logs in (1, 2) else shifts[:, 1]  # Logs 3 / 4: measured at the shifted depths, filed under DEPT
    formation = os.path.join(d, "Formation_model_%d.txt" % (1 if logs in (1, 3) else 2))
    model = _run(THIN, depths, formation, os.path.join(d, "Borehole_model_correct_rm.txt"))
    rel = {t: np.abs(model.logs[t][:, 1] - gold[:, k + 1]) / gold[:, k + 1] for k, t in enumerate(THIN)}
    st = _stats("thin_bedded_logs_%d" % logs, rel, {"tasks": len(model.task_records), "iters_max": int(max(max(r["iters"]) for r in model.task_records)),
                                                     "noconv_tasks": sum(1 for r in model.task_records if "noconv" in r)})
    for t, (mx, med) in zip(THIN, [(3.5e-3, 1e-3), (1.2e-2, 3.5e-3), (2.2e-2, 6e-3), (6e-2, 3.5e-2)]):
        assert st[t]["max"] < mx and st[t]["median"] < med, (t, st[t])
