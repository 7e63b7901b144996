"""Generate golden vectors for the HOST logic of the hot path by importing the real
reference (`/root/reference/remo3d/remo3d.py`) in the build container.

The reference imports mpi4py and matplotlib at module scope (remo3d.py:3-10); neither is
installed here and neither is touched by the functions exercised below, so empty stub
modules are registered first.  Nothing else of the reference is modified.

Run (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_host_golden.py
Writes tests/golden/host_golden.json .
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/remo3d"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden.json")


def _stub(name, attrs=()):
    m = types.ModuleType(name)
    for a in attrs:
        setattr(m, a, type(a, (), {}))
    sys.modules[name] = m
    return m


def load_reference_model():
    mpi = _stub("mpi4py")
    mpi.MPI = types.SimpleNamespace()
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("matplotlib.patches", ["Polygon"])
    _stub("matplotlib.lines", ["Line2D"])
    _stub("matplotlib.collections", ["PatchCollection"])
    sys.modules["matplotlib"].ticker = _stub("matplotlib.ticker")
    sys.path.insert(0, REF)
    import remo3d as ref  # the reference module file remo3d/remo3d.py

    return ref.Model


def tolist(x):
    if isinstance(x, np.ndarray):
        return [tolist(v) for v in x.tolist()] if x.ndim else x.item()
    if isinstance(x, (list, tuple)):
        return [tolist(v) for v in x]
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    if isinstance(x, float) and x != x:
        return "nan"
    return x


def nan_to_str(o):
    if isinstance(o, list):
        return [nan_to_str(v) for v in o]
    if isinstance(o, float) and o != o:
        return "nan"
    return o


def main():
    Model = load_reference_model()
    gold = {"tools": {}, "tool_errors": {}, "planner": [], "loaders": {}}

    tool_sets = [
        (["B5.7A0.4M", "B4.48A1.62M", "M1.0A0.1B", "A2.0M0.5N", "N0.5M2.0A", "M4.0A0.5B"], True),
        (["B5.7A0.4M", "B4.48A1.62M", "M1.0A0.1B", "A2.0M0.5N", "N0.5M2.0A", "M4.0A0.5B"], False),
        (["N2.5M0.25A", "A0.5M0.25N", "M0.3N2.0B", "A0.2B3.0M", "N1.0A0.5B"], True),
        (["N2.5M0.25A", "A0.5M0.25N", "M0.3N2.0B", "A0.2B3.0M", "N1.0A0.5B"], False),
    ]
    for tools, fsec in tool_sets:
        m = Model(tools, force_single_electrode_configuration=fsec)
        key = "%s|%s" % (",".join(tools), fsec)
        gold["tools"][key] = {"sec": bool(m.sec), "params": {t: tolist(m.tools[t]) for t in tools}}

    for bad in (["A1.0M1.0N"], ["A1.0M"], ["A1.0A2.0M"], ["X1.0M2.0N"], ["A-1.0M2.0N"], ["A1.0M2.0N0.5B"]):
        try:
            Model(bad)
            gold["tool_errors"][bad[0]] = None
        except ValueError as e:
            gold["tool_errors"][bad[0]] = str(e)
    for bad_arg in ("A1.0M2.0N", [1.0]):
        try:
            Model(bad_arg)
        except ValueError as e:
            gold["tool_errors"][repr(bad_arg)] = str(e)

    plans = [
        (tool_sets[0][0], True, np.arange(0, 25.1, 0.1), 5),
        (tool_sets[0][0], True, np.arange(0, 25.1, 0.1), 10),
        (tool_sets[0][0], False, np.arange(0, 3.0, 0.1), 5),
        (["B5.7A0.4M"], True, np.arange(0, 10, 0.1), 5),
        (["A2.0M0.5N", "N0.5M2.0A", "B5.7A0.4M", "M4.0A0.5B"], True, np.arange(10, 12, 0.25), 5),
        (["A2.0M0.5N", "N0.5M2.0A"], True, np.arange(3, 4, 0.1), 1),
        (["A0.2B3.0M", "N1.0A0.5B"], False, np.arange(3, 4, 0.1), 4),
    ]
    for tools, fsec, depths, bs in plans:
        m = Model(tools, force_single_electrode_configuration=fsec)
        cd, tasks = m._prepare_simulation_depths_and_tasks(depths, bs)
        gold["planner"].append({
            "tools": tools, "fsec": fsec, "depths": tolist(depths), "batch_size": bs,
            "combined_depths": nan_to_str(tolist(cd)),
            "tasks": nan_to_str(tolist(tasks)),
        })

    ex = "/root/reference/Examples"
    files = {
        "ex01": (ex + "/Example_01/Input/Formation.txt", ex + "/Example_01/Input/Borehole.txt"),
        "bm2": (ex + "/Benchmark models/Benchmark model 2/Formation_BM2.txt",
                ex + "/Benchmark models/Benchmark model 2/Borehole_BM2.txt"),
    }
    for k, (ff, bf) in files.items():
        m = Model(["A2.0M0.5N"])
        m.set_model_parameters(ff, bf)
        gold["loaders"][k] = {
            "formation_text": open(ff).read(), "borehole_text": open(bf).read(),
            "formation": nan_to_str(tolist(m.formation_model)),
            "borehole": nan_to_str(tolist(m.borehole_model)),
        }
    # error behaviour of the setters (messages are part of the drop-in contract)
    errs = {}
    m = Model(["A2.0M0.5N"])
    cases = {
        "formation_gap": lambda: m.set_formation_parameters(np.array([[0., 1, np.nan, np.nan, 5], [1.5, 2, np.nan, np.nan, 5]])),
        "formation_neg_res": lambda: m.set_formation_parameters(np.array([[0., 1, np.nan, np.nan, -5]])),
        "formation_unit": lambda: m.set_formation_parameters(np.array([[0., 1, np.nan, np.nan, 5]]), ["M", "KM", "M"]),
        "borehole_one_row": lambda: m.set_borehole_parameters(np.array([[0., 0.2, 1.0]])),
        "borehole_neg": lambda: m.set_borehole_parameters(np.array([[0., -0.2, 1.0], [1.0, 0.2, 1.0]])),
        "borehole_type": lambda: m.set_borehole_parameters(np.array([[0., 0.2, 1.0], [1.0, 0.2, 1.0]]), "circumference"),
        "borehole_rm": lambda: m.set_borehole_parameters(np.array([[0., 0.2, 0.0], [1.0, 0.2, 1.0]])),
        "borehole_unit": lambda: m.set_borehole_parameters(np.array([[0., 0.2, 1.0], [1.0, 0.2, 1.0]]), borehole_units=["M", "YD"]),
        "dip_90": lambda: m.set_dip(90),
        "dip_neg": lambda: m.set_dip(-1),
    }
    for k, fn in cases.items():
        try:
            fn()
            errs[k] = None
        except ValueError as e:
            errs[k] = str(e)
    gold["setter_errors"] = errs
    gold["dip_30"] = tolist(list(m.set_dip(30)))

    with open(OUT, "w") as f:
        json.dump(gold, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
