"""Logging-tool strings -> electrode geometry, source terms, geometric factor.

Host-side restatement of the reference's tool parser
(`/root/reference/remo3d/remo3d.py:178-340`, `Model.set_tools_parameters`,
`Model._set_tool_parameters`, `Model._str2float`).  Pure NumPy; no GPU involved.

A tool is three electrodes out of {A, B, M, N} listed top to bottom with the two spacings
in metres between them, e.g. ``"N0.5M2.0A"``.  A/B inject current, M/N measure potential.
The result for one tool is the 2x4 array the rest of the reference passes around::

    [[z_1, z_2, z_3, K          ],     z ascending, relative to the current electrode(s)
     [s_1, s_2, s_3, depth_shift]]     s = source term (+1/-1 current, 0 potential)

Error messages are part of the drop-in contract and are kept verbatim.
"""
import itertools
import re

import numpy as np

_TOKEN = re.compile(r"[A-Za-z]+|[^A-Za-z]+")
_RECIPROCAL = str.maketrans("ABMN", "MNAB")  # remo3d.py:213, current <-> potential swap
_VALID = set(itertools.permutations("ABMN", 3))


def _tokens(name):
    """Split into alphabetic / non-alphabetic runs; numeric runs become floats (remo3d.py:214-218, 323-340)."""
    out = []
    for tok in _TOKEN.findall(name):
        try:
            out.append(float(tok))
        except ValueError:
            out.append(tok)
    return out


def _bad(tool):
    return ValueError("{} logging tool specification is uncorrect".format(tool))


def tool_parameters(tool, electrodes, distances):
    """One tool -> 2x4 parameter array (remo3d.py:231-321)."""
    if len(electrodes) != 3 or len(distances) != 2 or min(distances) <= 0:
        raise _bad(tool)
    if tuple(electrodes) not in _VALID:
        raise _bad(tool)
    d0, d1 = distances
    # measurement point: midpoint of the closer electrode pair (remo3d.py:258-264)
    if d0 < d1:
        z_mp = d0 / 2
    elif d0 > d1:
        z_mp = d0 + d1 / 2
    else:
        raise _bad(tool)
    top_down = np.array([0, 0 + d0, 0 + d0 + d1])
    z = {e: top_down[i] - z_mp for i, e in enumerate(electrodes)}
    missing = (set("ABMN") - set(electrodes)).pop()

    if missing in "AB":
        # one current electrode C, two potential electrodes (remo3d.py:281-294)
        c = "B" if missing == "A" else "A"
        r_m = abs(z[c] - z["M"])
        r_n = abs(z[c] - z["N"])
        k = abs(4 * np.pi * r_m * r_n / (r_n - r_m))
        shift = z[c]
        pos = np.array([z[c], z["M"], z["N"]])
        src = np.array([1, 0, 0])
    else:
        # two current electrodes, one potential electrode P (remo3d.py:295-308)
        p = "N" if missing == "M" else "M"
        r_a = abs(z["A"] - z[p])
        r_b = abs(z["B"] - z[p])
        k = abs(4 * np.pi * r_a * r_b / (r_a - r_b)) if p == "N" else abs(4 * np.pi * r_a * r_b / (r_b - r_a))
        shift = (z["A"] + z["B"]) / 2
        pos = np.array([z["A"], z["B"], z[p]])
        src = np.array([1, -1, 0])

    order = np.argsort(pos)
    out = np.empty((2, 4))
    out[0, :3] = pos[order]
    out[1, :3] = src[order]
    out[0, 3] = k
    out[1, 3] = shift
    out[0, :3] -= shift  # centre on the current electrode(s), remo3d.py:319
    return out


def set_tools_parameters(tools, force_single_electrode_configuration=True):
    """list of tool names -> (dict name -> 2x4 array, single_electrode_mode) (remo3d.py:178-228)."""
    if type(tools) != list or not all(isinstance(s, str) for s in tools):
        raise ValueError("Tools names have to be provided in the form of list of strings")
    if type(force_single_electrode_configuration) != bool:
        raise ValueError("The value of parameter force_single_electrode_configuration can be set only to True or False")

    params = {}
    for tool in tools:
        name = tool
        if force_single_electrode_configuration and "A" in tool and "B" in tool:
            name = tool.translate(_RECIPROCAL)  # reciprocity: ABM -> MNA (remo3d.py:211-214)
        toks = _tokens(name)
        electrodes = tuple(t for t in toks if isinstance(t, str))
        distances = [t for t in toks if isinstance(t, float)]
        params[tool] = tool_parameters(tool, electrodes, distances)

    sec = True
    for p in params.values():
        if np.isclose(np.sum(p[1, :3]), 0):
            sec = False
    return params, sec
