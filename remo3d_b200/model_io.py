"""Formation / borehole model loading, unit conversion and sanity checks (host side).

Restates `/root/reference/remo3d/remo3d.py:344-548` (`load_/set_formation_parameters`,
`load_/set_borehole_parameters`, `set_dip`, `_check_model_geometry`) and `:694-720`
(`_add_points_to_borehole`).  File formats: tab separated, two header lines, the
second one holding the units; formation columns `TOP BOTTOM FZ_RADIUS FZ_VALUE UZ_VALUE`,
borehole columns `DEPT CALx RM`.  Error messages are kept verbatim (drop-in contract).
"""
import numpy as np

# remo3d.py:26
CONVERSION_TABLE = {"M": 1.0, "DM": 0.1, "CM": 0.01, "MM": 0.001, "IN": 0.0254, "FT": 0.3048}


def _second_line_tokens(path):
    with open(path) as f:
        f.readline()
        return f.readline().split()


def _convert(columns, units, what):
    for i, unit in enumerate(units):
        if unit not in CONVERSION_TABLE:
            raise ValueError("{} unit in {} model file not recognized. Allowed units: M, DM, CM, MM, IN, FT".format(unit, what))
        columns[:, i] *= CONVERSION_TABLE[unit]


def set_formation_parameters(formation, units=("M", "M", "M")):
    """remo3d.py:406-437.  Converts the geometry columns in place and validates."""
    _convert(formation, list(units), "formation")
    tops_bottoms = formation[:, :2]
    if (np.diff(tops_bottoms, axis=0) <= 0.0).any() or (formation[1:, 0] != formation[:-1, 1]).any():
        raise ValueError("Uncorrect formation model geometry")
    if np.nanmin(formation[:, [3, 4]]) <= 0.0:
        raise ValueError("Formation resistivies have to be higher than 0 ohmm")
    return formation


def load_formation_parameters(path):
    """remo3d.py:380-403: units are all but the last two tokens of header line 2."""
    data = np.atleast_2d(np.loadtxt(path, delimiter="\t", skiprows=2))
    return set_formation_parameters(data, _second_line_tokens(path)[:-2])


def set_borehole_parameters(borehole, geometry_type="diameter", units=("M", "M")):
    """remo3d.py:470-514."""
    if np.shape(borehole)[0] < 2:
        raise ValueError("Borehole paramaters have to be defined for at least two depths")
    _convert(borehole, list(units), "borehole")
    if (np.diff(borehole[:, 0], axis=0) <= 0.0).any() or (borehole[:, 1] <= 0.0).any():
        raise ValueError("Uncorrect borehole model geometry")
    if geometry_type == "diameter":
        borehole[:, 1] /= 2
    elif geometry_type != "radius":
        raise ValueError("Uncorrect borehole geometry type - use 'diameter' or 'radius' to specify borehole geometry")
    if np.nanmin(borehole[:, 2]) <= 0.0:
        raise ValueError("Drilling mud resistivies have to be higher than 0 ohmm")
    return borehole


def load_borehole_parameters(path, geometry_type="diameter"):
    """remo3d.py:440-467: units are all but the last token of header line 2."""
    data = np.atleast_2d(np.loadtxt(path, delimiter="\t", skiprows=2))
    return set_borehole_parameters(data, geometry_type, _second_line_tokens(path)[:-1])


def set_dip(dip):
    """remo3d.py:517-536 -> (degrees, radians)."""
    if dip < 0 or dip >= 90:
        raise ValueError("Uncorrect dip angle")
    return dip, dip * np.pi / 180


def check_model_geometry(formation, borehole):
    """remo3d.py:538-548: the borehole must stay inside every flushed zone it crosses."""
    for top, bottom, fz_radius in formation[:, :3]:
        inside = (borehole[:, 0] >= top) & (borehole[:, 0] <= bottom)
        if np.any(borehole[inside, 1] >= fz_radius):
            raise ValueError("Borehole radius have to be smaller than the extend of the filtration zone")


def densify_borehole(borehole, maximal_distance=0.15):
    """remo3d.py:694-720 (`_add_points_to_borehole`): insert interpolated caliper rows where the
    samples are further apart than `maximal_distance` (needed by 3D meshing).  Unlike the reference
    (which leaves its return value unbound when nothing was added) the unchanged
    model is returned in that case."""
    depths = [borehole[0, 0]]
    for i in range(1, borehole.shape[0]):
        gap = borehole[i, 0] - borehole[i - 1, 0]
        if gap > maximal_distance:
            extra = np.linspace(borehole[i - 1, 0], borehole[i, 0], np.max([3, int(gap * 10 + 1)]))
            depths.extend(extra[1:])
        else:
            depths.append(borehole[i, 0])
    depths = np.asarray(depths, dtype=float)
    if depths.shape[0] <= borehole.shape[0]:
        return borehole
    return np.vstack([depths,
                      np.interp(depths, borehole[:, 0], borehole[:, 1]),
                      np.interp(depths, borehole[:, 0], borehole[:, 2])]).T
