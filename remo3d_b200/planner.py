"""Measurement depths x tools -> mesh tasks (which right-hand sides share a mesh).

Host-side restatement of `Model._prepare_simulation_depths_and_tasks`
(`/root/reference/remo3d/remo3d.py:602-692`).  The nesting of the returned task list is the
one `workers/worker.py:75-80, 104-134` unpacks:

    task            = [batch_index, batch_electrodes (2 x K), [source_task, ...]]
    source_task     = [simulation_depth_index, electrodes (2 x k), [log_point, ...]]
    log_point       = [measurement_depth_index, tool_index, offset]
    electrodes[0]   = z relative to the mesh centre (ascending), electrodes[1] = 1 current / 0 potential
                      (in two-current-electrode mode: the tool's own +1/-1/0 source terms)

One task = one mesh; every source_task is one right-hand side on that mesh; every log_point is
one entry of the result array.  `flatten_task` converts a task into the flat arrays the
C-ABI takes.
"""
import numpy as np


def _electrode_table(potential, current):
    """Union of electrode depths -> 2 x n array [z; flag], current wins over potential, z ascending
    (remo3d.py:661-666, 681-688)."""
    cur = np.unique(current)
    pot = np.unique(potential)
    pot = pot[~np.isin(pot, cur)]
    table = np.hstack([np.vstack([pot, np.zeros_like(pot)]), np.vstack([cur, np.ones_like(cur)])])
    return table[:, table[0, :].argsort()]


def prepare_simulation_depths_and_tasks(tools, sec, measurement_depths, batch_size):
    """(tools dict, single-electrode mode, depths, batch size) -> (mesh-centre depths, tasks)."""
    names = list(tools.keys())
    n_meas = len(measurement_depths)
    # absolute depth of each tool's current electrode at every measurement depth (remo3d.py:606-608)
    source_depths = {t: np.round(measurement_depths + tools[t][1, 3], decimals=4) for t in names}

    if sec:
        sim = np.unique(np.hstack([source_depths[t] for t in names]))
        sim_tool = None
    else:
        sim = np.hstack([source_depths[t] for t in names])
        owner = [ti for ti in range(len(names)) for _ in range(n_meas)]
        order = np.argsort(sim)
        sim = sim[order]
        sim_tool = [owner[i] for i in order]

    n_batches = int(np.ceil(sim.size / batch_size))
    grid = np.pad(sim.astype(float), (0, n_batches * batch_size - sim.size), mode="constant",
                  constant_values=np.nan).reshape(n_batches, batch_size)
    centres = np.round(np.nanmean(grid, axis=1), decimals=4)
    offsets = np.round(grid - centres[:, None], decimals=4)

    def shifted(tool, offset):
        e = tools[tool][:, :3].copy()
        e[0, :3] += offset
        return np.round(e, 4)

    tasks = []
    for b in range(n_batches):
        batch_pot, batch_cur, source_tasks = [], [], []
        for j in range(batch_size):
            depth = grid[b, j]
            if np.isnan(depth):
                break
            sim_index = b * batch_size + j
            off = offsets[b, j]
            points = []
            if sec:
                pot, cur = [], []
                for ti, t in enumerate(names):
                    if not np.any(np.isclose(source_depths[t], depth)):
                        continue
                    mi = np.argwhere(np.isclose(measurement_depths + tools[t][1, 3], depth))[0][0]
                    points.append([mi, ti, off])
                    e = shifted(t, off)
                    c = list(e[0, e[1, :] != 0])
                    p = list(e[0, e[1, :] == 0])
                    cur += c
                    pot += p
                    batch_cur += c
                    batch_pot += p
                electrodes = _electrode_table(pot, cur)
            else:
                ti = sim_tool[sim_index]
                t = names[ti]
                mi = np.argwhere(np.isclose(measurement_depths + tools[t][1, 3], depth))[0][0]
                points.append([mi, ti, off])
                e = shifted(t, off)
                batch_cur += list(e[0, e[1, :] != 0])
                batch_pot += list(e[0, e[1, :] == 0])
                electrodes = e[:, e[0, :].argsort()]
            source_tasks.append([sim_index, electrodes, points])
        tasks.append([b, _electrode_table(batch_pot, batch_cur), source_tasks])
    return centres, tasks


def flatten_task(task, tools, three_d):
    """One task -> flat arrays for the batched solve + Ra kernel.

    Returns dict with
      src_ptr (nrhs+1), src_z, src_fac : point sources of every right-hand side
                                         (`ngsolve_functions.py:41-44`: only non-zero source terms)
      pt_rhs, pt_z0, pt_z1, pt_k, pt_depth, pt_tool : one row per log point; z1 = NaN when the tool
                                         has a single potential electrode (`worker.py:113-131`)
      scale : 0.5 on the 3D half-ball, 1.0 in 2D (`worker.py:129-131`)
    """
    names = list(tools.keys())
    src_ptr, src_z, src_fac = [0], [], []
    pt = {k: [] for k in ("rhs", "z0", "z1", "k", "depth", "tool")}
    for r, (_, electrodes, points) in enumerate(task[2]):
        z, s = electrodes[0, :], electrodes[1, :]
        for zz, ss in zip(z, s):
            if ss != 0.0:
                src_z.append(float(zz))
                src_fac.append(float(ss))
        src_ptr.append(len(src_z))
        for depth_idx, tool_idx, offset in points:
            p = tools[names[tool_idx]]
            geom = p[0, :3] + offset
            meas = geom[p[1, :3] == 0]
            pt["rhs"].append(r)
            pt["z0"].append(meas[0])
            pt["z1"].append(meas[1] if meas.shape[0] == 2 else np.nan)
            pt["k"].append(p[0, 3])
            pt["depth"].append(int(depth_idx))
            pt["tool"].append(int(tool_idx))
    return {
        "src_ptr": np.asarray(src_ptr, dtype=np.int64),
        "src_z": np.asarray(src_z, dtype=np.float64),
        "src_fac": np.asarray(src_fac, dtype=np.float64),
        "pt_rhs": np.asarray(pt["rhs"], dtype=np.int32),
        "pt_z0": np.asarray(pt["z0"], dtype=np.float64),
        "pt_z1": np.asarray(pt["z1"], dtype=np.float64),
        "pt_k": np.asarray(pt["k"], dtype=np.float64),
        "pt_depth": np.asarray(pt["depth"], dtype=np.int64),
        "pt_tool": np.asarray(pt["tool"], dtype=np.int64),
        "scale": 0.5 if three_d else 1.0,
    }
