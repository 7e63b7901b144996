"""Host side of the measurement pipeline without a GPU: the per-task failure contract, right-hand-side blocks, shared 3D
geometry, the results log + resume, and `.msh` input -- `Model.simulate_logs` driven by an oracle-backed stand-in for the
GPU context (tests/helpers.OracleContext).  Reference behaviour: `workers/worker.py:74-138`, `remo3d.py:723-884`."""
import json
import multiprocessing
import os

import numpy as np
import pytest

from remo3d_b200 import model_mesh, msh_reader, planner, tools as tl, worker
from remo3d_b200.remo3d import Model
from tests import helpers

FORMATION = np.array([[-50.0, 4.0, np.nan, np.nan, 10.0], [4.0, 6.0, 0.5, 5.0, 100.0], [6.0, 60.0, np.nan, np.nan, 20.0]])
BOREHOLE = np.array([[-50.0, 0.2, 1.0], [60.0, 0.2, 1.2]])
MESH2D = {"h_electrode": 0.05, "h_axis": 0.3, "h_borehole": 0.4, "grading": 0.7}


def _model(tools, ctxs):
    m = Model(tools)
    m.set_model_parameters(FORMATION.copy(), BOREHOLE.copy())
    m.cpu_workers, m.gpu_workers = 2, 1
    m._mesh_pool = multiprocessing.get_context("fork").Pool(2)
    m._contexts = ctxs
    return m


def test_rhs_blocks_split_and_reindex():
    params, sec = tl.set_tools_parameters(["A2.0M0.5N", "N0.5M2.0A"])
    _, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.arange(0, 4.0, 0.1), 70)
    flat = planner.flatten_task(tasks[0], params, three_d=False)
    nrhs = flat["src_ptr"].shape[0] - 1
    assert nrhs == 70
    seen = np.zeros(flat["pt_rhs"].shape[0], bool)
    blocks = list(worker.rhs_blocks(flat))
    assert [b["src_ptr"].shape[0] - 1 for b, _ in blocks] == [32, 32, 6]
    for lo, (b, sel) in zip((0, 32, 64), blocks):
        assert b["src_ptr"][0] == 0 and b["src_ptr"][-1] == b["src_z"].shape[0]
        np.testing.assert_array_equal(b["pt_rhs"] + lo, flat["pt_rhs"][sel])
        n0, n1 = flat["src_ptr"][lo], flat["src_ptr"][lo + b["src_ptr"].shape[0] - 1]
        np.testing.assert_array_equal(b["src_z"], flat["src_z"][n0:n1])
        assert not seen[sel].any()
        seen[sel] = True
    assert seen.all()
    only, sel = next(worker.rhs_blocks(planner.flatten_task(tasks[0], params, False), max_rhs=100))
    assert only["src_ptr"].shape[0] - 1 == 70 and sel.shape[0] == seen.shape[0]


def test_failures_stay_inside_their_task():
    """A mesh that cannot be built, a task that cannot be flattened and a solver error each give NaN for the log points of
    that task only (worker.py:135-138); everything else is solved."""
    params, sec = tl.set_tools_parameters(["A2.0M0.5N"])
    depths = np.arange(3.0, 4.5, 0.1)
    centres, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, depths, 5)
    assert len(tasks) == 3
    jobs = []
    for t in tasks:
        mesh, sigma = model_mesh.build_task_mesh(FORMATION, BOREHOLE[:, :2], 0.0, centres[t[0]], t[1][0], 1.0, 50.0, MESH2D)
        jobs.append([t[0], t, mesh, sigma])
    jobs[1][2] = ValueError("injected mesh failure")
    out = {i: (tr, rec) for i, tr, rec in worker.run_tasks(helpers.OracleContext(fail_on=(2,)), jobs, params, order=1)}
    assert "injected mesh failure" in out[1][1]["error"] and "injected solver failure" in out[2][1]["error"] and "error" not in out[0][1]
    assert all(np.isnan(r) for _, _, r in out[1][0] + out[2][0]) and len(out[1][0]) == 5
    assert all(np.isfinite(r) and r > 0 for _, _, r in out[0][0])
    # depth / tool indices of the NaN points are the task's own
    assert sorted(d for d, _, _ in out[1][0]) == sorted(int(p[0]) for st in tasks[1][2] for p in st[2])


def test_results_log_and_resume(tmp_path):
    tools = ["A2.0M0.5N", "N0.5M2.0A"]
    depths = np.arange(3.0, 4.0, 0.1)
    log = str(tmp_path / "results.jsonl")
    m = _model(tools, [helpers.OracleContext()])
    try:
        m.simulate_logs(depths, mesh_options=MESH2D, order=1, results_log=log)
        full = {t: m.logs[t].copy() for t in tools}
        lines = [json.loads(x) for x in open(log)]
        assert len(lines) == len(m.task_records) and all("error" not in r["record"] for r in lines)
        assert m.pipeline_stats["tasks"] == len(lines) and 0.0 < m.pipeline_stats["gpu_busy_fraction"] <= 1.0
        # crash after two tasks: keep two complete lines and a torn third one
        with open(log, "w") as f:
            f.write(json.dumps(lines[0]) + "\n" + json.dumps(lines[2]) + "\n" + json.dumps(lines[1])[:40])
        ctx = helpers.OracleContext()
        m._contexts = [ctx]
        m.simulate_logs(depths, mesh_options=MESH2D, order=1, results_log=log, resume=True)
        assert ctx.calls == len(lines) - 2  # only the missing tasks were solved
        for t in tools:
            np.testing.assert_array_equal(m.logs[t], full[t])
        # a log written for another plan is ignored
        ctx2 = helpers.OracleContext()
        m._contexts = [ctx2]
        m.simulate_logs(depths[:5], mesh_options=MESH2D, order=1, results_log=log, resume=True)
        assert ctx2.calls == len([r for r in m.task_records])
    finally:
        m._contexts = None
        m.shutdown_workers()


def test_mesh_failure_inside_the_pool_does_not_abort_the_run():
    m = _model(["A2.0M0.5N"], [helpers.OracleContext()])
    try:
        # an option the 2D mesher does not know makes it raise inside the pool for every task: all NaN, no exception here
        m.simulate_logs(np.arange(3.0, 3.6, 0.1), mesh_options={"no_such_option": 1}, order=1)
        assert np.isnan(m.logs["A2.0M0.5N"][:, 1]).all()
        assert all("error" in r for r in m.task_records)
    finally:
        m._contexts = None
        m.shutdown_workers()


def test_shared_geometry_gives_the_same_task_mesh():
    ez = np.array([-2.5, -2.0, -0.4, 0.0, 0.1])
    opts = {"h_electrode": 0.1, "h_axis": 0.4, "grading": 0.6}
    dip = np.deg2rad(20.0)
    key = model_mesh.geometry_key(dip, ez, 30.0, opts)
    assert key == model_mesh.geometry_key(dip, ez + 1e-9, 30.0, dict(opts)) and key != model_mesh.geometry_key(dip, ez + 0.1, 30.0, opts)
    assert model_mesh.geometry_key(0.0, ez, 30.0, opts) is None
    geo = model_mesh.build_geometry_3d(ez, 30.0, opts)
    for depth in (4.5, 5.5):
        a, sa = model_mesh.build_task_mesh(FORMATION, BOREHOLE[:, :2], dip, depth, ez, 1.1, 30.0, opts)
        b, sb = model_mesh.build_task_mesh(FORMATION, BOREHOLE[:, :2], dip, depth, ez, 1.1, 30.0, opts, geometry=geo)
        assert sa == sb
        for name in ("points", "elems", "mat", "bfacets", "bc"):
            np.testing.assert_array_equal(getattr(a, name), getattr(b, name))
    assert len(np.unique(b.mat)) >= 3  # mud, flushed zone, beds all present at this depth


def test_msh_files_as_task_meshes(tmp_path):
    """mesh_generator="gmsh" + mesh_options["msh_path"]: every task's mesh comes from a MSH 2.2 file through msh_reader
    (the reference's ReadGmsh path, worker.py:82-92); same logs as with the meshes built in memory."""
    tools = ["A2.0M0.5N"]
    depths = np.arange(3.0, 3.6, 0.1)
    params, sec = tl.set_tools_parameters(tools)
    centres, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, depths, 5)
    m = _model(tools, [helpers.OracleContext()])
    for t in tasks:
        mesh, _ = model_mesh.build_task_mesh(m.formation_model, m.borehole_model[:, :2], 0.0, centres[t[0]], t[1][0], 1.0, 50.0, MESH2D)
        names = [(1, i + 1, n) for i, n in enumerate(mesh.bc_names)] + [(2, 100 + k, "mat%d" % k) for k in range(mesh.nmat)]
        # Gmsh writes the surfaces physical group by physical group (gmsh_functions.py:592-624), so the reader's material index
        # -- order of first appearance of the elementary tag -- is the index into the sigma list: elements sorted by material
        o = np.argsort(mesh.mat, kind="stable")
        msh_reader.write_msh(str(tmp_path / ("task_%d.msh" % t[0])), mesh.points, mesh.elems[o], [(100 + k, 100 + k) for k in mesh.mat[o]],
                             mesh.bfacets, [(int(b), int(b)) for b in mesh.bc], names)
    try:
        with pytest.raises(ValueError):
            m.simulate_logs(depths, mesh_generator="netgen", mesh_options={"msh_path": "x"}, order=1)
        m.simulate_logs(depths, mesh_options=MESH2D, order=1)
        direct = m.logs[tools[0]].copy()
        m.simulate_logs(depths, mesh_generator="gmsh", mesh_options={"msh_path": str(tmp_path / "task_{}.msh")}, order=1)
        assert all("error" not in r for r in m.task_records), m.task_records
        np.testing.assert_allclose(m.logs[tools[0]][:, 1], direct[:, 1], rtol=1e-9)
    finally:
        m._contexts = None
        m.shutdown_workers()


def test_pattern_aware_shards_partition_the_tasks():
    """`worker.shard_tasks` with the electrode-pattern key: disjoint, complete, balanced to one task, and every rank touches
    few patterns (interleaved shards touch all of them)."""
    rng = np.random.default_rng(0)
    n = 413
    pattern = {i: int(p) for i, p in enumerate(rng.choice(11, size=n, p=np.r_[np.full(8, 0.11), np.full(3, 0.04)]))}
    todo = [i for i in range(n) if i % 17 != 3]  # a resumed run: some tasks already done
    for world in (2, 4, 8):
        parts = [worker.shard_tasks(todo, n, r, world, pattern) for r in range(world)]
        assert sorted(i for p in parts for i in p) == todo
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        touched = [len({pattern[i] for i in p}) for p in parts]
        inter = [len({pattern[i] for i in worker.shard_tasks(todo, n, r, world)}) for r in range(world)]
        assert max(touched) <= 11 // world + 3 and sum(touched) < sum(inter), (world, touched, inter)
    assert worker.shard_tasks(todo, n, 0, 1, pattern) == todo
