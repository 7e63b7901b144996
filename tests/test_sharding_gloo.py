"""Multi-GPU path on CPU: world_size-2 gloo run of the task sharding + final gather (remo3d.py:843-874 semantics)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from remo3d_b200 import planner, tools as tl, worker

TOOLS = ["A2.0M0.5N", "N0.5M2.0A", "B5.7A0.4M"]
DEPTHS = np.arange(0, 3, 0.1)


def _fake_ra(d, t):
    return 1.0 + 10.0 * d + 0.1 * t


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, sec = tl.set_tools_parameters(TOOLS)
    _, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, DEPTHS, 5)
    mine = worker.shard(len(tasks), rank, world)
    local = []
    for i in mine:  # stands in for run_tasks(): same triples layout, synthetic values (no GPU here)
        flat = planner.flatten_task(tasks[i], params, three_d=True)
        local += [[int(d), int(t), _fake_ra(d, t)] for d, t in zip(flat["pt_depth"], flat["pt_tool"])]
    allt = worker.gather_results(local, world)
    if rank == 0:
        logs = worker.results_to_logs(allt, params, DEPTHS)
        np.save(out, np.stack([logs[t] for t in TOOLS]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "logs.npy")
    mp.spawn(_rank_main, args=(2, port, out), nprocs=2, join=True)
    logs = np.load(out)
    assert logs.shape == (len(TOOLS), DEPTHS.shape[0], 2)
    for t in range(len(TOOLS)):
        np.testing.assert_array_equal(logs[t, :, 0], DEPTHS)
        np.testing.assert_allclose(logs[t, :, 1], [_fake_ra(d, t) for d in range(DEPTHS.shape[0])])


def test_shards_partition_the_tasks():
    for n, w in ((164, 8), (7, 2), (3, 4), (1, 1)):
        parts = [worker.shard(n, r, w) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


# ---- the real Model.simulate_logs over two ranks (pattern-aware shards of the shared-geometry 3D mode + the final gather)
FORMATION3 = np.array([[-1000.0, 1.0, np.nan, np.nan, 10.0], [1.0, 2.5, np.nan, np.nan, 100.0], [2.5, 2000.0, np.nan, np.nan, 10.0]])
BOREHOLE3 = np.array([[-1000.0, 0.2, 1.0], [2000.0, 0.2, 1.0]])
DEPTHS3 = np.round(np.arange(0.0, 2.0, 0.1), 4)
OPTS3 = {"h_electrode": 0.25, "h_axis": 0.8, "grading": 0.7, "conforming": False}


def _simulate(task_shard):
    import multiprocessing

    from remo3d_b200 import Model
    from tests import helpers

    m = Model(["A2.0M0.5N", "N0.5M2.0A"])
    m.set_model_parameters(FORMATION3, BOREHOLE3, dip=20)
    m.cpu_workers, m.gpu_workers = 2, 1
    m._mesh_pool = multiprocessing.get_context("fork").Pool(2)
    m._contexts = [helpers.OracleContext()]
    try:
        m.simulate_logs(DEPTHS3, order=1, mesh_options=OPTS3, task_shard=task_shard)
    finally:
        m._mesh_pool.terminate()
        m._contexts = None
    return m


def _model_rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _simulate((rank, world))
    assert m.pipeline_stats["shared_geometry"] and 0 < m.pipeline_stats["tasks"] < 8  # this rank solved only its slice
    if rank == 0:
        np.save(out, np.stack([m.logs[t] for t in m.tools]))
    dist.barrier()
    dist.destroy_process_group()


def test_model_simulate_logs_over_two_ranks(tmp_path):
    """`Model.simulate_logs(task_shard=(rank, 2))` on two gloo ranks (oracle as the solver): every rank meshes and solves its
    pattern-aware slice, the gathered logs equal the single-process run bit for bit (same meshes, same direct solves)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "logs3.npy")
    mp.spawn(_model_rank_main, args=(2, port, out), nprocs=2, join=True)
    two = np.load(out)
    one = _simulate(None)
    ref = np.stack([one.logs[t] for t in one.tools])
    assert np.isfinite(two[:, :, 1]).all()
    np.testing.assert_array_equal(two, ref)
