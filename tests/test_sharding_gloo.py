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
