"""Per-GPU solve worker: the B200 counterpart of `/root/reference/remo3d/workers/worker.py`.

The reference spawns MPI processes that each loop `mesh -> SolveBVP per source -> sample -> Ra`
(`worker.py:74-138`).  Here a worker is a thread bound to one GPU context; a task (= one mesh and all the sources
that share it, `remo3d.py:624-690`) is solved as ONE multi-right-hand-side system.  The failure contract is kept:
any exception inside a task turns every log point of that task into NaN (`worker.py:135-138`); the message is kept
in `errors` instead of being swallowed.
"""
import numpy as np

from . import _cabi, planner

DIRICHLET = "dirichlet_boundary"  # worker.py:90


def solve_task(ctx, mesh, sigma, flat, order=3, preconditioner="multigrid", rtol=1e-10, maxit=1000):
    """One mesh task on one context -> (Ra per log point, per-task record).  Raises on failure."""
    ctx.mesh_set(mesh.dim, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, mesh.dirichlet_flags(DIRICHLET), mesh.axis_vertices())
    ndof, _ = ctx.space_build(order)  # the CSR pattern (and its nnz) is built only on demand: not on this path
    ctx.assemble(np.asarray(sigma, dtype=np.float64))
    ctx.precond_setup(preconditioner)
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
    iters, relres = ctx.solve(rtol=rtol, maxit=maxit)
    ra = ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"])
    rec = {"ndof": ndof, "iters": iters.tolist(), "relres": float(relres.max())}
    rec.update(ctx.stage_times())
    return ra, rec


def run_tasks(ctx, jobs, tools, order=3, preconditioner="multigrid", rtol=1e-10, maxit=1000):
    """jobs: iterable of (task_index, task, mesh, sigma).  Yields (task_index, [[depth_idx, tool_idx, Ra], ...], record)."""
    for index, task, mesh, sigma in jobs:
        flat = planner.flatten_task(task, tools, three_d=(mesh.dim == 3))
        try:
            ra, rec = solve_task(ctx, mesh, sigma, flat, order, preconditioner, rtol, maxit)
        except Exception as exc:  # NaN-on-failure contract of worker.py:135-138
            ra = np.full(flat["pt_rhs"].shape[0], np.nan)
            rec = {"error": "%s: %s" % (type(exc).__name__, exc)}
        triples = [[int(d), int(t), float(r)] for d, t, r in zip(flat["pt_depth"], flat["pt_tool"], ra)]
        yield index, triples, rec


def shard(n_tasks, rank, world):
    """Static interleaved partition of task indices over ranks (independent units, no data-path collective)."""
    return list(range(rank, n_tasks, world))


def gather_results(local_triples, world=1):
    """The reference's single final gather (`remo3d.py:865`): every rank contributes its [depth, tool, Ra] triples.
    Uses torch.distributed (gloo or nccl) when a process group is initialised, else returns the local list."""
    if world <= 1:
        return list(local_triples)
    import torch.distributed as dist

    out = [None] * world
    dist.all_gather_object(out, list(local_triples))
    return [t for part in out for t in part]


def results_to_logs(triples, tools, measurement_depths):
    """`remo3d.py:868-874`: scatter the triples into logs[tool] = (n_depths, 2) arrays [depth, Ra]."""
    names = list(tools.keys())
    res = np.full((len(measurement_depths), len(names)), np.nan)
    for d, t, r in triples:
        res[d, t] = r
    return {name: np.vstack([measurement_depths, res[:, i]]).T for i, name in enumerate(names)}
