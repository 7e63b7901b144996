"""Per-GPU solve worker: the B200 counterpart of `/root/reference/remo3d/workers/worker.py`.

The reference spawns MPI processes that each loop `mesh -> SolveBVP per source -> sample -> Ra`
(`worker.py:74-138`).  Here a worker is a thread bound to one GPU context; a task (= one mesh and all the sources
that share it, `remo3d.py:624-690`) is solved as ONE multi-right-hand-side system (blocks of at most 32 right-hand
sides: a task with more sources, `batch_size > 32`, is solved block after block on the same matrix).  The failure
contract is kept: any exception inside a task -- mesh generation included, `worker.py:82-138` wraps both -- turns every
log point of that task into NaN and the run goes on; the message is kept in the task record instead of being swallowed.
A solve that stops at `maxit` is NOT a failure (NGSolve's `CGSolver(maxsteps=1000)` returns its last iterate too,
`ngsolve_functions.py:50-51`): the values are kept and the record carries `noconv` with the residual reached.
"""
import warnings

import numpy as np

from . import _cabi, planner

DIRICHLET = "dirichlet_boundary"  # worker.py:90
DEFAULT_MAXIT = {"multigrid": 1000, "local": 20000}  # ngsolve_functions.py:50 caps both at 1000; Jacobi needs ~2000 at 5 M dofs


def rhs_blocks(flat, max_rhs=_cabi.MAX_RHS):
    """Split the right-hand sides of a flattened task into blocks of at most `max_rhs`: yields (flat_block, point index)."""
    nrhs = flat["src_ptr"].shape[0] - 1
    if nrhs <= max_rhs:
        yield flat, np.arange(flat["pt_rhs"].shape[0])
        return
    for lo in range(0, nrhs, max_rhs):
        hi = min(nrhs, lo + max_rhs)
        s0, s1 = int(flat["src_ptr"][lo]), int(flat["src_ptr"][hi])
        sel = np.nonzero((flat["pt_rhs"] >= lo) & (flat["pt_rhs"] < hi))[0]
        block = dict(flat)
        block["src_ptr"] = flat["src_ptr"][lo:hi + 1] - s0
        block["src_z"], block["src_fac"] = flat["src_z"][s0:s1], flat["src_fac"][s0:s1]
        for k in ("pt_z0", "pt_z1", "pt_k", "pt_depth", "pt_tool"):
            block[k] = flat[k][sel]
        block["pt_rhs"] = (flat["pt_rhs"][sel] - lo).astype(np.int32)
        yield block, sel


def solve_task(ctx, mesh, sigma, flat, order=3, preconditioner="multigrid", rtol=1e-10, maxit=None):
    """One mesh task on one context -> (Ra per log point, per-task record).  Raises on failure."""
    if maxit is None:
        maxit = DEFAULT_MAXIT[preconditioner]
    ctx.mesh_set(mesh.dim, mesh.points, mesh.elems, mesh.mat, mesh.bfacets, mesh.dirichlet_flags(DIRICHLET), mesh.axis_vertices())
    ndof, _ = ctx.space_build(order)  # the CSR pattern (and its nnz) is built only on demand: not on this path
    ctx.assemble(np.asarray(sigma, dtype=np.float64))
    ctx.precond_setup(preconditioner)
    ra = np.empty(flat["pt_rhs"].shape[0])
    iters, relres = [], []
    for block, sel in rhs_blocks(flat):
        ctx.rhs_point_sources(block["src_ptr"], block["src_z"], block["src_fac"])
        it, rr = ctx.solve(rtol=rtol, maxit=maxit, raise_on_noconv=False)
        ra[sel] = ctx.apparent_resistivity(block["pt_rhs"], block["pt_z0"], block["pt_z1"], block["pt_k"], block["scale"])
        iters += it.tolist()
        relres += rr.tolist()
    rec = {"ndof": ndof, "iters": iters, "relres": float(max(relres))}
    if rec["relres"] > rtol:
        rec["noconv"] = "PCG stopped at maxit=%d with relative residual %.2e (> rtol %.1e): last iterate kept, like CGSolver(maxsteps)" % (
            maxit, rec["relres"], rtol)
        warnings.warn(rec["noconv"])
    rec.update(ctx.stage_times())
    return ra, rec


def run_tasks(ctx, jobs, tools, order=3, preconditioner="multigrid", rtol=1e-10, maxit=None):
    """jobs: iterable of (task_index, task, mesh, sigma) -- `mesh` may be an Exception raised by the mesh generator.
    Yields (task_index, [[depth_idx, tool_idx, Ra], ...], record)."""
    for index, task, mesh, sigma in jobs:
        points = [(int(p[0]), int(p[1])) for st in task[2] for p in st[2]]  # worker.py:104-134: the task's log points
        try:
            if isinstance(mesh, BaseException):
                raise mesh
            flat = planner.flatten_task(task, tools, three_d=(mesh.dim == 3))
            ra, rec = solve_task(ctx, mesh, sigma, flat, order, preconditioner, rtol, maxit)
            triples = [[int(d), int(t), float(r)] for d, t, r in zip(flat["pt_depth"], flat["pt_tool"], ra)]
        except Exception as exc:  # NaN-on-failure contract of worker.py:135-138
            triples = [[d, t, float("nan")] for d, t in points]
            rec = {"error": "%s: %s" % (type(exc).__name__, exc)}
        yield index, triples, rec


def shard(n_tasks, rank, world):
    """Static interleaved partition of task indices over ranks (independent units, no data-path collective)."""
    return list(range(rank, n_tasks, world))


def shard_tasks(todo, n_tasks, rank, world, pattern_of=None):
    """The tasks of `todo` this rank solves.  Default: the interleaved partition of `shard`.  With `pattern_of` (task -> key of
    the triangulation it shares with other tasks, `model_mesh.geometry_key`) the shards are pattern-aware: tasks ordered by
    (size of their pattern group, pattern), every rank takes a contiguous slice -- a rank then builds 2-3 of the
    triangulations instead of every rank building all of them (measured with interleaved shards at N = 8: 42 % GPU busy, the
    host pools of all ranks triangulating the same 11 patterns)."""
    if world <= 1:
        return list(todo)
    if pattern_of is None:
        mine = set(shard(n_tasks, rank, world))
        return [i for i in todo if i in mine]
    size = {}
    for i in todo:
        size[pattern_of[i]] = size.get(pattern_of[i], 0) + 1
    order = sorted(todo, key=lambda i: (-size[pattern_of[i]], pattern_of[i], i))
    lo, hi = (len(order) * rank) // world, (len(order) * (rank + 1)) // world
    return sorted(order[lo:hi])


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
