#!/usr/bin/env python
"""bench.py -- log points / second of the ReMo3D forward-solve hot path on B200.

A *step* is one pass of the hot path over one mesh task (workload C4):
    upload mesh -> symbolic phase (H1 order-2 space, CSR pattern) -> numeric assembly ->
    preconditioner setup -> point-source right-hand sides -> multi-RHS PCG to 1e-10 ->
    axis sampling -> apparent resistivity of every log point of the task.
The mesh is the synthetic 3D dipping-bed + inclusion model (half-ball R = 50 m, graded to the
electrodes, seed 0) sized to ~5 M degrees of freedom at order 2; the task is one batch of the
planner (batch_size 5 sources, the 4 tools of config C5) and yields `points_per_step` log points.

  value : whole-job log points/s with the mesh arrays already resident in HBM (CUDA tensors)
  e2e   : the same through the public API with HOST (pinned) mesh arrays and the result read back
          to the host every step (h2d / d2h inside the timed region)
  --impl reference : the CPU restatement of the reference path (oracle/, NumPy/SciPy; NGSolve itself is
          not installable) farmed over all host cores on a bounded sample of the workload.

Launch: `python bench.py --gpus 1 --steps K --warmup W`, or under torchrun with one rank per GPU
(points are sharded over ranks, no collective in the data path; timing = max over ranks).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOOLS = ["A2.0M0.5N", "N0.5M2.0A", "B5.7A0.4M", "M4.0A0.5B"]
SIGMA = [1 / 1.0, 1 / 10.0, 1 / 100.0, 1 / 10.0, 1 / 2.0]       # mud, 10/100/10 ohm-m beds, inclusion
SIZES = {
    # name: (h_electrode, h_axis, grading, h_max)  -> dofs at order 2
    "20M": (0.006, 0.013, 0.075, 3.0),   # config C5: ~2.6 M vertices, ~20 M dofs
    "5M": (0.008, 0.0222, 0.125, 4.0),
    "1M": (0.01, 0.04, 0.19, 5.0),
    "200k": (0.03, 0.1, 0.33, 6.0),
    "60k": (0.06, 0.25, 0.5, 6.0),
}


def make_task():
    from remo3d_b200 import planner, tools as tl

    params, sec = tl.set_tools_parameters(TOOLS)
    _, tasks = planner.prepare_simulation_depths_and_tasks(params, sec, np.arange(0, 100, 0.1), 5)
    task = tasks[len(tasks) // 2]
    return task, planner.flatten_task(task, params, three_d=True)


MESH_SLIVER_ROUNDS = 6


def mesh_rounds():
    return int(os.environ.get("REMO_BENCH_MESH_IMPROVE", MESH_SLIVER_ROUNDS))


def make_mesh(size, task, log=lambda *a: None):
    """Synthetic C4 mesh (cached under the system temp dir: all ranks and both arms share it)."""
    from remo3d_b200 import meshgen

    he, ha, g, hm = SIZES[size]
    # sliver pass of the mesher (meshgen.half_ball_mesh(improve=N)): Delaunay meshes of well-spaced points still hold a few
    # slivers, and the PCG iteration count follows the worst of them (profiles/r01_notes.md: 408-420 -> 189-196 iterations
    # at 1.4 M dofs on B200).  Gmsh / Netgen, which the reference meshes with, optimise their tets the same way.
    # REMO_BENCH_MESH_IMPROVE=0 gives the meshes of the first sessions of round 1.
    improve = mesh_rounds()
    key = hashlib.sha1(repr((size, he, ha, g, hm, task[1][0].tolist(), 3) + (("sliver-pass-v3", improve) if improve else ())).encode()).hexdigest()[:12]
    name = "remo3d_bench_mesh_%s.npz" % key
    path = os.path.join(os.environ.get("REMO_MESH_CACHE", tempfile.gettempdir()), name)
    # caches: REMO_MESH_CACHE / the system temp dir (written below), and <repo>/.mesh_cache (read only: a mesh generated
    # ahead with this same function travels with the working tree, so a fresh box need not spend minutes in Qhull)
    for cand in (path, os.path.join(ROOT, ".mesh_cache", name)):
        if os.path.exists(cand):
            try:
                z = np.load(cand)
                return {k: z[k] for k in z.files}
            except Exception as exc:  # truncated copy: generate
                log("mesh cache %s unreadable (%s)" % (cand, exc))
    t0 = time.time()
    material = meshgen.layered_material([-1.0, 1.5], dip_rad=np.deg2rad(30.0), borehole_radius=0.1, inclusion=((3.0, 2.0, 1.0), 1.5))
    m = meshgen.half_ball_mesh(50.0, task[1][0], material=material, h_electrode=he, h_axis=ha, grading=g, h_max=hm, seed=0, improve=improve)
    from remo3d_b200.mesh import Mesh

    mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
    out = {"points": mesh.points, "elems": mesh.elems, "mat": mesh.mat, "bfacets": mesh.bfacets,
           "bdir": mesh.dirichlet_flags("dirichlet_boundary"), "axis": mesh.axis_vertices()}
    tmp = path + ".%d.tmp.npz" % os.getpid()
    np.savez(tmp, **out)
    os.replace(tmp, path)
    log("mesh %s generated in %.1f s: %d vertices, %d tets" % (size, time.time() - t0, mesh.nv, mesh.ne))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


SPMM_KERNELS = {
    0: "k_spmm_p / k_spmm_g (CSR PCG SpMM + fused p.q)",
    1: "k_spmm_stream8 (SELL-8 PCG SpMM + fused p.q, cp.async-staged matrix stream)",
    2: "k_spmm_ebe (element-wise PCG product Q = A P + fused p.q from 10 metric numbers per tet, no assembled matrix read; "
       "includes the memset of Q; bytes counted as the CSR SpMM it replaces, SURVEY 8d)",
}


def spmm_bytes(nnz, ndof, k):
    """Algorithmic bytes of one SpMM launch: fp64 values + int32 columns, int64 row pointers, P read once,
    Q written once: 12 nnz + N (8 + 16 k)."""
    return 12.0 * nnz + ndof * (8.0 + 16.0 * k)


def run_b200(args):
    import torch
    import torch.distributed as dist

    from remo3d_b200 import _cabi

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; this arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    log = (lambda *a: print(*a, file=sys.stderr, flush=True)) if rank == 0 else (lambda *a: None)

    task, flat = make_task()
    if rank == 0:
        m = make_mesh(args.size, task, log)
    if world > 1:
        dist.barrier()
    if rank != 0:
        m = make_mesh(args.size, task)

    # `--contexts` solver contexts per GPU, each with its own stream and host thread (ctypes releases the GIL): the mesh
    # tasks of a rank are independent, so while one context is in the small launch-bound kernels of its V-cycle or in the
    # sort-heavy symbolic phase of its next mesh, the other one's SpMM fills the SMs.  Same work per step, same kernels.
    nctx = max(1, args.contexts)
    main_stream = torch.cuda.Stream()
    streams = [torch.cuda.Stream() for _ in range(nctx)]
    ctxs = [_cabi.Context(local) for _ in range(nctx)]
    # REMO_BENCH_OPTS="name=value,...": remo_set_option on every context (A/B runs of solver options, e.g. amg_agg=0)
    bench_opts = [kv.split("=") for kv in os.environ.get("REMO_BENCH_OPTS", "").split(",") if "=" in kv]
    for cx, sx in zip(ctxs, streams):
        cx.set_stream(sx.cuda_stream)
        for name, value in bench_opts:
            cx.set_option(name.strip(), float(value))
    ctx, stream = ctxs[0], streams[0]
    names = ["points", "elems", "mat", "bfacets", "bdir", "axis"]

    def load(mesh):
        h = {k: torch.from_numpy(np.ascontiguousarray(mesh[k])).pin_memory() for k in names}
        return h, {k: h[k].cuda() for k in names}

    host, dev = load(m)
    npts = flat["pt_rhs"].shape[0]
    nrhs = flat["src_ptr"].shape[0] - 1
    ra_hosts = [torch.empty(npts, dtype=torch.float64).pin_memory() for _ in range(nctx)]
    info = {}

    cur_order = [args.order]  # the order-3 companion leg switches it

    def step(a, i=0):
        cx = ctxs[i]
        cx.mesh_set(3, a["points"], a["elems"], a["mat"], a["bfacets"], a["bdir"], a["axis"])
        cx.space_build(cur_order[0])
        cx.assemble(SIGMA)
        cx.precond_setup(args.preconditioner)
        cx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
        it, rel = cx.solve(rtol=1e-10, maxit=args.maxit)
        cx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"], out=ra_hosts[i].numpy())
        info.update(iters=it.tolist(), relres=float(rel.max()))

    def timed(a, steps, contexts=None):
        """Exactly `steps` steps, pulled from a shared counter by one host thread per context; device time from an event
        recorded before any context may start to one recorded after every context's stream has drained."""
        use = nctx if contexts is None else contexts
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_stream)
        for sx in streams[:use]:
            sx.wait_event(e0)
        lock, nxt, errs = threading.Lock(), [0], []

        def work(i):
            torch.cuda.set_device(local)
            try:
                while True:
                    with lock:
                        j = nxt[0]
                        nxt[0] += 1
                    if j >= steps:
                        break
                    step(a, i)
            except Exception as exc:  # surfaced below: a failed step invalidates the measurement
                errs.append(exc)

        if use == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(i,)) for i in range(use)]
            [t.start() for t in th]
            [t.join() for t in th]
        if errs:
            raise errs[0]
        for sx in streams[:use]:
            main_stream.wait_stream(sx)
        e1.record(main_stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if mesh_rounds() > 0:
        # the sliver pass only changes the INPUT (vertex positions of a few thousand tets); should the path reject such a
        # mesh, say so loudly and measure on the plain mesh instead of measuring nothing
        try:
            step(dev, 0)
        except _cabi.RemoError as exc:
            log("bench.py: the mesh with the sliver pass was rejected (%s); falling back to REMO_BENCH_MESH_IMPROVE=0" % exc)
            os.environ["REMO_BENCH_MESH_IMPROVE"] = "0"
            m = make_mesh(args.size, task, log)
            host, dev = load(m)
    for i in range(nctx):
        for _ in range(args.warmup):
            step(dev, i)
        step(host, i)  # also warm the host-buffer path once (first use of the pinned buffers), untimed
    torch.cuda.synchronize()
    launches0 = sum(cx.launch_count() for cx in ctxs)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev = timed(dev, args.steps)
    launches = sum(cx.launch_count() for cx in ctxs) - launches0
    ms_e2e = timed(host, args.steps)
    # roofline leg: the same steps once more with CUDA events around every SpMM launch of the PCG (remo_profile).  Events
    # cannot sit inside a CUDA graph, so this leg replays the iterations as plain launches; the SpMM kernel is identical.
    ctx.profile(True)
    ms_prof = timed(dev, args.steps, contexts=1)
    spmm_ms, spmm_n = ctx.profile_get()
    ctx.profile(False)
    kind = ctx.spmm_kind()  # 0 CSR, 1 SELL copy, 2 element-wise (remo_spmm_kind)
    stage = ctx.stage_times()  # of the last step of the single-context leg (under two contexts the stages interleave)
    if os.environ.get("REMO_BENCH_DEBUG"):
        log("debug: dev %.1f ms, e2e %.1f ms, profiled (no graph) %.1f ms" % (ms_dev, ms_e2e, ms_prof))
    clocks = sampler.stop() if sampler else None
    # everything that describes the main workload is read BEFORE the companion legs below load other meshes
    ndof, nnz = ctx.ndof, ctx.nnz  # nnz: builds the CSR pattern now, outside the timed regions (the PCG path never needs it)
    amg_levels = ctx.precond_get()[2] if args.preconditioner == "multigrid" else []
    try:
        asm_ms = float(ctx.kernel_time(1, nrhs, 3))
    except Exception as exc:  # reporting only
        asm_ms = None
        log("assembly timing failed: %r" % exc)
    main_info = dict(info)

    def companion(mesh, steps):
        """The same timed loop on another mesh of the same generator (all contexts, device-resident arrays)."""
        _, d2 = load(mesh)
        for i in range(nctx):
            step(d2, i)
        ms = timed(d2, steps)
        return {"value": npts * steps * world / (ms / 1e3), "ms_per_step": ms / steps, "iterations": info.get("iters"), "max_relres": info.get("relres"),
                "ndof": ctxs[0].ndof, "steps": steps}, ra_hosts[0].numpy().copy()

    plain, like, order3 = None, None, None
    if world == 1 and not args.no_companions and args.order == 2:
        # the reference hard-wires order 3 (ngsolve_functions.py:27): the same task on the 1M-size mesh of the generator at order 3
        # has the dof count of the headline workload; its product runs on the order-3 instantiation of the element-wise kernel
        try:
            cur_order[0] = 3
            m3 = make_mesh("1M", task, log)
            o3, _ = companion(m3, max(4, args.steps // 2))
            _, d3 = load(m3)
            ctx.profile(True)
            timed(d3, 1, contexts=1)
            ms3, n3 = ctx.profile_get()
            ctx.profile(False)
            nd3, nz3, kind3 = ctx.ndof, ctx.nnz, ctx.spmm_kind()
            per3 = ms3 / max(n3, 1) / 1e3
            o3.update({"order": 3, "nnz": nz3, "mesh": "size class 1M of the same generator (%d vertices, %d tets)" % (m3["points"].shape[0], m3["elems"].shape[0]),
                       "product_kind": kind3, "product_avg_launch_ms": per3 * 1e3, "product_launches_timed": int(n3),
                       "product_bytes_per_launch": spmm_bytes(nz3, nd3, nrhs),
                       "product_achieved_gbs": spmm_bytes(nz3, nd3, nrhs) / per3 / 1e9 if n3 else None})
            order3 = o3
        except Exception as exc:
            log("order-3 leg failed: %r" % exc)
        finally:
            cur_order[0] = args.order
    if world == 1 and not args.no_companions:
        if mesh_rounds() > 0:
            # like-for-like with round 1's first sessions: the same workload on the mesh WITHOUT the sliver pass
            saved = os.environ.get("REMO_BENCH_MESH_IMPROVE")
            os.environ["REMO_BENCH_MESH_IMPROVE"] = "0"
            try:
                plain, _ = companion(make_mesh(args.size, task, log), args.steps)
            except Exception as exc:
                log("plain-mesh leg failed: %r" % exc)
            finally:
                if saved is None:
                    os.environ.pop("REMO_BENCH_MESH_IMPROVE", None)
                else:
                    os.environ["REMO_BENCH_MESH_IMPROVE"] = saved
        if not args.no_cpu_baseline:
            # the CPU arm's larger sample through the GPU arm: both sides of `like_for_like` are measured on one mesh, and
            # the apparent resistivities of the two arms are compared (parity on the benchmarked generator)
            like, like_ra = companion(make_mesh(args.cpu_size2, task, log), max(8, args.steps))

    if rank == 0:
        total_pts = npts * args.steps * world
        value = total_pts / (ms_dev / 1e3)
        e2e = total_pts / (ms_e2e / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        if order3 and order3.get("product_achieved_gbs"):
            order3["product_frac_of_measured_hbm"] = order3["product_achieved_gbs"] / peak
        per_launch = spmm_ms / max(spmm_n, 1) / 1e3
        traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per SpMM launch from the committed ncu --set full capture
        for rnd in ("r02", "r01"):  # the newest committed ncu --set full capture of this kernel at this size
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "%s_spmm%s_traffic_%s.json" % (rnd, "_ebe" if kind == 2 else "", args.size))))
                if tj.get("order") == args.order and tj.get("nrhs") == nrhs and kind != 0:
                    traffic = tj["traffic_bytes_per_launch"]
                    break
            except Exception:
                pass
        achieved = spmm_bytes(nnz, ndof, nrhs) / per_launch / 1e9 if spmm_n else 0.0
        h2d = sum(host[k].numel() * host[k].element_size() for k in names) + flat["src_z"].nbytes * 2 + flat["src_ptr"].nbytes \
            + flat["pt_rhs"].nbytes + 3 * flat["pt_z0"].nbytes + len(SIGMA) * 8
        line = {
            "metric": "log points/sec", "value": value, "unit": "log points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "C4: synthetic 3D dipping-bed (30 deg, 10/100/10 ohm-m) + spherical inclusion, half-ball R=50 m, "
                            "order-%d H1, one mesh task per step (batch_size 5, 4 tools), size %s" % (args.order, args.size),
                "ndof": ndof, "nnz": nnz, "nnz_per_row": nnz / ndof, "vertices": int(m["points"].shape[0]), "tets": int(m["elems"].shape[0]),
                "nrhs": nrhs, "points_per_step": npts, "solves_per_step": nrhs, "preconditioner": args.preconditioner,
                "rtol": 1e-10, "iterations": main_info.get("iters"), "max_relres": main_info.get("relres"),
                "l2": "inputs larger than L2 (matrix %.0f MB + vectors %.0f MB vs 126 MB L2); no explicit flush" % (12e-6 * nnz, 48e-6 * ndof * nrhs),
                "sharding": "independent mesh tasks per rank, no data-path collective",
                "contexts_per_gpu": nctx, "stage_ms_one_context_alone": stage,
                "mesh_sliver_pass_rounds": mesh_rounds(),
                "options": {k: float(v) for k, v in bench_opts},
                "amg_levels": amg_levels,
                "value_plain_mesh": plain,
                "order3_companion": order3,
            },
            "e2e": {"value": e2e, "unit": "log points/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(npts * 8),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": SPMM_KERNELS[kind], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "bytes_per_launch": spmm_bytes(nnz, ndof, nrhs), "avg_launch_ms": per_launch * 1e3, "launches_timed": int(spmm_n),
                         "frac_of_8TBs_spec": achieved / 8000.0, "spmm_share_of_step": spmm_ms / ms_prof,
                         "bytes_note": "achieved / frac use the ALGORITHMIC bytes of the CSR SpMM this launch replaces (SURVEY 8d: 12 nnz + N (8 + 16 k)); "
                                       "the element-wise kernel (kind 2) reads no matrix and really moves `traffic` bytes",
                         "frac_of_measured_traffic": (traffic / per_launch / 1e9 / peak) if (traffic and spmm_n) else None,
                         "limiter": ("shared-memory / L1 data pipe (ncu: l1tex throughput ~87 %, DRAM ~25-40 %): bound=hbm names the roofline the "
                                     "contract asks for, not the unit that saturates") if kind == 2 else "latency x concurrency of the scattered P-row gathers",
                         "timed_with": "CUDA events around every SpMM launch in %d extra steps run right after the timed regions (plain launches instead of the CUDA graph)" % args.steps},
        }
        try:
            # numeric assembly (remo_assemble: geometry + atomic-free row-gather kernels + the sigma copy) against the same
            # peak; algorithmic bytes per element as in SURVEY 8d: vertex ids + coordinates + material + the ldof^2 scatter
            # map (4 B) and values (8 B)
            ldof = {1: 4, 2: 10, 3: 20}[args.order]
            abytes = float(m["elems"].shape[0]) * (16 + 96 + 4 + 12 * ldof * ldof)
            ams = asm_ms
            line["assembly"] = {"ms": ams, "algorithmic_bytes": abytes, "achieved": abytes / ams / 1e6, "unit": "GB/s",
                                "frac": abytes / ams / 1e6 / peak, "nnz_per_s": nnz / ams * 1e3,
                                "note": "element metrics + atomic-free row-gather assembly of the whole CSR matrix (remo_kernel_time, CUDA events, "
                                        "3 repetitions after the timed regions).  NOT part of a step any more: the element-wise PCG path needs "
                                        "only the element metrics (stage 'assemble'), diag(A) and the vertex block; the CSR matrix is built on demand"}
        except Exception as e:  # reporting only
            line["assembly"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(args, log)
            ra_cpu = np.asarray(cb.pop("_ra_sample"))
            line["cpu_baseline"] = cb
            if like is not None:
                err = float(np.max(np.abs(like_ra - ra_cpu) / np.abs(ra_cpu)))
                line["parity"] = {"max_rel_err_ra": err, "n_points": int(ra_cpu.shape[0]), "tol": 1e-6, "ndof": like["ndof"],
                                  "what": "apparent resistivities of one C4 mesh task (size %s, sliver pass %d) from the GPU arm vs the CPU arm "
                                          "(oracle, two-level PCG), both to relative residual 1e-10" % (args.cpu_size2, mesh_rounds())}
                line["like_for_like"] = {"ndof": like["ndof"], "gpu_value": like["value"], "gpu_ms_per_step": like["ms_per_step"],
                                         "gpu_iterations": like["iterations"], "cpu_value": cb["sample_value_unscaled"],
                                         "cpu_iterations": cb["sample_iterations"], "ratio": like["value"] / cb["sample_value_unscaled"],
                                         "unit": "log points/s", "note": "both arms MEASURED on the same mesh (no extrapolation)"}
                if not err <= 1e-6:
                    emit(line)
                    raise SystemExit("bench.py: GPU and CPU arms disagree on the apparent resistivities: max rel err %.3e > 1e-6" % err)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for cx in ctxs:
        cx.close()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy/SciPy restatement of the reference path) farmed over the host cores
# ------------------------------------------------------------------------------------------------
CPU_RTOL = 1e-10  # the same stopping criterion as the GPU arm


def _cpu_task(payload):
    from threadpoolctl import threadpool_limits

    from oracle import fem_oracle as fo

    threadpool_limits(1)  # one BLAS/OpenMP thread per worker process: the farm supplies the parallelism
    m, flat, order = payload
    t0 = time.time()
    # solver="multigrid": the reference's default preconditioner (remo3d.py:82 -> ngsolve_functions.py:46), restated as
    # exact sparse LU of the P1 block + Jacobi on the high-order dofs, all right-hand sides of the task in one block
    res = fo.solve_task(m["points"], m["elems"], m["mat"], SIGMA, m["bfacets"], m["bdir"].astype(bool), order, flat, solver="multigrid",
                        rtol=CPU_RTOL)
    return time.time() - t0, res["ra"].tolist(), res["space"].ndof, [int(i) for i in res["iters"]]


def cpu_farm(size, order, steps, workers):
    """`steps` rounds; in each round every worker solves one mesh task (the reference's MPI farm, remo3d.py:843-860)."""
    import multiprocessing as mp

    task, flat = make_task()
    m = make_mesh(size, task)
    npts = flat["pt_rhs"].shape[0]
    ctxm = mp.get_context("fork")
    rounds = []
    with ctxm.Pool(workers) as pool:
        t0 = time.time()
        for _ in range(steps):
            t1 = time.time()
            out = pool.map(_cpu_task, [(m, flat, order)] * workers)
            rounds.append(time.time() - t1)
        wall = time.time() - t0
    # steady-state round time: the first round of a pool carries the warm-up of its processes (imports, page cache, the first
    # LU), so with more than one round it is left out of the median; the two samples of cpu_scaled are then comparable and the
    # fitted exponent stops moving by +-0.05 from run to run (which was +-20 % on the extrapolated figure)
    steady = rounds[1:] if len(rounds) > 1 else rounds
    return {"value": npts * workers * steps / wall, "wall_s": wall, "ndof": out[0][2], "points": npts * workers * steps,
            "round_s": float(np.median(steady)), "rounds_s": rounds, "ra": out[0][1], "iters": out[0][3], "size": size}


def cpu_scaled(args, rounds, log):
    """CPU arm on BOUNDED samples of workload C4.

    The oracle cannot finish a ~5 M-dof task in minutes (SciPy assembly of 137 M non-zeros + a sparse LU of the 613 k-row
    P1 block per worker process), so the farm is MEASURED on the same generator at two reduced sizes (same order, same 5
    right-hand sides, same tolerance, the reference's default two-level preconditioner).  Two numbers come out:
      * `like_for_like`: the larger sample (--cpu-size2, ~202 k dofs) is also run through the GPU arm by run_b200, so both
        sides of that ratio are measurements on the same mesh;
      * `value`: the round time extrapolated to the dof count of `--size` with the growth exponent p MEASURED between
        the two samples (t ~ ndof^p; clamped to [1.2, 1.8]: the sparse LU of the P1 block and the iteration count both
        grow faster than linearly in 3D)."""
    workers = os.cpu_count() or 1
    task, flat = make_task()
    npts = flat["pt_rhs"].shape[0]
    small = cpu_farm(args.cpu_size, args.order, 4, workers)           # ~3 s per round
    big = cpu_farm(args.cpu_size2, args.order, min(max(3, rounds), 4), workers)  # ~13 s per round: the whole arm stays under 1.5 min
    p_raw = float(np.log(big["round_s"] / small["round_s"]) / np.log(big["ndof"] / small["ndof"]))
    p = min(max(p_raw, 1.2), 1.8)
    full = FULL_DOFS.get((args.size, args.order))
    if full is None:  # dof count of the GPU arm's mesh: vertices + edges (order 2), counted from the cached mesh
        from oracle import fem_oracle as fo

        mm = make_mesh(args.size, task)
        full = fo.Space(mm["points"].shape[0], mm["elems"], args.order, 3).ndof if mm["elems"].shape[0] < 400000 else None
    if full is None:
        full = big["ndof"]
    t_full = big["round_s"] * (full / big["ndof"]) ** p
    value = npts * workers / t_full
    log("cpu arm: %d workers; %s: %d dofs %.1f s/round; %s: %d dofs %s s/round (iterations %s); exponent %.2f (raw %.2f) -> %.0f s/round at %d dofs" % (
        workers, args.cpu_size, small["ndof"], small["round_s"], args.cpu_size2, big["ndof"], ["%.1f" % r for r in big["rounds_s"]], big["iters"],
        p, p_raw, t_full, full))
    sample = ("oracle/fem_oracle.py (NumPy/SciPy restatement of the reference path with its default preconditioner: exact sparse LU of the P1 "
              "block + Jacobi on the high-order dofs, PCG to 1e-10, 5 right-hand sides per task in one block; NGSolve is not installable) farmed "
              "over %d worker processes on bounded samples of workload C4: size %s (%d dofs) %.1f s per round, size %s (%d dofs) median %.1f s "
              "per round of %d rounds (%s iterations); time per task ~ ndof^%.2f (measured %.2f, clamped to [1.2, 1.8]) extrapolated to %d dofs "
              "-> %.0f s per round of %d tasks; MEASURED throughput on the %s sample: %.2f log points/s"
              % (workers, args.cpu_size, small["ndof"], small["round_s"], args.cpu_size2, big["ndof"], big["round_s"], len(big["rounds_s"]),
                 big["iters"], p, p_raw, full, t_full, workers, args.cpu_size2, npts * workers / big["round_s"]))
    return {"value": value, "unit": "log points/s", "cores": workers, "kind": "port", "sample": sample,
            "sample_value_unscaled": npts * workers / big["round_s"], "sample_ndof": big["ndof"], "sample_size": args.cpu_size2,
            "sample_iterations": big["iters"], "exponent": p, "exponent_measured": p_raw,
            "wall_s": small["wall_s"] + big["wall_s"], "round_points": npts * workers, "_ra_sample": big["ra"]}


FULL_DOFS = {("5M", 2): 4821007, ("1M", 2): 1422109}


def cpu_baseline(args, log):
    return cpu_scaled(args, 2, log)


# ------------------------------------------------------------------------------------------------
# pipeline mode: the whole measurement pipeline, host mesh generation inside the timed region
# ------------------------------------------------------------------------------------------------
def run_pipeline(args):
    """`Model.simulate_logs` end to end: planner -> host mesh pool (one triangulation per electrode pattern,
    materials per task) -> GPU workers -> gather, on the C5-shaped plan (--pipeline-depths depths x 4 tools, batch 5) of a
    3-layer 10/100/10 ohm-m model dipping 30 degrees.  Everything is inside the timed region (wall clock: the host is part of
    what is measured).  Under torchrun every rank drives its own GPU and every world-th task; one gather at the end."""
    import torch
    import torch.distributed as dist

    from remo3d_b200 import Model

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; this arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    he, ha, g, hm = SIZES[args.pipeline_size]
    formation = np.array([[-1000.0, 49.0, np.nan, np.nan, 10.0], [49.0, 51.5, np.nan, np.nan, 100.0], [51.5, 2000.0, np.nan, np.nan, 10.0]])
    borehole = np.array([[-1000.0, 0.2, 1.0], [2000.0, 0.2, 1.0]])
    depths = np.round(np.arange(args.pipeline_depths) * 0.1, 4)
    model = Model(TOOLS)
    model.set_model_parameters(formation, borehole, dip=30)
    cpu = max(1, (os.cpu_count() or 1) // world)
    model.initialize_workers(cpu_workers=cpu, gpu_workers=1, devices=[local], contexts_per_gpu=args.contexts)
    # --pipeline-conforming 1: every task meshes its own interfaces (the Model default, reference-like); 0: one triangulation
    # per electrode pattern, materials per tet centroid (the host-light mode the sharing was built for)
    opts = {"h_electrode": he, "h_axis": ha, "grading": g, "h_max": hm, "improve": mesh_rounds() if args.pipeline_improve else 0,
            "conforming": bool(args.pipeline_conforming)}
    try:
        # warm-up: one small call (CUDA context, kernels, pool processes), untimed
        model.simulate_logs(depths[:2], order=args.order, preconditioner=args.preconditioner, mesh_options=dict(opts, h_electrode=0.1, h_axis=0.4, grading=0.6, improve=0))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        model.simulate_logs(depths, order=args.order, preconditioner=args.preconditioner, mesh_options=opts,
                            task_shard=(rank, world) if world > 1 else None)
        torch.cuda.synchronize()
        wall = time.time() - t0
    finally:
        model.shutdown_workers()
    st = model.pipeline_stats
    if world > 1:
        t = torch.tensor([wall, st["gpu_busy_fraction"], float(st["tasks"])], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        wall, busy, tasks = float(tmax[0]), float(t[1]) / world, int(t[2])
    else:
        busy, tasks = st["gpu_busy_fraction"], st["tasks"]
    if rank == 0:
        npts = sum(int(np.isfinite(model.logs[t][:, 1]).sum()) for t in TOOLS)
        recs = [r for r in model.task_records if r and "error" not in r]
        errs = [r for r in model.task_records if r and "error" in r]
        line = {"metric": "log points/sec", "mode": "pipeline", "value": npts / wall, "unit": "log points/s", "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "pipeline: Model.simulate_logs, %d depths x 4 tools, batch 5, 3 layers dipping 30 deg, half-ball R=50, order %d, "
                                       "mesh size class %s; mesh generation inside the timed region" % (args.pipeline_depths, args.order, args.pipeline_size),
                           "tasks": tasks, "log_points": npts, "failed_tasks": len(errs), "first_error": errs[0]["error"] if errs else None,
                           "ndof_median": float(np.median([r["ndof"] for r in recs])) if recs else None,
                           "iterations_median": float(np.median([max(r["iters"]) for r in recs])) if recs else None,
                           "gpu_busy_fraction": busy, "host_material_s_rank0": st["host_material_s"], "shared_geometry": st["shared_geometry"],
                           "cpu_workers_per_rank": cpu, "contexts_per_gpu": args.contexts, "conforming_interfaces": bool(args.pipeline_conforming), "mesh_sliver_pass_rounds": opts["improve"],
                           "solve_ms_median": float(np.median([r["solve"] for r in recs])) if recs else None,
                           "timing": "wall clock around simulate_logs (host work is part of the measurement), max over ranks"}}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_example01(args):
    """Config C1: the reference's own Example_01 inputs (tests/golden/example_01), one normal tool
    B5.7A0.4M, 100 measurement points, defaults of `Model.compute_synthetic_logs` (2D axisymmetric, order 3, "multigrid",
    batch 5), mesh generation inside the timed region.  Reports log points/s (the README's 15-30 s for such a run on a
    Ryzen 2600 is 3.3-6.7 log points/s: `vs_baseline` against the upper figure) and the parity of the log against the
    reference's committed output."""
    import torch

    from remo3d_b200 import Model

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; this arm has no CPU fallback")
    d = os.path.join(ROOT, "tests", "golden", "example_01")
    gold = np.loadtxt(os.path.join(d, "Results_1.txt"), skiprows=2)
    names = open(os.path.join(d, "Results_1.txt")).readline().split()[1:]
    tool = "B5.7A0.4M"
    depths = gold[75:175, 0].copy()  # 100 points, 7.5 .. 17.4 m
    cpu = os.cpu_count() or 4
    model = Model([tool])
    model.set_model_parameters(os.path.join(d, "Formation.txt"), os.path.join(d, "Borehole.txt"))
    model.initialize_workers(cpu_workers=cpu, gpu_workers=1, contexts_per_gpu=args.contexts)
    try:
        model.simulate_logs(depths[:5])  # warm-up (CUDA context, pool processes), untimed
        torch.cuda.synchronize()
        t0 = time.time()
        model.simulate_logs(depths)
        torch.cuda.synchronize()
        wall = time.time() - t0
    finally:
        model.shutdown_workers()
    ref = gold[75:175, names.index(tool) + 1]
    rel = np.abs(model.logs[tool][:, 1] - ref) / ref
    recs = [r for r in model.task_records if r and "error" not in r]
    published = 100.0 / 15.0  # README.md:25-26, best case
    emit({"metric": "log points/sec", "mode": "example01", "value": depths.shape[0] / wall, "unit": "log points/s", "n_gpus": 1, "steps": 1, "warmup": 1,
          "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": depths.shape[0] / wall / published, "dtype": "f64",
          "data": "reference example inputs (tests/golden/example_01)",
          "config": {"workload": "C1: Examples/Example_01, tool %s, 100 depths, 2D axisymmetric, order 3, multigrid, batch 5; mesh generation inside the timed region" % tool,
                     "tasks": len(model.task_records), "ndof_median": float(np.median([r["ndof"] for r in recs])),
                     "iterations_median": float(np.median([max(r["iters"]) for r in recs])), "cpu_workers": cpu,
                     "gpu_busy_fraction": model.pipeline_stats["gpu_busy_fraction"],
                     "published": "README.md:25-26: 15-30 s for 100 points x 1 tool on a Ryzen 2600 (3.3-6.7 log points/s, end to end); vs_baseline uses 6.7",
                     "timing": "wall clock around simulate_logs"},
          "parity": {"max_rel_err_ra": float(rel.max()), "median_rel_err_ra": float(np.median(rel)), "n_points": int(rel.shape[0]),
                     "against": "Examples/Example_01/Output/Results_2024_08_17__18_59_29/Results_1.txt (the reference's committed log; its own two example outputs differ by 3.1e-4)"}})


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    log = lambda *a: print(*a, file=sys.stderr, flush=True)
    r = cpu_scaled(args, max(1, args.steps), log)
    line = {"impl": "reference", "metric": "log points/sec", "value": r["value"], "unit": "log points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["round_points"] / r["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4 (same generator, order %d), CPU arm on bounded samples scaled to size %s" % (args.order, args.size)},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "sample_value_unscaled", "sample_ndof", "exponent",
                                               "exponent_measured")},
            "e2e": {"value": r["value"], "unit": "log points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # libraries chat on stdout (NCCL prints its version there under torchrun): keep fd 1 for the JSON line alone and send
    # everything else to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", default="5M", choices=list(SIZES))
    ap.add_argument("--cpu-size", default="60k", choices=list(SIZES))
    ap.add_argument("--cpu-size2", default="200k", choices=list(SIZES))
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--preconditioner", default="multigrid", choices=["local", "multigrid"])
    ap.add_argument("--maxit", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-companions", action="store_true", help="skip the plain-mesh and like-for-like GPU legs (N = 1 only)")
    ap.add_argument("--contexts", type=int, default=3, help="solver contexts (stream + host thread) per GPU")
    ap.add_argument("--mode", default="step", choices=["step", "pipeline", "example01"])
    ap.add_argument("--pipeline-depths", type=int, default=1000)
    ap.add_argument("--pipeline-size", default="200k", choices=list(SIZES))
    ap.add_argument("--pipeline-improve", type=int, default=0)
    ap.add_argument("--pipeline-conforming", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "pipeline":
        run_pipeline(args)
    elif args.mode == "example01":
        run_example01(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
