"""`Model`: drop-in for `/root/reference/remo3d/remo3d.py` `class Model` on the forward-solve path.

Same constructor, same method names and keyword arguments, same `logs` layout and the same ValueError messages;
the MPI worker farm (`remo3d.py:552-599, 809-899`) is replaced by an in-process queue: a host thread pool builds
meshes (`cpu_workers`), one worker thread per GPU (`gpu_workers`) solves them.  Extra keyword arguments that the
reference hard-codes: `order` (3, ngsolve_functions.py:27), `rtol` (CG tolerance), `mesh_options`.
Plotting (`save_results` figures) is out of scope; the text writer keeps the reference's file format.
"""
import datetime
import json
import multiprocessing
import os
import queue
import threading

import numpy as np

from . import _cabi, model_io, model_mesh, planner, tools as tl, worker


def _build_mesh_job(args):
    """Runs in a mesh-pool process: one task's mesh + sigma list (pure host work).  Exceptions are RETURNED, not raised:
    a mesh that cannot be built turns its own task into NaN log points (`worker.py:82-138`), not the whole run."""
    try:
        return model_mesh.build_task_mesh(*args)
    except Exception as exc:  # noqa: BLE001 -- the failure contract covers every error
        return exc, None


def _build_geometry_job(args):
    try:
        return model_mesh.build_geometry_3d(*args)
    except Exception as exc:  # noqa: BLE001
        return exc


class Model:
    conversion_table = model_io.CONVERSION_TABLE  # remo3d.py:26

    def __init__(self, tools, force_single_electrode_configuration=True):
        self.tools, self.sec = self.set_tools_parameters(tools, force_single_electrode_configuration=force_single_electrode_configuration)
        self.formation_model = None
        self.borehole_model = None
        self.dip_deg = None
        self.dip_rad = None
        self.cpu_workers = None
        self.gpu_workers = None
        self.logs = None
        self.task_records = None
        self.pipeline_stats = None
        self._contexts = None
        self._mesh_pool = None

    @classmethod
    def compute_synthetic_logs(cls, tools, measurement_depths, formation_model, borehole_model,
                               force_single_electrode_configuration=True, formation_units=["M", "M", "M"],
                               borehole_geometry_type="diameter", borehole_units=["M", "M"], dip=0, cpu_workers=4,
                               gpu_workers=0, domain_radius=50, batch_size=5, mesh_generator="auto",
                               preconditioner="multigrid", condense=True, **extra):
        """remo3d.py:65-174."""
        model = cls(tools, force_single_electrode_configuration=force_single_electrode_configuration)
        model.set_model_parameters(formation_model, borehole_model, borehole_geometry_type=borehole_geometry_type, dip=dip)
        model.initialize_workers(cpu_workers=cpu_workers, gpu_workers=gpu_workers)
        try:
            model.simulate_logs(measurement_depths, domain_radius=domain_radius, batch_size=batch_size, mesh_generator=mesh_generator,
                                preconditioner=preconditioner, condense=condense, **extra)
        finally:
            model.shutdown_workers()
        return model

    # ---- tools (remo3d.py:178-340)
    def set_tools_parameters(self, tools, force_single_electrode_configuration=True):
        return tl.set_tools_parameters(tools, force_single_electrode_configuration)

    def _set_tool_parameters(self, tool, electrodes, distances):
        return tl.tool_parameters(tool, electrodes, distances)

    # ---- model (remo3d.py:344-548)
    def set_model_parameters(self, formation_model, borehole_model, borehole_geometry_type="diameter", dip=0):
        if isinstance(formation_model, str):
            self.formation_model = self.load_formation_parameters(formation_model)
        elif isinstance(formation_model, np.ndarray):
            self.formation_model = self.set_formation_parameters(formation_model)
        if isinstance(borehole_model, str):
            self.borehole_model = self.load_borehole_parameters(borehole_model, borehole_geometry_type)
        elif isinstance(borehole_model, np.ndarray):
            self.borehole_model = self.set_borehole_parameters(borehole_model, borehole_geometry_type)
        self.dip_deg, self.dip_rad = self.set_dip(dip)
        self._check_model_geometry()

    def load_formation_parameters(self, formation_model_file):
        return model_io.load_formation_parameters(formation_model_file)

    def set_formation_parameters(self, formation_parameters, formation_units=["M", "M", "M"]):
        return model_io.set_formation_parameters(formation_parameters, formation_units)

    def load_borehole_parameters(self, borehole_model_file, borehole_geometry_type="diameter"):
        return model_io.load_borehole_parameters(borehole_model_file, borehole_geometry_type)

    def set_borehole_parameters(self, borehole_parameters, borehole_geometry_type="diameter", borehole_units=["M", "M"]):
        return model_io.set_borehole_parameters(borehole_parameters, borehole_geometry_type, borehole_units)

    def set_dip(self, dip):
        return model_io.set_dip(dip)

    def _check_model_geometry(self):
        model_io.check_model_geometry(self.formation_model, self.borehole_model)

    def _add_points_to_borehole(self, maximal_distance=0.15):
        return model_io.densify_borehole(self.borehole_model, maximal_distance)

    def _prepare_simulation_depths_and_tasks(self, measurement_depths, batch_size):
        return planner.prepare_simulation_depths_and_tasks(self.tools, self.sec, measurement_depths, batch_size)

    # ---- workers (remo3d.py:552-599, 887-899)
    def initialize_workers(self, cpu_workers=4, gpu_workers=0, contexts_per_gpu=3, devices=None):
        """`gpu_workers` = number of GPUs to shard the mesh tasks over (0 is promoted to 1: there is no CPU solve
        path in this package); `cpu_workers` = host processes that build meshes ahead of the GPUs.  Every GPU runs
        `contexts_per_gpu` solver contexts (own stream + host thread): mesh tasks are independent, and several in flight
        hide the launch-bound parts of one another (measured at 4.8 M dofs: two +5-10 %, three another +2.5-4 %)."""
        if type(cpu_workers) != int or type(gpu_workers) != int:
            raise ValueError("The number of processes have to be an intager")
        if cpu_workers < 1:
            raise ValueError("Minimal number of cpu workers is 1")
        if gpu_workers < 0:
            raise ValueError("Minimal number of gpu workers is 0")
        self.cpu_workers = cpu_workers
        self.gpu_workers = max(1, gpu_workers)
        # host mesh-generation pool: worker PROCESSES (mesh generation is Python/NumPy/Qhull and would serialise on the GIL
        # in threads), forked before any CUDA context exists in this process; they never touch the GPU
        self._mesh_pool = multiprocessing.get_context("fork").Pool(cpu_workers)
        try:
            # fails loudly without a B200
            # `devices`: explicit CUDA device numbers (one rank per GPU under torchrun passes [LOCAL_RANK]); default 0..gpu_workers-1
            devs = list(devices) if devices is not None else list(range(self.gpu_workers))
            self._contexts = [_cabi.Context(d) for d in devs for _ in range(max(1, int(contexts_per_gpu)))]
        except Exception:
            self._mesh_pool.terminate()
            self._mesh_pool = None
            raise

    def shutdown_workers(self):
        for c in self._contexts or []:
            c.close()
        self._contexts = None
        if getattr(self, "_mesh_pool", None) is not None:
            self._mesh_pool.close()
            self._mesh_pool.join()
            self._mesh_pool = None

    # ---- simulation (remo3d.py:723-884)
    def simulate_logs(self, measurement_depths, domain_radius=50, batch_size=5, mesh_generator="auto", preconditioner="multigrid",
                      condense=True, order=3, rtol=1e-10, maxit=None, mesh_options=None, results_log=None, resume=False,
                      share_geometry=True, task_shard=None):
        """`remo3d.py:723-884`.  Additions to the reference's keywords: `order` (the reference hard-codes 3), `rtol`, `maxit`
        (None: 1000 for "multigrid" like `ngsolve_functions.py:50`, 20000 for "local"), `task_shard=(rank, world)` (this
        process solves every world-th task and the [depth, tool, Ra] triples are gathered over torch.distributed at the end,
        the reference's single MPI gather, `remo3d.py:865`), `mesh_options` (mesh sizes; with
        `mesh_generator="gmsh"` a `"msh_path"` entry reads every task's mesh from a Gmsh MSH 2.2 file), `results_log` (a
        JSON-lines file that receives every finished task; with `resume=True` the tasks already in it are not solved
        again), `share_geometry` (3D with `mesh_options={"conforming": False}`: tasks with the same electrode pattern share one
        triangulation and the material is taken per tet centroid; the default 3D meshes carry the interfaces and are per task).
        Per-task failures (meshing or solving) give NaN log points and an `error` entry in `task_records`."""
        start_time = datetime.datetime.now()
        measurement_depths = np.asarray(measurement_depths, dtype=float)
        domain_radius_alert = False
        for tool in self.tools.keys():
            reach = np.max(np.abs(self.tools[tool][0, :3]))
            if reach > domain_radius:
                raise ValueError("Some electrodes are locate outside the simulation domain. Domain size have to be increased")
            elif reach > 0.75 * domain_radius:
                domain_radius_alert = True
        if domain_radius_alert:
            print("Some electrodes are located close to the boundary of the simulation domain. This may cause problems during simulation. Consider increase of the domain size")
        if mesh_generator == "auto":
            mesh_generator = "netgen" if np.isclose(self.dip_deg, 0) else "gmsh"
        if mesh_generator not in ("netgen", "gmsh"):
            raise ValueError("mesh_generator must be 'auto', 'netgen' or 'gmsh'")
        if ~np.isclose(self.dip_deg, 0) and mesh_generator != "gmsh":
            raise ValueError("The only mesh generator supported in 3D models is gmsh")
        if preconditioner not in _cabi.PRECOND:
            raise ValueError("preconditioner must be 'local' or 'multigrid'")
        if type(batch_size) != int or batch_size < 1:
            raise ValueError("batch_size must be a positive integer")
        if (mesh_options or {}).get("msh_path") is not None and mesh_generator != "gmsh":
            raise ValueError("mesh_options['msh_path'] needs mesh_generator='gmsh'")
        if self._contexts is None:
            raise RuntimeError("initialize_workers() must be called before simulate_logs()")
        borehole_model = self.borehole_model
        if self.dip_deg != 0:
            borehole_model = model_io.densify_borehole(self.borehole_model)

        simulation_depths, task_list = self._prepare_simulation_depths_and_tasks(measurement_depths, batch_size)
        n_tasks = len(task_list)
        borehole_geometry = np.ascontiguousarray(borehole_model[:, :2])
        mud_resistivities = np.interp(simulation_depths, borehole_model[:, 0], borehole_model[:, 2])
        print("{} simulation tasks prepared".format(n_tasks))

        triples, records, lock = [], [None] * n_tasks, threading.Lock()
        # ---- resume: tasks already in the results log are taken from it (same plan: same depths, tools, batch size)
        done = set()
        plan_id = "%d tasks, %d depths, tools %s, batch %d, order %d, R %g" % (n_tasks, measurement_depths.shape[0], list(self.tools.keys()),
                                                                             batch_size, order, domain_radius)
        log_file = None
        if results_log is not None:
            if resume and os.path.exists(results_log):
                with open(results_log) as f:
                    for line in f:
                        try:
                            rec = json.loads(line)
                        except ValueError:
                            continue  # a line cut off by the crash that made the resume necessary
                        if rec.get("plan") == plan_id and 0 <= rec.get("task", -1) < n_tasks and "error" not in rec.get("record", {}):
                            if rec["task"] not in done:
                                done.add(rec["task"])
                                triples.extend(rec["triples"])
                                records[rec["task"]] = rec["record"]
                print("{} tasks taken from {}".format(len(done), results_log))
            log_file = open(results_log, "a" if resume else "w")
        todo = [i for i in range(n_tasks) if i not in done]
        three_d = not np.isclose(self.dip_rad, 0.0)
        # 3D meshes carry the material interfaces (`mesh_options["conforming"]`, default True: the reference's Gmsh path
        # fragments the domain by them, `gmsh_functions.py:576-624`), so they depend on the depth of the batch; with
        # conforming=False the material is taken per tet centroid and one triangulation serves every task of a pattern
        conforming = bool((mesh_options or {}).get("conforming", True))
        use_shared = three_d and share_geometry and not conforming and (mesh_options or {}).get("msh_path") is None
        if task_shard is not None:
            keyof = None
            if use_shared:
                keyof = {i: model_mesh.geometry_key(self.dip_rad, task_list[i][1][0], domain_radius, mesh_options) for i in todo}
            todo = worker.shard_tasks(todo, n_tasks, int(task_shard[0]), int(task_shard[1]), keyof)

        mesh_opts_task = dict(mesh_options or {})
        mesh_opts_task.setdefault("conforming", True)

        def job_args(i, geometry=None):
            t = task_list[i]
            return (self.formation_model, borehole_geometry, self.dip_rad, simulation_depths[t[0]], t[1][0], mud_resistivities[t[0]],
                    domain_radius, mesh_opts_task, geometry, i)

        jobs = queue.Queue(maxsize=2 * len(self._contexts) + 2)
        stop = threading.Event()
        busy = [0.0] * len(self._contexts)

        def gpu_loop(k, ctx):
            def feed():
                while not stop.is_set():
                    try:
                        job = jobs.get(timeout=0.2)
                    except queue.Empty:
                        continue
                    if job is None:
                        return
                    yield job
            t0 = [0.0]

            def timed_feed():
                for job in feed():
                    t0[0] = datetime.datetime.now().timestamp()
                    yield job
            for index, t, rec in worker.run_tasks(ctx, timed_feed(), self.tools, order, preconditioner, rtol, maxit):
                busy[k] += datetime.datetime.now().timestamp() - t0[0]
                with lock:
                    triples.extend(t)
                    records[index] = rec
                    if log_file is not None:
                        log_file.write(json.dumps({"plan": plan_id, "task": index, "triples": t, "record": rec}) + "\n")
                        log_file.flush()

        threads = [threading.Thread(target=gpu_loop, args=(k, c), daemon=True) for k, c in enumerate(self._contexts)]
        for th in threads:
            th.start()

        def put(item):
            """Bounded put that notices dead consumers instead of blocking forever."""
            while True:
                try:
                    jobs.put(item, timeout=0.5)
                    return
                except queue.Full:
                    if not any(th.is_alive() for th in threads):
                        raise RuntimeError("all GPU worker threads died")

        t_mesh0 = datetime.datetime.now().timestamp()
        mesh_seconds = 0.0
        try:
            if use_shared:
                # one triangulation per electrode pattern, built ahead by the pool (largest groups first); the per-task part
                # (material of every tet at the task's depth) is cheap and runs here while the GPUs solve
                groups = {}
                for i in todo:
                    groups.setdefault(model_mesh.geometry_key(self.dip_rad, task_list[i][1][0], domain_radius, mesh_options), []).append(i)
                keys = sorted(groups, key=lambda k: -len(groups[k]))
                geo_args = [(task_list[groups[k][0]][1][0], domain_radius, mesh_options) for k in keys]
                for k, geometry in zip(keys, self._mesh_pool.imap(_build_geometry_job, geo_args, chunksize=1)):
                    for i in groups[k]:
                        if isinstance(geometry, BaseException):
                            put((i, task_list[i], geometry, None))
                            continue
                        tm = datetime.datetime.now().timestamp()
                        mesh, sigma = _build_mesh_job(job_args(i, geometry))
                        mesh_seconds += datetime.datetime.now().timestamp() - tm
                        put((i, task_list[i], mesh, sigma))
            else:
                # meshes are produced ahead by the process pool, in task order, while the GPU workers solve
                for i, (mesh, sigma) in zip(todo, self._mesh_pool.imap(_build_mesh_job, [job_args(i) for i in todo], chunksize=1)):
                    put((i, task_list[i], mesh, sigma))
            for _ in threads:
                put(None)
            for th in threads:
                th.join()
        finally:
            # whatever happened, no worker thread may still be inside a context when the caller closes it
            stop.set()
            for th in threads:
                th.join()
            if log_file is not None:
                log_file.close()

        wall = datetime.datetime.now().timestamp() - t_mesh0
        if task_shard is not None and int(task_shard[1]) > 1:
            triples = worker.gather_results(triples, int(task_shard[1]))
        self.logs = worker.results_to_logs(triples, self.tools, measurement_depths)
        self.task_records = records
        self.pipeline_stats = {"tasks": len(todo), "wall_s": wall, "gpu_busy_s": list(busy),
                               "gpu_busy_fraction": (sum(busy) / (len(busy) * wall)) if wall > 0 else 0.0,
                               "host_material_s": mesh_seconds, "shared_geometry": bool(len(todo)) and use_shared}
        print("\nProcessed in: ", datetime.datetime.now() - start_time)

    # ---- results (text part of remo3d.py:902-990)
    def save_results(self, output_folder=None, measurements_to_save="auto", **_plot_options):
        """Writes Results_<n>.txt exactly like the reference (`remo3d.py:957-990`); plotting is out of scope."""
        if output_folder is None:
            return None
        sub = os.path.join(output_folder, "Results_{}/".format(datetime.datetime.now().strftime("%Y_%m_%d__%H_%M_%S")))
        os.makedirs(sub, exist_ok=True)
        todo = list(self.logs.keys()) if measurements_to_save == "auto" else list(measurements_to_save)
        number = 1
        while todo:
            group = [todo[0]]
            for name in todo[1:]:
                a, b = self.logs[todo[0]][:, 0], self.logs[name][:, 0]
                if a.shape[0] == b.shape[0] and np.all(np.isclose(a, b)):
                    group.append(name)
            for name in group:
                todo.remove(name)
            table = self.logs[group[0]]
            for name in group[1:]:
                table = np.hstack([table, np.atleast_2d(self.logs[name][:, 1]).T])
            header = "\t".join(["DEPTH"] + group) + "\n" + "\t".join(["M"] + ["OHMM"] * len(group))
            np.savetxt(sub + "Results_{}.txt".format(number), table, fmt="%.4f", delimiter="\t", header=header, comments="")
            number += 1
        return sub
