"""`Model`: drop-in for `/root/reference/remo3d/remo3d.py` `class Model` on the forward-solve path.

Same constructor, same method names and keyword arguments, same `logs` layout and the same ValueError messages;
the MPI worker farm (`remo3d.py:552-599, 809-899`) is replaced by an in-process queue: a host thread pool builds
meshes (`cpu_workers`), one worker thread per GPU (`gpu_workers`) solves them.  Extra keyword arguments that the
reference hard-codes: `order` (3, ngsolve_functions.py:27), `rtol` (CG tolerance), `mesh_options`.
Plotting (`save_results` figures) is out of scope; the text writer keeps the reference's file format.
"""
import datetime
import multiprocessing
import os
import queue
import threading

import numpy as np

from . import _cabi, model_io, model_mesh, planner, tools as tl, worker


def _build_mesh_job(args):
    """Runs in a mesh-pool process: one task's mesh + sigma list (pure host work)."""
    return model_mesh.build_task_mesh(*args)


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
    def initialize_workers(self, cpu_workers=4, gpu_workers=0, contexts_per_gpu=2):
        """`gpu_workers` = number of GPUs to shard the mesh tasks over (0 is promoted to 1: there is no CPU solve
        path in this package); `cpu_workers` = host processes that build meshes ahead of the GPUs.  Every GPU runs
        `contexts_per_gpu` solver contexts (own stream + host thread): mesh tasks are independent, and two in flight
        hide the launch-bound parts of one another (measured +10 % throughput at 4.8 M dofs)."""
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
            self._contexts = [_cabi.Context(d) for d in range(self.gpu_workers) for _ in range(max(1, int(contexts_per_gpu)))]
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
                      condense=True, order=3, rtol=1e-10, maxit=1000, mesh_options=None):
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
        if ~np.isclose(self.dip_deg, 0) and mesh_generator != "gmsh":
            raise ValueError("The only mesh generator supported in 3D models is gmsh")
        if preconditioner not in _cabi.PRECOND:
            raise ValueError("preconditioner must be 'local' or 'multigrid'")
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

        job_args = [(self.formation_model, borehole_geometry, self.dip_rad, simulation_depths[t[0]], t[1][0], mud_resistivities[t[0]],
                     domain_radius, mesh_options) for t in task_list]

        jobs = queue.Queue(maxsize=2 * len(self._contexts) + 2)
        triples, records, lock = [], [None] * n_tasks, threading.Lock()

        def gpu_loop(ctx):
            def feed():
                while True:
                    job = jobs.get()
                    if job is None:
                        return
                    yield job
            for index, t, rec in worker.run_tasks(ctx, feed(), self.tools, order, preconditioner, rtol, maxit):
                with lock:
                    triples.extend(t)
                    records[index] = rec

        threads = [threading.Thread(target=gpu_loop, args=(c,), daemon=True) for c in self._contexts]
        for th in threads:
            th.start()
        # meshes are produced ahead by the process pool, in task order, while the GPU workers solve
        for i, (mesh, sigma) in enumerate(self._mesh_pool.imap(_build_mesh_job, job_args, chunksize=1)):
            jobs.put((i, task_list[i], mesh, sigma))
        for _ in threads:
            jobs.put(None)
        for th in threads:
            th.join()

        self.logs = worker.results_to_logs(triples, self.tools, measurement_depths)
        self.task_records = records
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
