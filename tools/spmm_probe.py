"""GPU probe: SpMM / vector-update / assembly kernel times and achieved GB/s on a bench mesh."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from remo3d_b200 import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--size", default=os.environ.get("REMO_PROBE_SIZE", "1M"))
ap.add_argument("--order", type=int, default=2)
ap.add_argument("--ks", default="1,5,8")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
task, flat = bench.make_task()
m = bench.make_mesh(a.size, task, print)
ctx = _cabi.Context(0)
for kv in os.environ.get("REMO_BENCH_OPTS", "").split(","):
    if "=" in kv:
        ctx.set_option(kv.split("=")[0].strip(), float(kv.split("=")[1]))
ctx.mesh_set(3, m["points"], m["elems"], m["mat"], m["bfacets"], m["bdir"], m["axis"])
ndof, _ = ctx.space_build(a.order); nnz = ctx.nnz
ctx.assemble(bench.SIGMA)
ctx.precond_setup("local")
print("ndof", ndof, "nnz", nnz, "variant", os.environ.get("REMO_SPMM_VARIANT"), ctx.stage_times())
for k in [int(x) for x in a.ks.split(",")]:
    ptr = np.arange(k + 1, dtype=np.int64)
    z = np.resize(flat["src_z"], k)
    ctx.rhs_point_sources(ptr, z, np.ones(k))
    ctx.solve(rtol=1e-3, maxit=64, raise_on_noconv=False)   # fills P with a realistic dense vector
    ms = ctx.kernel_time(0, k, a.reps)
    mv = ctx.kernel_time(2, k, a.reps)
    b = bench.spmm_bytes(nnz, ndof, k)
    vb = ndof * k * 8 * 11.0
    print("k=%d  spmm %.4f ms  %.0f GB/s (%.1f%% of 6553)   vec-updates %.4f ms %.0f GB/s" % (k, ms, b / ms / 1e6, b / ms / 1e6 / 65.533, mv, vb / mv / 1e6))
ms = ctx.kernel_time(1, 1, 5)
print("assembly kernels %.3f ms" % ms)
