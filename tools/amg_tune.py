"""GPU probe: PCG iteration counts / solve time of the multigrid preconditioner for several tunables."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from remo3d_b200 import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="1M")
ap.add_argument("--order", type=int, default=2)
ap.add_argument("--combos", default="1:1.0:1.3,2:1.0:1.3,1:1.5:1.3,2:1.5:1.3,2:1.0:1.6,3:1.0:1.3")
a = ap.parse_args()
task, flat = bench.make_task()
m = bench.make_mesh(a.size, task, print)
ctx = _cabi.Context(0)
ctx.mesh_set(3, m["points"], m["elems"], m["mat"], m["bfacets"], m["bdir"], m["axis"])
ndof, _ = ctx.space_build(a.order); nnz = ctx.nnz
ctx.assemble(bench.SIGMA)
ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"])
print("ndof", ndof, "nnz", nnz)
for combo in a.combos.split(","):
    sw, al, om = combo.split(":")
    ctx.set_option("amg_sweeps", float(sw)); ctx.set_option("amg_alpha", float(al)); ctx.set_option("amg_omega_scale", float(om))
    ctx.precond_setup("multigrid")
    it, rel = ctx.solve(rtol=1e-10, maxit=3000, raise_on_noconv=False)
    t = ctx.stage_times()
    print("sweeps=%s alpha=%s omega_scale=%s  iters=%d relres=%.1e solve=%.1f ms  (%.3f ms/iter) setup=%.1f ms" % (sw, al, om, it.max(), rel.max(), t["solve"], t["solve"] / max(1, it.max()), t["precond_setup"]))
