"""GPU probe: per-step stage times of the bench step under different host-side conditions."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from remo3d_b200 import _cabi
size = sys.argv[1] if len(sys.argv) > 1 else "5M"
task, flat = bench.make_task()
m = bench.make_mesh(size, task, print)
names = ["points", "elems", "mat", "bfacets", "bdir", "axis"]
host = {k: torch.from_numpy(np.ascontiguousarray(m[k])).pin_memory() for k in names}
dev = {k: host[k].cuda() for k in names}
ctx = _cabi.Context(0)
def step(a):
    t0 = time.time()
    ctx.mesh_set(3, a["points"], a["elems"], a["mat"], a["bfacets"], a["bdir"], a["axis"]); t1 = time.time()
    ctx.space_build(2); t2 = time.time()
    ctx.assemble(bench.SIGMA); t3 = time.time()
    ctx.precond_setup("multigrid"); t4 = time.time()
    ctx.rhs_point_sources(flat["src_ptr"], flat["src_z"], flat["src_fac"]); t5 = time.time()
    ctx.solve(rtol=1e-10, maxit=3000); t6 = time.time()
    ctx.apparent_resistivity(flat["pt_rhs"], flat["pt_z0"], flat["pt_z1"], flat["pt_k"], flat["scale"]); t7 = time.time()
    s = ctx.stage_times()
    print("host ms: mesh %.1f space %.1f asm %.1f pre %.1f rhs %.1f solve %.1f ra %.1f | gpu ms: space %.1f pre %.1f solve %.1f" % (
        1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), 1e3*(t6-t5), 1e3*(t7-t6), s["space_build"], s["precond_setup"], s["solve"]))
print("-- own stream, dev arrays"); [step(dev) for _ in range(3)]
print("-- profile on"); ctx.profile(True); [step(dev) for _ in range(3)]; ctx.profile(False)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
print("-- torch stream"); [step(dev) for _ in range(3)]
print("-- torch stream, host arrays"); [step(host) for _ in range(2)]
