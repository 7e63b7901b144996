import os, sys, time, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from remo3d_b200 import meshgen
from remo3d_b200.mesh import Mesh
from mesh_quality_iters import iters
size, improve = sys.argv[1], int(sys.argv[2])
qmin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.12
task, flat = bench.make_task()
he, ha, g, hm = bench.SIZES[size]
material = meshgen.layered_material([-1.0, 1.5], dip_rad=np.deg2rad(30.0), borehole_radius=0.1, inclusion=((3.0, 2.0, 1.0), 1.5))
t0 = time.time()
m = meshgen.half_ball_mesh(50.0, task[1][0], material=material, h_electrode=he, h_axis=ha, grading=g, h_max=hm, seed=0, improve=improve, improve_quality=qmin, improve_mode=(sys.argv[4] if len(sys.argv) > 4 else "normal"))
print("mesh %.0fs" % (time.time() - t0))
mesh = Mesh(m["points"], m["elems"], m["mat"], m["bfacets"], m["bc"], m["bc_names"])
m["bdir"] = mesh.dirichlet_flags("dirichlet_boundary")
iters(m, flat, "improve=%d qmin=%g" % (improve, qmin))
