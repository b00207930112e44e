"""A/B sweep of an environment switch read at solver creation / step time, on one mesh build:
python gpu_envsweep.py N steps VAR v0,v1,...   (e.g. QGD_BND_FORK 0,1,0,1)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as B
from qgdsolver_b200 import api

n, steps, var = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
vals = sys.argv[4].split(",")
api.init(0)
c = B.build_case(n)
dmesh = api.Mesh(c.mesh)
ref = None
for v in vals:
    os.environ[var] = v
    s = c.make_solver(api, dmesh)
    s.step(5)
    api.synchronize()
    api.timer_begin()
    s.step(steps)
    ms = api.timer_end() / steps
    s.profile(True)
    s.step(10)
    kt = s.kernel_times()
    s.profile(False)
    r = s.get("rhoE")
    if ref is None:
        ref = r
    print(json.dumps({var: v, "ms_per_step": ms, "mcups": c.mesh.n_cells / ms / 1e3, "points_ms": kt["points_ms"] / kt["steps"],
                      "face_ms": kt["face_ms"] / kt["steps"], "cell_ms": kt["cell_ms"] / kt["steps"],
                      "bitwise_equal_to_first": bool(np.array_equal(r, ref))}), flush=True)
    del s
