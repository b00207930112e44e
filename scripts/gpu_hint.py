"""L2 cache-policy sweep of the TMA face kernel (QGD_FACE_L2HINT = 0..3) on one mesh build: python gpu_hint.py N [steps]."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as B
from qgdsolver_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hints = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3, 0]
api.init(0)
c = B.build_case(n)
dmesh = api.Mesh(c.mesh)
ref = None
for h in hints:
    os.environ["QGD_FACE_L2HINT"] = str(h)
    s = c.make_solver(api, dmesh)
    s.step(5)
    api.synchronize()
    s.profile(True)
    api.timer_begin()
    s.step(steps)
    ms = api.timer_end() / steps
    kt = s.kernel_times()
    s.profile(False)
    r = s.get("rhoE")
    if ref is None:
        ref = r
    print(json.dumps({"l2hint": h, "ms_per_step": ms, "mcups": c.mesh.n_cells / ms / 1e3, "points_ms": kt["points_ms"] / kt["steps"],
                      "face_ms": kt["face_ms"] / kt["steps"], "cell_ms": kt["cell_ms"] / kt["steps"],
                      "bitwise_equal_to_first": bool(np.array_equal(r, ref))}), flush=True)
    del s
