"""Tuning sweep of the pipelined step at size N: python gpu_tune.py N [steps]  (one mesh build, many configurations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from qgdsolver_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
api.init(0)
c = B.build_case(n)
s = c.make_solver(api)
configs = [(0, 0, -1, 0), (1, 0, -1, 0)]
if len(sys.argv) > 3:
    configs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[3:]]
for cfg in configs:
    s.set_pipeline(*cfg)
    s.step(5)
    api.synchronize()
    s.profile(True)
    api.timer_begin()
    s.step(steps)
    ms = api.timer_end() / steps
    kt = s.kernel_times()
    s.profile(False)
    print(json.dumps({"cfg": cfg, "pipe": s.get_pipeline(), "ms_per_step": ms, "mcups": c.mesh.n_cells / ms / 1e3,
                      "points_ms": kt["points_ms"] / kt["steps"], "face_ms": kt["face_ms"] / kt["steps"],
                      "cell_ms": kt["cell_ms"] / kt["steps"]}), flush=True)
