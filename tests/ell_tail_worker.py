"""Worker of test_ell_tails_and_degenerate_sizes: run with QGD_ELL_MAXW=4 in a fresh process so that every stencil row
longer than 4 (cell->faces, point->cells, PCG rows) goes through the CSR tail path of the device layout."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle as O  # noqa: E402
from qgdsolver_b200 import api  # noqa: E402

TOL = 1e-10
api.init(0)
bad = []
for name, mk, pipe in (("hex3d-mixed", lambda: cases.case_hex3d(perturb=0.2, bcs="mixed"), 0),
                       ("hex3d-mixed-pipelined", lambda: cases.case_hex3d(perturb=0.2, bcs="mixed"), 1),
                       ("poly-qgdflux", lambda: cases.case_poly(bcs="qgdflux"), 0),
                       ("poly-qgdflux-pipelined", lambda: cases.case_poly(bcs="qgdflux"), 1),
                       ("hex3d-implicit", lambda: cases.case_hex3d(bcs="fixed", implicit=True), 0)):
    c = mk()
    o = c.make_oracle(O)
    s = c.make_solver(api)
    if pipe:
        s.set_pipeline(1, 32, 1, 0)
    c.oracle_step(o, 30)
    s.step(30)
    err = {f: float(np.abs(s.get(f) - o.get(f)).max() / np.abs(o.get(f)).max()) for f in ("rho", "rhoU", "rhoE")}
    print(name, err, flush=True)
    if not all(v < TOL for v in err.values()):
        bad.append(name)
q = cases.qhd_cavity(n=(7, 6, 5), dims=3, dt=1e-3)
o = q.make_oracle(O)
s = q.make_solver(api)
q.oracle_step(o, 20)
s.step(20)
err = {f: float(np.abs(s.get(f) - o.qhd_get(f)).max() / np.abs(o.qhd_get(f)).max()) for f in ("U", "T", "p")}
print("qhd-cavity-3d", err, flush=True)
if not all(v < TOL for v in err.values()):
    bad.append("qhd-cavity-3d")
print("TAILS_OK" if not bad else "TAILS_BAD %s" % bad, flush=True)
