import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases, oracle as O
from qgdsolver_b200 import api
api.init(0)
which, pipe, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mk = {"poly": lambda: cases.case_poly(bcs='qgdflux'), "impl": lambda: cases.case_hex3d(bcs='fixed', implicit=True),
      "polyzg": lambda: cases.case_poly(bcs='zg')}[which]
c = mk(); o = c.make_oracle(O); s = c.make_solver(api)
if pipe: s.set_pipeline(1, 32, 1, 0)
c.oracle_step(o, n); s.step(n)
print(which, "MAXW", os.environ.get("QGD_ELL_MAXW"), "pipe", pipe, "n", n, {f: float(np.abs(s.get(f) - o.get(f)).max() / np.abs(o.get(f)).max()) for f in ('rho', 'rhoU', 'rhoE')},
      s.diffusion_iterations() if c.implicit else "")
