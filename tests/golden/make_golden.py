"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).
The reference cannot be executed here (no OpenFOAM), so these are regression vectors of the oracle restatement,
small enough to commit; they also serve as fixed inputs/outputs for the GPU parity run on the box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from test_oracle_kat import GOLDEN_CASES, golden_fields  # noqa: E402

only = set(sys.argv[1:])
for name, mk in GOLDEN_CASES.items():
    if only and name not in only:
        continue
    c = mk()
    o = c.make_oracle(O)
    steps = 25
    c.oracle_step(o, steps)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"{name}.npz"), steps=steps, **golden_fields(c, o))
    print(name, "written")
