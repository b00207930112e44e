"""Worker of test_two_rank_gloo_exchange: the rank-local halo lists drive a real send/recv exchange (gloo, CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from qgdsolver_b200 import decompose  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mesh = cases.pm.hex_box(8, 6, 5, perturb=0.2, seed=3)
cell_rank = decompose.geometric_split(mesh, world)
sd = decompose.extended_submeshes(mesh, cell_rank, ranks=[rank])[0]
field = np.sin(np.arange(mesh.n_cells) * 0.37) + 2.0          # "global truth"
local = np.full(sd.mesh.n_cells, np.nan)
local[:sd.n_owned] = field[sd.cell_global[:sd.n_owned]]
reqs, bufs = [], {}
for s, ids in sd.send_cells.items():
    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(local[ids])), s))
for s, ids in sd.recv_cells.items():
    bufs[s] = torch.empty(ids.size, dtype=torch.float64)
    reqs.append(dist.irecv(bufs[s], s))
for r in reqs:
    r.wait()
for s, ids in sd.recv_cells.items():
    local[ids] = bufs[s].numpy()
assert np.array_equal(local, field[sd.cell_global]), "halo exchange mismatch"
# global reduction stand-in for the Courant all-reduce
t = torch.tensor([float(local[:sd.n_owned].max())], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert abs(t.item() - field.max()) < 1e-15
print("HALO_OK", rank, flush=True)
dist.destroy_process_group()
