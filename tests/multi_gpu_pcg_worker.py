"""Worker of test_decomposed_pcg_matches_oracle: N ranks (one per GPU) solve one global symmetric system with
qgd_pcg_solve_multi (stepwise PCG, NCCL halo exchange of the search direction, all-reduced dot products); rank 0 gathers the
owned parts and compares them with the oracle (or_pcg_solve_blocks; none / diagonal preconditioning do not depend on the blocks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from qgdsolver_b200 import api, decompose, multigpu  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
api.init(local)
multigpu.init_comm(rank, world)

LOG = None
if rank == 0 and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
    LOG = open(os.path.join(ROOT, "gpurun_out", f"multi_pcg_parity_n{world}.log"), "w")


def say(msg):
    print(msg, flush=True)
    if LOG:
        LOG.write(msg + "\n"); LOG.flush()


mesh = cases.pm.hex_box(14, 12, 10, perturb=0.1, seed=4)
nI, nC = mesh.n_internal, mesh.n_cells
upper_g = -(mesh.magSf[:nI] * mesh.deltaCoeffs[:nI])
diag_g = np.zeros(nC)
np.subtract.at(diag_g, mesh.owner[:nI], upper_g); np.subtract.at(diag_g, mesh.neighbour, upper_g)
diag_g += 1e-3 * mesh.V / mesh.V.mean()
b_g = np.random.default_rng(0).standard_normal(nC)
cell_rank = decompose.geometric_split(mesh, world)
sub = decompose.extended_submeshes(mesh, cell_rank, ranks=[rank])[0]
dm = api.Mesh(sub.mesh, n_owned=sub.n_owned, coupled_face=sub.coupled_face)
cg = sub.cell_global
ok = True
# DIC in a decomposed run is block-local: tiles inside every rank's owned cells (qgd_mesh_make_pcg_blocks); the oracle gets the
# same map in global numbering (rank r's blocks follow those of ranks < r)
blk_local = dm.make_pcg_blocks(60)[:sub.n_owned]
np.save(f"/tmp/qgd_multi_pcg_blk_{rank}.npy", blk_local)
dist.barrier()
for precond in ("diagonal", "none", "DIC"):
    x, it, r0, r1 = api.pcg_solve_multi(dm, sub, diag_g[cg], upper_g[sub.face_global[:sub.mesh.n_internal]], b_g[cg], np.zeros(cg.size),
                                        tol=1e-12, max_iter=3000, precond=precond)
    np.save(f"/tmp/qgd_multi_pcg_{rank}.npy", x[:sub.n_owned])
    api.synchronize()
    dist.barrier()
    if rank == 0:
        import oracle as O
        o = O.Oracle(mesh)
        blocks = cell_rank
        if precond == "DIC":
            blocks, base = np.zeros(nC, np.int32), 0
            for r in range(world):
                s_r = decompose.extended_submeshes(mesh, cell_rank, ranks=[r])[0]
                b_r = np.load(f"/tmp/qgd_multi_pcg_blk_{r}.npy")
                blocks[s_r.cell_global[:s_r.n_owned]] = base + b_r
                base += int(b_r.max()) + 1
        xs, its, r0s, r1s = o.pcg_solve(diag_g, upper_g, b_g, np.zeros(nC), tol=1e-12, maxIter=3000, precond=O.PRECONDS[precond], cell_block=blocks)
        got = np.zeros(nC)
        for r in range(world):
            s_r = decompose.extended_submeshes(mesh, cell_rank, ranks=[r])[0]
            got[s_r.cell_global[:s_r.n_owned]] = np.load(f"/tmp/qgd_multi_pcg_{r}.npy")
        err = float(np.abs(got - xs).max() / np.abs(xs).max())
        good = err < 1e-9 and abs(it - its) <= 1 and abs(r0 - r0s) < 1e-10 * r0s
        ok = ok and good
        say(f"MPCG n={world} {precond} iterations {it} (oracle {its}) relLinf={err:.3e} {'ok' if good else 'FAIL'}")
    # the halo entries of the returned solution are the owners' values
    full = np.zeros(nC)
    dist.barrier()
if rank == 0:
    say("MPCG_ALL_OK" if ok else "MPCG_FAILED")
api.comm_finalize()
dist.destroy_process_group()
