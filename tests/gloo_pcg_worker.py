"""Worker of test_two_rank_gloo_pcg: PCG on a decomposed matrix with real message passing (gloo, CPU).  Each rank holds
the rows of its owned cells; every iteration exchanges the search direction on the face-neighbour halo
(SubDomain.recv_face_cells / send_face_cells) and all-reduces three scalars; the preconditioner is DIC on the rank's own
block.  Rank 0 compares the gathered solution and the iteration count with the oracle's decomposed-run solver
(or_pcg_solve_blocks).  This is the algorithm the multi-GPU pressure solver has to implement."""
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
mesh = cases.pm.hex_box(9, 8, 6, perturb=0.1, seed=4)
nI, nC = mesh.n_internal, mesh.n_cells
upper_g = -(mesh.magSf[:nI] * mesh.deltaCoeffs[:nI])
diag_g = np.zeros(nC)
np.subtract.at(diag_g, mesh.owner[:nI], upper_g); np.subtract.at(diag_g, mesh.neighbour, upper_g)
diag_g += 1e-3 * mesh.V / mesh.V.mean()
b_g = np.random.default_rng(0).standard_normal(nC)
cell_rank = decompose.geometric_split(mesh, world)
sd = decompose.extended_submeshes(mesh, cell_rank, ranks=[rank])[0]
m, nO = sd.mesh, sd.n_owned
nIl = m.n_internal
own, nei = m.owner[:nIl], m.neighbour
up = upper_g[sd.face_global[:nIl]]
rows = own < nO                                   # faces with an owned owner: owned-owned and owned-halo (coupled)
inblk = nei < nO                                  # ... of which both cells are owned: the rank's own block
diag, b = diag_g[sd.cell_global[:nO]], b_g[sd.cell_global[:nO]]


def allsum(*vals):
    t = torch.tensor(vals, dtype=torch.float64)
    dist.all_reduce(t)
    return t.tolist()


def exchange(v):
    """v: local vector (owned + halo); fills the face-neighbour halo entries from their owners"""
    reqs, bufs = [], {}
    for s, ids in sd.send_face_cells.items():
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(v[ids])), s))
    for s, ids in sd.recv_face_cells.items():
        bufs[s] = torch.empty(ids.size, dtype=torch.float64)
        reqs.append(dist.irecv(bufs[s], s))
    for r in reqs:
        r.wait()
    for s, ids in sd.recv_face_cells.items():
        v[ids] = bufs[s].numpy()


def amul(v):
    exchange(v)
    y = diag * v[:nO]
    f = np.nonzero(rows)[0]
    np.add.at(y, own[f], up[f] * v[nei[f]])
    g = f[inblk[f]]
    np.add.at(y, nei[g], up[g] * v[own[g]])
    return y


x = np.zeros(m.n_cells)
wA = amul(x)
rA = b - wA
(sx, n) = allsum(float(x[:nO].sum()), float(nO))
xRef = sx / n
sumA = diag.copy()
f = np.nonzero(rows)[0]
np.add.at(sumA, own[f], up[f])
g = f[inblk[f]]
np.add.at(sumA, nei[g], up[g])
(normFactor,) = allsum(float((np.abs(wA - sumA * xRef) + np.abs(b - sumA * xRef)).sum()))
normFactor += 1e-20
(res,) = allsum(float(np.abs(rA).sum()))
res /= normFactor
tol, it = 1e-13, 0
blk = np.nonzero(rows & inblk)[0]                 # ascending local order == ascending global face order inside the block
rD = diag.copy()
for fi in blk:
    rD[nei[fi]] -= up[fi] * up[fi] / rD[own[fi]]
rD = 1.0 / rD
pA = np.zeros(m.n_cells)
wArA = 1e20
while res >= tol and it < 3000:
    wArAold = wArA
    w = rD * rA
    for fi in blk:
        w[nei[fi]] -= rD[nei[fi]] * up[fi] * w[own[fi]]
    for fi in blk[::-1]:
        w[own[fi]] -= rD[own[fi]] * up[fi] * w[nei[fi]]
    (wArA,) = allsum(float((w * rA).sum()))
    pA[:nO] = w if it == 0 else w + (wArA / wArAold) * pA[:nO]
    wA = amul(pA)
    (wApA,) = allsum(float((wA * pA[:nO]).sum()))
    alpha = wArA / wApA
    x[:nO] += alpha * pA[:nO]
    rA -= alpha * wA
    (res,) = allsum(float(np.abs(rA).sum()))
    res /= normFactor
    it += 1
np.save(f"/tmp/qgd_gloo_pcg_{rank}.npy", x[:nO])
dist.barrier()
if rank == 0:
    import oracle as O
    o = O.Oracle(mesh)
    xs, its, r0, r1 = o.pcg_solve(diag_g, upper_g, b_g, np.zeros(nC), tol=tol, maxIter=3000, precond=2, cell_block=cell_rank)
    got = np.zeros(nC)
    for r in range(world):
        sub = decompose.extended_submeshes(mesh, cell_rank, ranks=[r])[0]
        got[sub.cell_global[:sub.n_owned]] = np.load(f"/tmp/qgd_gloo_pcg_{r}.npy")
    err = float(np.abs(got - xs).max() / np.abs(xs).max())
    assert err < 1e-10, err
    assert abs(it - its) <= 1, (it, its)
    print("PCG_OK iterations", it, "oracle", its, "err", err, flush=True)
dist.barrier()
dist.destroy_process_group()
