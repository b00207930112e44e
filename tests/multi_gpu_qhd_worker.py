"""Worker of the multi-GPU QHDFoam parity test: N ranks (one per GPU) step the decomposed heated-cavity case through the
C-ABI (qgd_qhdfoam_set_halo: vertex-ring state exchange, stepwise PCG with NCCL exchange of the search direction and
all-reduced dot products, all-reduced reference shift / Courant number); rank 0 gathers the owned parts and compares them
with the serial CPU oracle (diagonal / no preconditioning are decomposition-independent)."""
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
    LOG = open(os.path.join(ROOT, "gpurun_out", f"multi_qhd_parity_n{world}.log"), "w")


def say(msg):
    print(msg, flush=True)
    if LOG:
        LOG.write(msg + "\n"); LOG.flush()


def fixed_p_case():
    c = cases.qhd_cavity(n=(20, 16), perturb=0.1, p_bc="zg", precond="diagonal", tol=1e-13, max_iter=20000)
    names = [p.name for p in c.mesh.patches]
    c.bcP[names.index("yMax")] = cases.FV           # a fixedValue patch: no reference cell (p.needReference() false on every rank)
    return c


CASES = {
    "cavity2d_diag": lambda: cases.qhd_cavity(n=(24, 20), precond="diagonal", tol=1e-13, max_iter=20000),
    "cavity2d_perturbed_refcell": lambda: cases.qhd_cavity(n=(20, 18), perturb=0.15, precond="none", tol=1e-13, max_iter=20000,
                                                           p_ref_cell=77, p_ref_value=0.3),
    "cavity3d_H2bynu": lambda: cases.qhd_cavity(n=(10, 9, 8), dims=3, perturb=0.1, model="H2bynuQHD", precond="diagonal", tol=1e-13,
                                                max_iter=20000),
    "cavity2d_HbyU_adjust": lambda: cases.qhd_cavity(n=(20, 16), perturb=0.1, model="HbyUQHD", precond="diagonal", tol=1e-13,
                                                     max_iter=20000, adjust_time_step=True, max_co=0.05),
    "cavity2d_fixed_p_patch": fixed_p_case,
}
UNIFORM = {"cavity2d_diag"}          # processor-patch rule kept: identical to the serial run on a uniform mesh
ok = True
NSTEPS = 30
for name, mk in CASES.items():
    c = mk()
    cell_rank = decompose.geometric_split(c.mesh, world)
    sub = decompose.extended_submeshes(c.mesh, cell_rank, ranks=[rank])[0]
    if name not in UNIFORM:
        # hQGDf on processor faces is |d| in the reference (QGDCoeffs.C:195-199) and 2 min(|C_P - C_f|, |C_N - C_f|) on internal
        # faces (:303-308): on a non-uniform mesh a decomposed reference run differs from the serial one wherever tau depends on
        # hQGD (H2bynuQHD, HbyUQHD).  The comparison with the SERIAL oracle therefore keeps the serial rule on the cut faces.
        sub.coupled_face[:] = 0
    dm = api.Mesh(sub.mesh, n_owned=sub.n_owned, coupled_face=sub.coupled_face)
    cg = sub.cell_global
    g2l = np.full(c.mesh.n_cells, -1, np.int64)
    g2l[cg[:sub.n_owned]] = np.arange(sub.n_owned)
    ds = c.diff_solver
    s = api.QHDFoam(dm, fvsc_scheme=c.scheme, qgd_coeffs=c.model, delta_t=c.dt, p_ref_cell=int(g2l[c.p_ref_cell]),
                    p_ref_value=c.p_ref_value, **c.fluid, **c.coeffs, **c.solver, **c.opts)
    nI_g = c.mesh.n_internal
    bf_g = sub.face_global[sub.mesh.n_internal:]
    phys = bf_g >= nI_g
    idx = np.where(phys, bf_g - nI_g, 0)
    pad = lambda k: np.concatenate([np.asarray(k, np.int32), [1]]).astype(np.int32)
    s.set_bcs(pad(c.bcU), pad(c.bcT), pad(c.bcP), np.where(phys[:, None], c.bvU[idx], 0.0), np.where(phys, c.bvT[idx], 0.0),
              np.where(phys, c.bvP[idx], 0.0))
    s.set_halo(sub)
    s.init_fields(c.U0[cg], c.T0[cg], c.p0[cg], None if c.alphaQGD is None else c.alphaQGD[cg])
    s.step(NSTEPS)
    res = {f: s.get(f)[:sub.n_owned] for f in ("U", "T", "p")}
    np.savez(f"/tmp/qgd_multi_qhd_{name}_{rank}.npz", cells=cg[:sub.n_owned], dt=s.scalars()["deltaT"], iters=s.solver_info()["iters"], **res)
    api.synchronize()
    dist.barrier()
    if rank == 0:
        import oracle as O
        o = c.make_oracle(O)
        c.oracle_step(o, NSTEPS)
        for f in ("U", "T", "p"):
            ref = o.qhd_get(f)
            got = np.zeros_like(ref)
            for r in range(world):
                z = np.load(f"/tmp/qgd_multi_qhd_{name}_{r}.npz")
                got[z["cells"]] = z[f]
            err = float(np.abs(got - ref).max() / np.abs(ref).max())
            good = err < 1e-10
            ok = ok and good
            say(f"MULTIQHD n={world} {name} steps={NSTEPS} {f} relLinf={err:.3e} {'ok' if good else 'FAIL'}")
        z = np.load(f"/tmp/qgd_multi_qhd_{name}_0.npz")
        it_o = o.qhd_solver_info()["iters"]
        say(f"MULTIQHD n={world} {name} pcg iterations {int(z['iters'])} (oracle {it_o})")
        if c.opts["adjust_time_step"]:
            derr = abs(float(z["dt"]) - o.qhd_deltaT()) / o.qhd_deltaT()
            say(f"MULTIQHD n={world} {name} deltaT rel err={derr:.3e} {'ok' if derr < 1e-10 else 'FAIL'}")
            ok = ok and derr < 1e-10
    dist.barrier()
if rank == 0:
    say("MULTIQHD_ALL_OK" if ok else "MULTIQHD_FAILED")
api.comm_finalize()
dist.destroy_process_group()
