"""Worker of the multi-GPU parity test: N ranks (one per GPU) step the decomposed case through the C-ABI with the NCCL
halo exchange; rank 0 gathers the owned parts and compares them with the serial CPU oracle."""
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
    LOG = open(os.path.join(ROOT, "gpurun_out", f"multi_parity_n{world}.log"), "w")


def say(msg):
    print(msg, flush=True)
    if LOG:
        LOG.write(msg + "\n"); LOG.flush()


def slip_case():
    c = cases.case_hex3d(n=(12, 10, 8), perturb=0.2, bcs="zg")
    for i in range(1, len(c.mesh.patches), 2):
        c.bcU[i] = cases.SLIP
    return c


CASES = {
    # coupled-face rule off: the decomposed run must reproduce the serial semantics on any mesh
    "perturbed_mixed_serialrule": (lambda: cases.case_hex3d(n=(12, 10, 8), perturb=0.2, bcs="mixed"), False),
    "prism_fixed_serialrule": (lambda: cases.case_prism(n=(6, 5, 4), bcs="fixed"), False),
    "2d_qgdflux_serialrule": (lambda: cases.case_2d((20, 16), perturb=0.15, bcs="qgdflux"), False),
    # reference processor-patch rule (hQGDf = |d| on coupled faces): identical to serial on uniform meshes
    "uniform_zg_procrule": (lambda: cases.case_hex3d(n=(12, 10, 8), bcs="mixed"), True),
    "uniform_adjust_procrule": (lambda: cases.case_hex3d(n=(12, 10, 8), bcs="fixed", adjust_time_step=True, dt=1e-3, max_co=0.1, c_tau=0.3), True),
    # polyhedral cells (14 faces, CSR tails), slip walls, a model with a per-cell ScQGD sensor
    "truncoct_mixed_serialrule": (lambda: cases.case_truncoct(n=(6, 5, 5), bcs="mixed"), False),
    "slip_perturbed_serialrule": (slip_case, False),
    "varSc7_fixed_serialrule": (lambda: cases.case_hex3d(n=(12, 10, 8), perturb=0.1, bcs="fixed", model="varScModel7",
                                                         varsc=dict(cSc1=3.0, minSc=0.02, maxSc=0.4)), False),
    # implicitDiffusion true on sub-meshes: stepwise PCG with NCCL exchange of the search direction / all-reduced dot products
    "perturbed_mixed_implicit_serialrule": (lambda: cases.case_hex3d(n=(12, 10, 8), perturb=0.2, bcs="mixed", implicit=True,
                                                                     diff_solver=dict(precond="diagonal")), False),
    "2d_qgdflux_implicit_adjust_serialrule": (lambda: cases.case_2d((20, 16), perturb=0.15, bcs="qgdflux", implicit=True, adjust_time_step=True,
                                                                    max_co=0.1, diff_solver=dict(precond="none")), False),
    # leastSquares on sub-meshes (2D): the vertex-ring halo holds every stencil cell of the faces this rank evaluates
    "2d_leastSquares_serialrule": (lambda: cases.case_2d((20, 16), perturb=0.2, bcs="mixed", scheme="leastSquares"), False),
    # boundary kernels forked onto the side stream in multi-GPU mode
    "perturbed_mixed_fork2": (lambda: cases.case_hex3d(n=(12, 10, 8), perturb=0.2, bcs="mixed"), False),
    # 64^3: sub-domains of many TMA tiles per CTA, halo lists of thousands of cells, 100 steps
    "hex64_mixed_procrule": (lambda: cases._with_bcs(cases.pm.hex_box(64, 64, 64), "mixed", cases.GAS, 8e-4), True),
}
if os.environ.get("QGD_MULTI_CASES"):
    CASES = {k: v for k, v in CASES.items() if k in os.environ["QGD_MULTI_CASES"].split(",")}
ok = True
for name, (mk, proc_rule) in CASES.items():
    c = mk()
    nsteps = 100 if name.startswith("hex64") else 50
    if name.endswith("fork2"):
        os.environ["QGD_BND_FORK"] = "2"
    else:
        os.environ.pop("QGD_BND_FORK", None)
    cell_rank = decompose.geometric_split(c.mesh, world)
    sub = decompose.extended_submeshes(c.mesh, cell_rank, ranks=[rank])[0]
    if not proc_rule:
        sub.coupled_face[:] = 0
    # same construction as multigpu.make_rank_solver, with the (possibly cleared) coupled flags
    dm = api.Mesh(sub.mesh, n_owned=sub.n_owned, coupled_face=sub.coupled_face)
    ds = c.diff_solver
    s = api.QGDFoam(dm, fvsc_scheme=c.scheme, qgd_coeffs=c.model, delta_t=c.dt, varsc_cSc1=c.varsc["cSc1"], varsc_minSc=c.varsc["minSc"],
                    varsc_maxSc=c.varsc["maxSc"], implicit_diffusion=c.implicit, diff_tol=ds["tol"], diff_rel_tol=ds["rel_tol"],
                    diff_max_iter=ds["max_iter"], diff_precond=ds["precond"], **c.gas, **c.opts)
    nI_g = c.mesh.n_internal
    bf_g = sub.face_global[sub.mesh.n_internal:]
    phys = bf_g >= nI_g
    idx = np.where(phys, bf_g - nI_g, 0)
    pad = lambda k: np.concatenate([np.asarray(k, np.int32), [1]]).astype(np.int32)
    s.set_bcs(pad(c.bcU), pad(c.bcT), pad(c.bcP), np.where(phys[:, None], c.bvU[idx], 0.0), np.where(phys, c.bvT[idx], 1.0),
              np.where(phys, c.bvP[idx], 1.0))
    cg = sub.cell_global
    s.init_fields(c.U0[cg], c.T0[cg], c.p0[cg], None)
    s.set_halo(sub)
    s.step(nsteps)
    res = {f: s.get(f)[:sub.n_owned] for f in ("rho", "rhoU", "rhoE", "e", "p")}
    np.savez(f"/tmp/qgd_multi_{name}_{rank}.npz", cells=cg[:sub.n_owned], dt=s.scalars()["deltaT"], **res)
    api.synchronize()
    dist.barrier()
    if rank == 0:
        import oracle as O
        o = c.make_oracle(O, n_threads=os.cpu_count() or 1)
        c.oracle_step(o, nsteps)
        for f in ("rho", "rhoU", "rhoE", "e", "p"):
            ref = o.get(f)
            got = np.zeros_like(ref)
            for r in range(world):
                z = np.load(f"/tmp/qgd_multi_{name}_{r}.npz")
                got[z["cells"]] = z[f]
            err = float(np.abs(got - ref).max() / np.abs(ref).max())
            status = "ok" if err < 1e-10 else "FAIL"
            if status == "FAIL":
                ok = False
            say(f"MULTI n={world} {name} steps={nsteps} face_kernel={s.face_kernel()[0]} {f} relLinf={err:.3e} {status}")
        if c.opts["adjust_time_step"]:
            z = np.load(f"/tmp/qgd_multi_{name}_0.npz")
            derr = abs(float(z["dt"]) - o.deltaT()) / o.deltaT()
            say(f"MULTI n={world} {name} deltaT rel err={derr:.3e} {'ok' if derr < 1e-10 else 'FAIL'}")
            ok = ok and derr < 1e-10
    dist.barrier()
if rank == 0:
    say("MULTI_ALL_OK" if ok else "MULTI_FAILED")
api.comm_finalize()
dist.destroy_process_group()
