"""One rank of a decomposed QGDFoam run with all ranks ON ONE GPU: `python tests/loopback_worker.py <rank> <world> <rendezvous dir>
<case,case,...>`.  Everything is the real library path of an N-GPU run (extended sub-mesh, exchange lists, pack / unpack kernels,
the global decisions that steer the collectives, device-side time-step control); only the transport underneath ncclSend / ncclRecv /
ncclAllReduce is the loopback stand-in tests/fake_nccl (first on LD_LIBRARY_PATH; torch is never imported, so the real libnccl is
not in the process).  Rank 0 gathers the owned parts and compares them with the serial CPU oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from qgdsolver_b200 import api, decompose  # noqa: E402

import signal  # noqa: E402

signal.alarm(330)                               # a rank left alone in an exchange ends by itself
rank, world, rdv = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
names = sys.argv[4].split(",")


def barrier(tag):
    open(os.path.join(rdv, f"bar_{tag}_{rank}"), "w").close()
    t0 = time.time()
    while not all(os.path.exists(os.path.join(rdv, f"bar_{tag}_{r}")) for r in range(world)):
        time.sleep(0.005)
        if time.time() - t0 > 600:
            raise SystemExit(f"rank {rank}: barrier {tag} timed out")


def slip_case():
    c = cases.case_hex3d(n=(12, 10, 8), perturb=0.2, bcs="zg")
    for i in range(1, len(c.mesh.patches), 2):
        c.bcU[i] = cases.SLIP
    return c


# the QGDFoam cases of tests/multi_gpu_worker.py (explicit branch); second entry: the reference's processor-patch rule for hQGDf
CASES = {
    # with 2 x 2 x 2 sub-domains a corner rank holds no face of the odd (qgdFlux) patches: the configuration that hung at N = 8
    "perturbed_mixed_serialrule": (lambda: cases.case_hex3d(n=(12, 10, 8), perturb=0.2, bcs="mixed"), False),
    "prism_fixed_serialrule": (lambda: cases.case_prism(n=(6, 5, 4), bcs="fixed"), False),
    "2d_qgdflux_serialrule": (lambda: cases.case_2d((20, 16), perturb=0.15, bcs="qgdflux"), False),
    "uniform_zg_procrule": (lambda: cases.case_hex3d(n=(12, 10, 8), bcs="mixed"), True),
    "uniform_adjust_procrule": (lambda: cases.case_hex3d(n=(12, 10, 8), bcs="fixed", adjust_time_step=True, dt=1e-3, max_co=0.1, c_tau=0.3), True),
    "truncoct_mixed_serialrule": (lambda: cases.case_truncoct(n=(6, 5, 5), bcs="mixed"), False),
    "slip_perturbed_serialrule": (slip_case, False),
    "varSc7_fixed_serialrule": (lambda: cases.case_hex3d(n=(12, 10, 8), perturb=0.1, bcs="fixed", model="varScModel7",
                                                         varsc=dict(cSc1=3.0, minSc=0.02, maxSc=0.4)), False),
    "2d_leastSquares_serialrule": (lambda: cases.case_2d((20, 16), perturb=0.2, bcs="mixed", scheme="leastSquares"), False),
    # implicitDiffusion true on sub-meshes: stepwise PCG, search direction exchanged every iteration, all-reduced dot products
    "perturbed_mixed_implicit_serialrule": (lambda: cases.case_hex3d(n=(12, 10, 8), perturb=0.2, bcs="mixed", implicit=True,
                                                                     diff_solver=dict(precond="diagonal")), False),
    "2d_qgdflux_implicit_adjust_serialrule": (lambda: cases.case_2d((20, 16), perturb=0.15, bcs="qgdflux", implicit=True, adjust_time_step=True,
                                                                    max_co=0.1, diff_solver=dict(precond="none")), False),
}

api.load_library()
api.init(0)                                     # every rank on GPU 0
idf = os.path.join(rdv, "id.bin")
if rank == 0:
    os.environ["FAKE_NCCL_DIR"] = rdv
    uid = api.comm_unique_id()
    with open(idf + ".tmp", "wb") as f:
        f.write(bytes(uid))
    os.rename(idf + ".tmp", idf)
else:
    while not os.path.exists(idf):
        time.sleep(0.01)
    uid = open(idf, "rb").read()
api.comm_init(rank, world, bytes(uid))
maps = open("/proc/self/maps").read()
assert "fake_nccl/libnccl.so.2" in maps, "the loopback transport is not the libnccl this process bound"
assert "site-packages" not in "".join(ln for ln in maps.splitlines() if "libnccl" in ln), "a real libnccl is mapped as well"

ok = True
for name in names:
    mk, proc_rule = CASES[name]
    c = mk()
    nsteps = 50
    cell_rank = decompose.geometric_split(c.mesh, world)
    sub = decompose.extended_submeshes(c.mesh, cell_rank, ranks=[rank])[0]
    if not proc_rule:
        sub.coupled_face[:] = 0
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
    # how many ranks hold no face of a qgdFlux patch (the rank-local decision that caused the N = 8 hang would differ there)
    pid = sub.mesh.patch_id_per_bface()
    has_qf_face = bool(np.any((pad(c.bcP)[pid] == cases.QF) & (sub.mesh.patch_kind_per_bface() != 1)))
    cg = sub.cell_global
    s.init_fields(c.U0[cg], c.T0[cg], c.p0[cg], None)
    s.set_halo(sub)
    s.step(nsteps)
    res = {f: s.get(f)[:sub.n_owned] for f in ("rho", "rhoU", "rhoE", "e", "p")}
    np.savez(os.path.join(rdv, f"res_{name}_{rank}.npz"), cells=cg[:sub.n_owned], dt=s.scalars()["deltaT"], has_qf=has_qf_face, **res)
    api.synchronize()
    barrier(name + "_done")
    if rank == 0:
        import oracle as O
        o = c.make_oracle(O, n_threads=2)
        c.oracle_step(o, nsteps)
        z = [np.load(os.path.join(rdv, f"res_{name}_{r}.npz")) for r in range(world)]
        any_qf = bool(np.any(np.asarray(c.bcP) == cases.QF))
        without = sum(1 for q in z if not bool(q["has_qf"])) if any_qf else 0
        for f in ("rho", "rhoU", "rhoE", "e", "p"):
            ref = o.get(f)
            got = np.zeros_like(ref)
            for q in z:
                got[q["cells"]] = q[f]
            err = float(np.abs(got - ref).max() / np.abs(ref).max())
            good = err < 1e-10
            ok = ok and good
            print(f"LOOPBACK n={world} {name} steps={nsteps} ranks_without_a_qgdFlux_face={without} {f} relLinf={err:.3e} {'ok' if good else 'FAIL'}", flush=True)
        if c.opts["adjust_time_step"]:
            derr = abs(float(z[0]["dt"]) - o.deltaT()) / o.deltaT()
            ok = ok and derr < 1e-10
            print(f"LOOPBACK n={world} {name} deltaT rel err={derr:.3e} {'ok' if derr < 1e-10 else 'FAIL'}", flush=True)
    barrier(name + "_checked")
if rank == 0:
    print("LOOPBACK_ALL_OK" if ok else "LOOPBACK_FAILED", flush=True)
api.comm_finalize()
