"""Multi-GPU driver glue (harness side): one process per GPU, torch.distributed only for the NCCL-id broadcast,
barriers and max-over-ranks timing; the data path (halo exchange, all-reduce) is inside libqgd_b200.so."""
from __future__ import annotations

import json
import os
import time

import numpy as np

from . import api, decompose


def init_comm(rank: int, world: int):
    import torch
    import torch.distributed as dist
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8).clone()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    uid = uid.to(dev)
    dist.broadcast(uid, 0)
    api.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))


def make_rank_solver(case, rank: int, world: int, cell_rank=None):
    """Extended sub-mesh of `rank` + solver with halo lists.  `case` is a tests/cases.Case on the GLOBAL mesh.
    `cell_rank`: cell -> processor map, e.g. foamcase.read_cell_decomposition(case_dir, n_cells) of a decomposePar'd case
    (the device decomposition is then the scotch decomposition, bit for bit); default: geometric split."""
    mesh = case.mesh
    if cell_rank is None:
        cell_rank = decompose.geometric_split(mesh, world)
    sub = decompose.extended_submeshes(mesh, cell_rank, ranks=[rank])[0]
    dm = api.Mesh(sub.mesh, n_owned=sub.n_owned, coupled_face=sub.coupled_face)
    vs = getattr(case, "varsc", None) or dict(cSc1=1.0, minSc=-1.0, maxSc=-1.0, const_sc_cells=None)
    s = api.QGDFoam(dm, fvsc_scheme=case.scheme, qgd_coeffs=getattr(case, 'model', 'constScPrModel1'), delta_t=case.dt,
                    varsc_cSc1=vs["cSc1"], varsc_minSc=vs["minSc"], varsc_maxSc=vs["maxSc"], **case.gas, **case.opts)
    if vs["const_sc_cells"] is not None:            # global cell ids -> local ids of the extended sub-mesh
        g2l = np.full(mesh.n_cells, -1, np.int64)
        g2l[sub.cell_global] = np.arange(sub.cell_global.size)
        loc = g2l[np.asarray(vs["const_sc_cells"], np.int64)]
        s.set_const_sc_cells(loc[loc >= 0].astype(np.int32))
    # per-patch BC kinds: global patches + the cut patch (kind irrelevant); per-face values follow the local faces
    nI_g = mesh.n_internal
    bf_g = sub.face_global[sub.mesh.n_internal:]
    is_phys = bf_g >= nI_g
    idx = np.where(is_phys, bf_g - nI_g, 0)
    pad = lambda k: np.concatenate([np.asarray(k, np.int32), [1]]).astype(np.int32)
    valU = np.where(is_phys[:, None], case.bvU[idx], 0.0)
    valT = np.where(is_phys, case.bvT[idx], 1.0)
    valP = np.where(is_phys, case.bvP[idx], 1.0)
    s.set_bcs(pad(case.bcU), pad(case.bcT), pad(case.bcP), valU, valT, valP)
    cg = sub.cell_global
    s.init_fields(case.U0[cg], case.T0[cg], case.p0[cg], None if case.alphaQGD is None else case.alphaQGD[cg])
    s.set_halo(sub)
    return s, sub, dm


def bench(args, rank: int, world: int, local: int):
    """bench.py body for --gpus > 1 (strong scaling of the 256^3 box; rank 0 prints the JSON line)."""
    import torch
    import torch.distributed as dist
    import bench as B
    init_comm(rank, world)
    case = B.build_case(args.size)
    mesh = case.mesh
    s, sub, dm = make_rank_solver(case, rank, world)
    s.step(args.warmup)
    api.synchronize()
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = s.launch_count()
    dist.barrier(); torch.cuda.synchronize(); api.synchronize()
    s.profile(True)
    api.timer_begin()
    s.step(args.steps)
    ms = api.timer_end()
    api.synchronize(); torch.cuda.synchronize()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    ms = float(t.item())
    kt = s.kernel_times()
    s.profile(False)
    launches = s.launch_count() - l0
    # ---- end to end through the C-ABI with pinned HOST buffers: every rank uploads U, T, p of its sub-domain (halo included),
    # the state is re-created from them (createFields.H semantics), one step with the NCCL halo exchange runs, U, T, p come back
    nL = sub.mesh.n_cells
    pinned = {k: torch.empty((nL, 3) if k == "U" else (nL,), dtype=torch.float64, pin_memory=True) for k in ("U", "T", "p")}
    fl = {k: v.numpy() for k, v in pinned.items()}
    s.step_fields_host(0, None, fl)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(3):
        s.step_fields_host(1, fl, fl)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.step_fields_host(1, fl, fl)
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    tb = torch.tensor([5.0 * 8 * nL], dtype=torch.float64, device="cuda")
    dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    e2e_s, e2e_bytes = float(te.item()), int(tb.item())
    kname, l2hint = s.face_kernel()
    if rank == 0:
        clocks = sampler.stop()
        ms_step = ms / args.steps
        value = mesh.n_cells / (ms_step * 1e-3) / 1e6
        ab = B.alg_bytes(mesh)
        peak, src = B.peaks()
        face_ms = kt["face_ms"] / max(kt["steps"], 1)
        ab_face_rank = 184 * sub.mesh.n_internal + 40 * sub.mesh.n_cells + 48 * sub.mesh.n_points
        halo_bytes = 8 * (16 * sum(a.size for a in sub.send_cells.values()) + 20 * sum(a.size for a in sub.send_bfaces.values()))
        line = {"metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"QGDFoam 3D synthetic hex box {args.size}^3 ({mesh.n_cells} cells), explicit, FP64, "
                                       f"geometric {'x'.join(map(str, decompose.split_factors(world)))} decomposition",
                           "fvsc": "GaussVolPoint", "QGDCoeffs": "constScPrModel1", "implicitDiffusion": False,
                           "deltaT": case.dt, "owned_cells_rank0": sub.n_owned, "halo_cells_rank0": sub.mesh.n_cells - sub.n_owned,
                           "halo_bytes_sent_per_step_rank0": halo_bytes,
                           "l2": "per-rank state and mesh records exceed the 126 MB L2; no flush"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": mesh.n_cells / e2e_s / 1e6, "unit": B.UNIT, "h2d_bytes_per_step": e2e_bytes,
                        "d2h_bytes_per_step": e2e_bytes, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                        "api": "qgd_qgdfoam_step_fields_host on every rank: U, T, p of the sub-domain (5 doubles/cell, halo included) "
                               "H2D, state re-created as createFields.H does, 1 step with NCCL halo exchange, U, T, p D2H per call, "
                               "pinned host buffers; max over ranks"},
                "roofline": {"bound": "hbm", "kernel": f"{kname} (rank 0)", "l2hint": l2hint, "achieved": ab_face_rank / (face_ms * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": ab_face_rank / (face_ms * 1e-3) / 1e9 / peak, "traffic": None,
                             "peak_source": src, "avg_launch_ms": face_ms,
                             "step": {"alg_bytes_global": ab["total"], "achieved_aggregate": ab["total"] / (ms_step * 1e-3) / 1e9,
                                      "frac_of_n_gpus_peak": ab["total"] / (ms_step * 1e-3) / 1e9 / (peak * world)}},
                "cpu_baseline": None}
        fd = getattr(args, "json_fd", None)
        if fd is not None:
            os.write(fd, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    api.comm_finalize()
    dist.destroy_process_group()
