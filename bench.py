#!/usr/bin/env python
"""bench.py — MCUPS of the QGDFoam step (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps K --warmup W            product arm (CUDA, through the C-ABI)
  python bench.py --impl reference --steps K --warmup W    reference arm: the CPU oracle port on the host cores
                                                           (the reference itself needs OpenFOAM v2312 and cannot
                                                           be built here - DESIGN.md)
A "step" is one pass of the QGDFoam.C:90-163 loop body over the whole mesh.  Workload at N=1: BASELINE.json configs[3],
QGDFoam 3D synthetic hex box 256^3 (16 777 216 cells, FP64, GaussVolPoint, constScPrModel1, explicit), which is the
configuration the metric's roofline target is quoted on; state (>2 GB) is far larger than L2, so no flush is needed.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "cell-updates/s per QGDFoam step"
UNIT = "MCUPS"
GAS = dict(R=1.0, Cp=3.5, Hf=0.0, Tref=0.0, Hsref=0.0, mu=1.8e-5, Pr=0.71, ScQGD=1.0, PrQGD=1.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_case(n: int, dt: float = None):
    """SURVEY 8(d) synthetic C4 inputs: uniform hex box n^3, six zeroGradient patches, Taylor-Green-like IC."""
    import cases
    mesh = _cached_hex_box(cases.pm, n)
    dt = dt if dt is not None else 2.0e-4 * (256.0 / n)     # Co ~ 0.06 at any n
    U0, T0, p0 = cases.smooth_ic(mesh, GAS)
    bc = cases.uniform_bcs(mesh)
    return cases.Case(mesh, U0, T0, p0, *bc, gas=GAS, dt=dt)


def _cached_hex_box(pm, n):
    """hex_box(n,n,n) with an on-box cache under /tmp (mesh synthesis is input generation, not the measured path)."""
    path = f"/tmp/qgd_hexbox_{n}.npz"
    fields = ("points", "face_offsets", "face_verts", "owner", "neighbour", "C", "V", "Cf", "Sf", "magSf", "weights",
              "deltaCoeffs", "nonOrthDeltaCoeffs", "neighb_cell_centres", "geometric_d")
    if n >= 128 and os.path.exists(path):
        try:
            z = np.load(path)
            m = pm.hex_box(2, 2, 2)
            patches = [pm.Patch(str(a), int(b), int(c), int(d)) for a, b, c, d in
                       zip(z["patch_name"], z["patch_kind"], z["patch_start"], z["patch_size"])]
            mesh = pm.PolyMesh(points=z["points"], face_offsets=z["face_offsets"], face_verts=z["face_verts"],
                               owner=z["owner"], neighbour=z["neighbour"], patches=patches, n_cells=n * n * n)
            for f in fields[5:]:
                setattr(mesh, f, z[f])
            return mesh
        except Exception:
            pass
    mesh = pm.hex_box(n, n, n)
    if n >= 128 and int(os.environ.get("RANK", "0")) == 0:
        try:
            np.savez(path, **{f: getattr(mesh, f) for f in fields},
                     patch_name=np.array([p.name for p in mesh.patches]), patch_kind=np.array([p.kind for p in mesh.patches]),
                     patch_start=np.array([p.start for p in mesh.patches]), patch_size=np.array([p.size for p in mesh.patches]))
        except Exception:
            pass
    return mesh


def alg_bytes(mesh):
    """Algorithmic bytes per step: total and per kernel."""
    nC, nF, nP = mesh.n_cells, mesh.n_internal, mesh.n_points
    face = 184 * nF + 40 * nC + 48 * nP
    points = 148 * nP
    cell = 40 * nF + 64 * nC
    return dict(total=face + points + cell, face=face, points=points, cell=cell)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_mem_available_gb() -> float:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return float(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


ORACLE_BYTES_PER_CELL = 2500      # measured: oracle context of an n^3 hex box (field-at-a-time arrays + 18-coefficient face records)
MESH_BYTES_PER_CELL = 1900        # python-side PolyMesh arrays of the same box


def cpu_sample_size(n: int, have_mesh: bool) -> int:
    """Edge of the hex box the CPU arm runs: the product arm's own n^3 when the host has the memory for the oracle's
    field-at-a-time working set (2.5 kB per cell, ~42 GB at 256^3), else the largest power-of-two fraction that fits."""
    avail = host_mem_available_gb() * 1e9
    m = n
    while m > 32 and (ORACLE_BYTES_PER_CELL + (0 if (have_mesh and m == n) else MESH_BYTES_PER_CELL)) * m ** 3 * 1.5 > avail:
        m //= 2
    return m


def time_oracle(n: int, budget_s: float, threads: int, min_steps: int = 2, case=None):
    """CPU oracle on an n^3 hex box of the same workload; returns MCUPS and the sample description."""
    import oracle as O
    O.build()
    c = case if case is not None else build_case(n)
    o = c.make_oracle(O, n_threads=threads)
    c.oracle_step(o, 1)                                      # warm-up
    steps, t0 = 0, time.perf_counter()
    while True:
        c.oracle_step(o, 1)
        steps += 1
        el = time.perf_counter() - t0
        if steps >= min_steps and el > budget_s:
            break
    return c.mesh.n_cells * steps / el / 1e6, f"{n}^3 hex box ({c.mesh.n_cells} cells), {steps} steps in {el:.1f} s", steps, el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # the product arm's own workload (args.size^3 = 256^3) on all host cores when the host memory allows it
    n = args.ref_size if args.ref_size > 0 else cpu_sample_size(args.size, False)
    import oracle as O
    O.build()
    c = build_case(n, dt=2.0e-4 * (256.0 / n))
    o = c.make_oracle(O, n_threads=threads)
    c.oracle_step(o, max(args.warmup, 1))
    t0 = time.perf_counter()
    c.oracle_step(o, args.steps)
    el = time.perf_counter() - t0
    val = c.mesh.n_cells * args.steps / el / 1e6
    same = n == args.size
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the same config block as the product arm prints for this workload (the arm itself is described beside it)
            "config": {"workload": f"QGDFoam 3D synthetic hex box {n}^3 ({c.mesh.n_cells} cells), explicit, FP64",
                       "fvsc": "GaussVolPoint", "QGDCoeffs": "constScPrModel1", "implicitDiffusion": False, "deltaT": c.dt,
                       "l2": "inputs (>2 GB of state and mesh records) exceed the 126 MB L2; no flush"},
            "arm": "CPU oracle port of the reference algorithm (the reference needs OpenFOAM v2312: unbuildable here), "
                   f"OpenMP over the whole mesh on {threads} host threads in place of mpirun -np {threads}",
            "same_workload_as_product_arm": same,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n}^3 hex box ({c.mesh.n_cells} cells) x {args.steps} steps, OpenMP {threads} threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_product(args):
    import torch
    from qgdsolver_b200 import api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    api.init(local)
    if world > 1:
        # keep stdout to the single JSON line: library banners (e.g. "NCCL version ...") go to stderr
        sys.stdout.flush()
        args.json_fd = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from qgdsolver_b200 import multigpu
        return multigpu.bench(args, rank, world, local)

    n = args.size
    c = build_case(n)
    mesh = c.mesh
    s = c.make_solver(api)
    ab = alg_bytes(mesh)
    # ---- device-resident loop
    s.step(args.warmup)
    api.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = s.launch_count()
    s.profile(True)
    torch.cuda.synchronize()
    api.timer_begin()
    s.step(args.steps)
    ms = api.timer_end()
    torch.cuda.synchronize()
    kt = s.kernel_times()
    s.profile(False)
    launches = s.launch_count() - l0
    clocks = sampler.stop()
    ms_step = ms / args.steps
    value = mesh.n_cells / (ms_step * 1e-3) / 1e6
    peak, peak_src = peaks()
    face_ms = kt["face_ms"] / max(kt["steps"], 1)
    pipe = s.get_pipeline()
    kname, l2hint = s.face_kernel()              # what the library launched, not a re-derivation
    kbytes = ab["face"] + ab["cell"] if pipe["mode"] == 1 else ab["face"]
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "face_flux_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
            if tj.get("n_cells") == mesh.n_cells and kname + "<" in tj.get("kernel", ""):
                if tj.get("l2hint", 0) == l2hint:
                    traffic = tj.get("dram_bytes_per_launch")
                    traffic_note = tj.get("source")
                else:       # an ncu capture of another variant of the kernel is not this kernel's traffic
                    traffic_note = (f"no ncu capture of the l2hint={l2hint} variant; l2hint={tj.get('l2hint', 0)} "
                                    f"moved {tj.get('dram_bytes_per_launch', 0) / 1e9:.2f} GB per launch ({tj.get('source', '')})")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": kbytes / (face_ms * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": kbytes / (face_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_note": traffic_note, "l2hint": l2hint,
                "peak_source": peak_src,
                "alg_bytes_per_launch": kbytes, "avg_launch_ms": face_ms, "pipeline": pipe,
                "step": {"alg_bytes": ab["total"], "achieved": ab["total"] / (ms_step * 1e-3) / 1e9,
                         "frac": ab["total"] / (ms_step * 1e-3) / 1e9 / peak},
                "kernel_ms": {"k_points": kt["points_ms"] / max(kt["steps"], 1), kname: face_ms,
                              "k_cell_update": kt["cell_ms"] / max(kt["steps"], 1)}}
    # ---- end-to-end through the C-ABI with host (pinned) buffers, every step: the fields of a time directory (U, T, p) go
    # up, the state is re-created from them as createFields.H does, one step runs, U, T, p come back (restart semantics)
    nC = mesh.n_cells
    e2e = e2e_fields(s, nC, max(3, min(args.steps, 10)))
    if args.e2e_full_state:
        e2e["full_state"] = e2e_full_state(s, nC, 5)
    # ---- CPU baseline (oracle port): the same n^3 box when the host memory allows it, bounded number of steps
    cpu = None
    if not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        m = args.ref_size if args.ref_size > 0 else cpu_sample_size(n, True)
        del s
        v, sample, _, _ = time_oracle(m, args.cpu_budget, threads, case=c if m == n else None)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"QGDFoam 3D synthetic hex box {n}^3 ({nC} cells), explicit, FP64",
                       "fvsc": "GaussVolPoint", "QGDCoeffs": "constScPrModel1", "implicitDiffusion": False,
                       "deltaT": c.dt, "l2": "inputs (>2 GB of state and mesh records) exceed the 126 MB L2; no flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def e2e_fields(s, nC, steps):
    import torch
    pinned = {k: torch.empty((nC, 3) if k == "U" else (nC,), dtype=torch.float64, pin_memory=True) for k in ("U", "T", "p")}
    fl = {k: v.numpy() for k, v in pinned.items()}
    s.step_fields_host(0, None, fl)
    for _ in range(3):
        s.step_fields_host(1, fl, fl)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step_fields_host(1, fl, fl)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    nbytes = 5 * 8 * nC
    return {"value": nC / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
            "ms_per_step": e2e_s * 1e3, "steps": steps,
            "api": "qgd_qgdfoam_step_fields_host: U, T, p (5 doubles/cell) H2D, state re-created as createFields.H does, 1 step, "
                   "U, T, p D2H - every call, pinned host buffers"}


def e2e_full_state(s, nC, steps):
    import torch
    names = ("rho", "U", "e", "p", "T", "rhoU", "rhoE", "mu")
    pinned = {k: torch.empty((nC, 3) if k in ("U", "rhoU") else (nC,), dtype=torch.float64, pin_memory=True) for k in names}
    st = {k: v.numpy() for k, v in pinned.items()}
    s.step_host(0, None, st)
    for _ in range(2):
        s.step_host(1, st, st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step_host(1, st, st)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    return {"value": nC / e2e_s / 1e6, "unit": UNIT, "bytes_each_way_per_step": 12 * 8 * nC, "ms_per_step": e2e_s * 1e3,
            "api": "qgd_qgdfoam_step_host: full cell state (12 doubles/cell) H2D + 1 step + D2H per call"}


def run_qgd2d(args):
    """Extra line (not the driver's default): BASELINE configs[1]-like, QGDFoam on a 2D n x n hex mesh (one cell thick, empty
    front/back), explicit, FP64.  Algorithmic bytes per step: 104 nC + 184 nF + 148 nP_used."""
    import torch
    import cases
    from qgdsolver_b200 import api
    torch.cuda.set_device(0)
    api.init(0)
    n = args.qhd_size
    if args.geometry == "step":
        # BASELINE configs[1]: Mach-3 forward-facing step (Woodward-Colella), slip walls, ~1.03 M hex cells at n = 640
        n = 640 if n == 1000 else n
        c = cases.case_forward_step(n=n)
        mesh = c.mesh
        what = f"Mach-3 forward-facing step, {3 * n} x {n} cells minus the step"
    else:
        mesh = cases.pm.hex_box(n, n, 1, lengths=(1.0, 1.0, 1.0 / n), patch_kinds={"zMin": "empty", "zMax": "empty"})
        c = cases._with_bcs(mesh, "zg", GAS, 2.0e-4 * (256.0 / n))
        what = f"hex mesh {n}x{n}"
    s = c.make_solver(api)
    s.step(args.warmup)
    api.synchronize()
    reps = []
    for _ in range(3):                       # state (~0.3 GB) is larger than L2; no flush
        api.timer_begin()
        s.step(args.steps)
        reps.append(api.timer_end() / args.steps)
    ms = min(reps)
    nC, nI = mesh.n_cells, mesh.n_internal
    alg = 104 * nC + 184 * nI + 148 * (mesh.n_points // 2)
    peak, src = peaks()
    line = {"metric": METRIC, "value": nC / ms / 1e3, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"QGDFoam 2D {what} ({nC} cells, empty front/back), explicit, FP64",
                       "l2": "state and mesh records (~0.4 GB) exceed the 126 MB L2; no flush", "face_kernel": s.face_kernel()[0],
                       "fvsc": "GaussVolPoint", "QGDCoeffs": "constScPrModel1", "implicitDiffusion": False},
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "alg_bytes_per_step": alg, "peak_source": src},
            "gpu_launches": int(s.launch_count())}
    print(json.dumps(line), flush=True)


def run_poly(args):
    """Extra line (not the driver's default): BASELINE configs[4]-like single-GPU share, QGDFoam on the truncated-octahedron
    polyhedral mesh (14 faces per cell: hexagons take the `other faces` branch, squares the six-point GaussVolPoint formula).
    Algorithmic bytes per step: 104 nC + 224 nF_quad + 136 nF_other + 196 nP."""
    import torch
    import cases
    from qgdsolver_b200 import api
    torch.cuda.set_device(0)
    api.init(0)
    n = args.poly_n
    c = cases.case_truncoct(n=(n, n, n))
    mesh = c.mesh
    s = c.make_solver(api)
    s.step(args.warmup)
    api.synchronize()
    reps = []
    for _ in range(3):
        api.timer_begin()
        s.step(args.steps)
        reps.append(api.timer_end() / args.steps)
    ms = min(reps)
    nC, nI = mesh.n_cells, mesh.n_internal
    nv = mesh.face_nverts()[:nI]
    nq, no = int((nv <= 4).sum()), int((nv > 4).sum())
    alg = 104 * nC + 224 * nq + 136 * no + 196 * mesh.n_points
    peak, src = peaks()
    line = {"metric": METRIC, "value": nC / ms / 1e3, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"QGDFoam 3D polyhedral mesh, truncated octahedra on a {n}^3 BCC lattice ({nC} cells, "
                                   f"{nI / nC:.2f} internal faces per cell, {no} polygon + {nq} quad faces), explicit, FP64",
                       "fvsc": "GaussVolPoint", "QGDCoeffs": "constScPrModel1", "implicitDiffusion": False,
                       "l2": "state and mesh records exceed the 126 MB L2 for n >= 60; no flush"},
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "alg_bytes_per_step": alg, "peak_source": src},
            "gpu_launches": int(s.launch_count())}
    print(json.dumps(line), flush=True)


def run_qhd(args):
    """Extra line (not the driver's default): BASELINE configs[2], QHDFoam 2D differentially heated cavity, n x n cells,
    pressure PCG on the device.  A step = one QHDFoam.C:83-139 pass including the whole PCG solve."""
    import torch
    import cases
    from qgdsolver_b200 import api
    torch.cuda.set_device(0)
    api.init(0)
    n = args.qhd_size
    blocks = args.pcg_blocks if args.precond == "DIC" else 0

    def make(max_iter):
        c = cases.qhd_cavity(n=(n, n), dt=args.qhd_dt, precond=args.precond, tol=args.p_tol, rel_tol=args.p_rel_tol, max_iter=max_iter,
                             model="constTau", coeffs=dict(Tau=args.qhd_dt))
        dm = api.Mesh(c.mesh)
        nb = 0
        if blocks:
            nb = int(dm.make_pcg_blocks(blocks).max()) + 1
        return c, c.make_solver(api, dm), nb

    def timed(s, steps):
        api.timer_begin()
        for _ in range(steps):
            s.step(1)
        return api.timer_end() / steps

    c, s, n_blocks = make(100000)
    s.step(args.warmup)
    api.synchronize()
    ms = timed(s, args.steps)
    info = s.solver_info()
    # the step without PCG iterations (maxIter 0: only the initial residual) separates the solver from the rest of the step
    _, s0, _ = make(0)
    s0.step(2)
    api.synchronize()
    ms0 = timed(s0, 3)
    nC, nI = c.mesh.n_cells, c.mesh.n_internal
    peak, src = peaks()
    it = max(info["iters"], 1)
    us_iter = (ms - ms0) * 1e3 / it
    pcg_bytes = 120 * nC + 48 * nI                  # SpMV + DIC sweeps + fused axpy / dots per iteration
    line = {"metric": "cell-updates/s per QHDFoam step", "value": nC / ms / 1e3, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"QHDFoam 2D buoyant differentially-heated cavity {n}x{n} ({nC} cells), explicit, FP64",
                       "fvsc": "GaussVolPoint", "QGDCoeffs": "constTau", "p_solver": f"PCG + {args.precond}",
                       "dic_blocks": n_blocks, "dic_block_target_cells": blocks,
                       "tolerance": args.p_tol, "relTol": args.p_rel_tol, "last_pcg_iterations": info["iters"],
                       "last_final_residual": info["final_residual"]},
            "pcg": {"iterations": info["iters"], "us_per_iteration": us_iter, "ms_step_without_iterations": ms0,
                    "alg_bytes_per_iteration": pcg_bytes},
            "roofline": {"bound": "hbm", "kernel": "k_pcg (one iteration)", "achieved": pcg_bytes / (us_iter * 1e-6) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": pcg_bytes / (us_iter * 1e-6) / 1e9 / peak, "traffic": None, "peak_source": src},
            "gpu_launches": int(s.launch_count())}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--size", type=int, default=256, help="hex box edge (cells); 256 = BASELINE configs[3]")
    ap.add_argument("--ref-size", type=int, default=0, help="edge of the CPU arm's hex box; 0 = --size when the host memory allows it")
    ap.add_argument("--e2e-full-state", action="store_true", help="also time the 12-doubles-per-cell qgd_qgdfoam_step_host hand-off")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--poly-n", type=int, default=100, help="--case poly: BCC lattice edge (cells ~ 2 n^3)")
    ap.add_argument("--case", default="qgd3d", choices=["qgd3d", "qgd2d", "qhd2d", "poly"],
                    help="qgd3d = BASELINE configs[3] (default); qgd2d = a configs[1]-sized 2D mesh; qhd2d = configs[2]; "
                         "poly = a configs[4]-shaped polyhedral mesh (one GPU's share)")
    ap.add_argument("--geometry", default="step", choices=["step", "box"], help="--case qgd2d: forward-facing step (configs[1]) or a plain box")
    ap.add_argument("--precond", default="diagonal")
    ap.add_argument("--pcg-blocks", type=int, default=128, help="--precond DIC: target cells per DIC block (0 = the serial, level-scheduled DIC)")
    ap.add_argument("--p-tol", type=float, default=1e-8)
    ap.add_argument("--p-rel-tol", type=float, default=0.0)
    ap.add_argument("--qhd-dt", type=float, default=1e-5, help="explicit QHD step: dt < h^2/(4 nu) = 2.5e-5 at 1000^2")
    ap.add_argument("--qhd-size", type=int, default=1000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.case == "qhd2d":
        run_qhd(args)
    elif args.case == "qgd2d":
        run_qgd2d(args)
    elif args.case == "poly":
        run_poly(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
