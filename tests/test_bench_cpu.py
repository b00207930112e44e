"""bench.py contract checks that need no GPU: the reference arm (the CPU oracle port timed on the host cores) prints one
JSON line with the keys the driver reads, and the algorithmic-byte model matches SURVEY 8(d) / DESIGN.md section 3."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-size", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "MCUPS" and j["higher_is_better"] is True and j["value"] > 0
    assert j["metric"] == "cell-updates/s per QGDFoam step" and j["dtype"] == "f64" and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    assert "workload" in j["config"] and "model" not in j["config"] and j["same_workload_as_product_arm"] is False and "OpenMP" in j["arm"]


def test_algorithmic_bytes_follow_the_survey_model():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import cases
    m = cases.pm.hex_box(6, 5, 4)
    ab = bench.alg_bytes(m)
    nC, nF, nP = m.n_cells, m.n_internal, m.n_points
    assert ab["total"] == 104 * nC + 224 * nF + 196 * nP          # b_c, b_f, b_p of the 3D tri/quad path
    assert ab["total"] == ab["face"] + ab["points"] + ab["cell"]
    # the figure quoted in DESIGN.md section 3 for the 256^3 box
    nC, nF, nP = 256 ** 3, 3 * 256 * 256 * 255, 257 ** 3
    assert abs((104 * nC + 224 * nF + 196 * nP) / 1e9 - 16.30) < 0.01


def test_product_arm_assembles_its_json_line_with_a_stand_in_device(monkeypatch, capsys):
    """Dry run of bench.run_product with the device API replaced by recorders (no GPU here): every key of the contract is
    present and consistent - this guards the line-assembly code that only ever runs on the GPU box."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import argparse
    import numpy as np
    import torch
    import bench
    from qgdsolver_b200 import api

    class FakeMesh:
        def __init__(self, mesh, **kw):
            self.mesh = mesh

    class FakeSolver:
        launches = 0

        def __init__(self, dmesh, **kw):
            self.n = 0
        def set_bcs(self, *a): pass
        def init_fields(self, *a): pass
        def step(self, n): self.n += n; FakeSolver.launches += 7 * n
        def launch_count(self): return FakeSolver.launches
        def profile(self, on): pass
        def kernel_times(self): return dict(points_ms=0.65 * 4, face_ms=2.2 * 4, cell_ms=1.2 * 4, steps=4)
        def get_pipeline(self): return dict(mode=0, chunk_cells=512, lag=0, ring_slots=0, n_chunks=0, grid=0)
        def step_host(self, n, a, b): self.n += n
        def step_fields_host(self, n, a, b): self.n += n
        def face_kernel(self): return ("k_face_flux_tma", 3)
    monkeypatch.setattr(api, "Mesh", FakeMesh)
    monkeypatch.setattr(api, "QGDFoam", FakeSolver)
    monkeypatch.setattr(api, "init", lambda d: None)
    monkeypatch.setattr(api, "synchronize", lambda: None)
    monkeypatch.setattr(api, "timer_begin", lambda: None)
    monkeypatch.setattr(api, "timer_end", lambda: 4.2 * 4)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, pin_memory=False, **k: real_empty(*a, **k))
    args = argparse.Namespace(gpus=1, steps=4, warmup=3, size=8, ref_size=0, cpu_budget=0.05, no_cpu_baseline=False, e2e_full_state=True)
    bench.run_product(args)
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(out) == 1
    j = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in j, key
    assert j["steps"] == 4 and j["warmup"] == 3 and j["ms_per_step"] == 4.2 and j["gpu_launches"] == 28 and j["dtype"] == "f64"
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["kernel"] == "k_face_flux_tma" and r["l2hint"] == 3 and r["traffic"] is None      # 8^3 is not the captured 256^3 kernel
    assert abs(r["step"]["frac"] - r["step"]["achieved"] / r["peak"]) < 1e-12
    assert j["e2e"]["h2d_bytes_per_step"] == 5 * 8 * 512 and j["e2e"]["full_state"]["bytes_each_way_per_step"] == 12 * 8 * 512 and\
        "8^3 hex box" in j["cpu_baseline"]["sample"] and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] > 0
    assert "workload" in j["config"] and j["vs_baseline"] is None
    # the extra polyhedral line (configs[4] shape) through the same stand-in device
    args.poly_n = 5
    bench.run_poly(args)
    jp = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][0])
    nC = 5 ** 3 + 4 ** 3
    assert jp["config"]["workload"].startswith("QGDFoam 3D polyhedral mesh") and f"({nC} cells" in jp["config"]["workload"]
    assert jp["roofline"]["alg_bytes_per_step"] > 104 * nC and jp["value"] > 0 and jp["unit"] == "MCUPS"
    # the traffic figure is only quoted for the kernel variant that was actually captured with ncu
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "profiles"))
        monkeypatch.setattr(bench, "ROOT", tmp)
        for hint, expect in ((0, None), (3, 123.0)):
            with open(os.path.join(tmp, "profiles", "face_flux_traffic.json"), "w") as f:
                json.dump({"n_cells": 512, "kernel": "k_face_flux_tma<0, 2>", "dram_bytes_per_launch": 123.0, "l2hint": hint, "source": "x"}, f)
            bench.run_product(args)
            r = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][0])["roofline"]
            assert r["traffic"] == expect and (r["traffic_note"] == "x") == (expect is not None)
