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
    assert "workload" in j["config"] and "model" not in j["config"]


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
