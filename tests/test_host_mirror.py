"""The C++ host mirror of the reference interface (qgdsolver_b200/host/: fvsc::grad/div, fvscStencil::New/lookupOrNew,
QGDCoeffs::New, QGDFoam / QHDFoam loop bodies) driven through its demo binary and checked against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "qgdsolver_b200", "host")
DEMO = os.path.join(HOST, "qgd_host_demo")


@pytest.fixture(scope="module")
def demo():
    from qgdsolver_b200 import build
    build.build()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return DEMO


def _run(args):
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_cpp_hex_mesh_equals_python_polymesh(demo, tmp_path):
    """The C++ blockMesh-like generator and qgdsolver_b200.polymesh.hex_box describe the same fvMesh: identical
    addressing (bit-exact), geometry equal to rounding."""
    nx, ny, nz = 5, 4, 3
    out = str(tmp_path / "mesh.bin")
    _run([demo, "mesh", str(nx), str(ny), str(nz), out])
    m = cases.pm.hex_box(nx, ny, nz)
    with open(out, "rb") as f:
        nC, nF, nI, nP = np.fromfile(f, np.int32, 4)
        assert (nC, nF, nI, nP) == (m.n_cells, m.n_faces, m.n_internal, m.n_points)
        pts = np.fromfile(f, np.float64, 3 * nP).reshape(-1, 3)
        fv = np.fromfile(f, np.int32, 4 * nF)
        own = np.fromfile(f, np.int32, nF)
        nei = np.fromfile(f, np.int32, nI)
        C = np.fromfile(f, np.float64, 3 * nC).reshape(-1, 3)
        V = np.fromfile(f, np.float64, nC)
        Cf = np.fromfile(f, np.float64, 3 * nF).reshape(-1, 3)
        Sf = np.fromfile(f, np.float64, 3 * nF).reshape(-1, 3)
        magSf = np.fromfile(f, np.float64, nF)
        w = np.fromfile(f, np.float64, nF)
        dC = np.fromfile(f, np.float64, nF)
    assert np.array_equal(fv, m.face_verts) and np.array_equal(own, m.owner) and np.array_equal(nei, m.neighbour)
    for a, b in ((pts, m.points), (C, m.C), (V, m.V), (Cf, m.Cf), (Sf, m.Sf), (magSf, m.magSf), (w, m.weights), (dC, m.deltaCoeffs)):
        assert np.allclose(a, b, rtol=1e-13, atol=1e-15)


@pytest.mark.gpu
def test_fvsc_free_functions_through_the_host_mirror(demo):
    out = _run([demo, "fvsc", "9", "8", "7"])
    assert "PASS fvsc host mirror" in out


@pytest.mark.gpu
def test_qgdfoam_loop_through_the_host_mirror_matches_oracle(demo, oracle_mod, tmp_path):
    n, steps = (10, 9, 8), 50
    c = cases.case_hex3d(n=n)                      # uniform box, zeroGradient patches, GAS, dt 2e-4 == the demo's dictionaries
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        for a in (c.U0, c.T0, c.p0):
            np.ascontiguousarray(a, np.float64).tofile(f)
    assert "PASS qgdfoam host mirror" in _run([demo, "qgdfoam", *map(str, n), str(steps), inp, out])
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, steps)
    nC = c.mesh.n_cells
    with open(out, "rb") as f:
        rho = np.fromfile(f, np.float64, nC)
        rhoU = np.fromfile(f, np.float64, 3 * nC).reshape(-1, 3)
        rhoE = np.fromfile(f, np.float64, nC)
        p = np.fromfile(f, np.float64, nC)
    for a, b in ((rho, o.get("rho")), (rhoU, o.get("rhoU")), (rhoE, o.get("rhoE")), (p, o.get("p"))):
        assert np.abs(a - b).max() / np.abs(b).max() < 1e-10


@pytest.mark.gpu
def test_qhdfoam_loop_through_the_host_mirror_matches_oracle(demo, oracle_mod, tmp_path):
    n, steps = (16, 14), 40
    c = cases.qhd_cavity(n=n, dt=1e-3)             # == the demo's dictionaries (constTau 1e-3, DIC, tol 1e-13, qhdFlux p)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        for a in (c.U0, c.T0, c.p0):
            np.ascontiguousarray(a, np.float64).tofile(f)
    assert "PASS qhdfoam host mirror" in _run([demo, "qhdfoam", str(n[0]), str(n[1]), str(steps), inp, out])
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, steps)
    nC = c.mesh.n_cells
    with open(out, "rb") as f:
        U = np.fromfile(f, np.float64, 3 * nC).reshape(-1, 3)
        T = np.fromfile(f, np.float64, nC)
        p = np.fromfile(f, np.float64, nC)
    for a, b in ((U, o.qhd_get("U")), (T, o.qhd_get("T")), (p, o.qhd_get("p"))):
        assert np.abs(a - b).max() / np.abs(b).max() < 1e-10
