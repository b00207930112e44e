"""More GPU parity cases against the CPU oracle: active qgdFlux gradients, runcase end to end, the leastSquares degenerate
face set, the truncated-octahedron polyhedral mesh (CSR tails on the default path), the stepwise and the decomposed PCG."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
TOL_STEP = 1e-10


def _open_qgdflux_case(**kw):
    """qgdFlux p on every patch with zeroGradient U: the boundary fluid moves, so phiwStar and the qgdFlux gradient are
    NOT identically zero (with the no-slip walls of the other qgdFlux cases every term of rhoW vanishes on the patch)."""
    c = cases.case_hex3d(perturb=0.15, bcs="qgdflux", **kw)
    c.bcU[:] = cases.ZG
    return c


@pytest.mark.parametrize("kw", [dict(), dict(implicit=True), dict(model="varScModel6"), dict(scheme="reduced")],
                         ids=["explicit", "implicit", "varSc6", "reduced"])
def test_active_qgdflux_gradient_matches_oracle(qgd, oracle_mod, kw):
    c = _open_qgdflux_case(**kw)
    o = c.make_oracle(oracle_mod)
    s = c.make_solver(qgd)
    c.oracle_step(o, 60)
    s.step(60)
    nI = c.mesh.n_internal
    if kw.get("scheme") != "reduced":
        assert np.abs(o.get_face("phiwStar")[nI:]).max() > 1e-8           # the condition is active
    for f in ("rho", "rhoU", "rhoE", "U", "e", "p", "T"):
        gc, gb = s.get(f, with_bnd=True)
        oc, ob = o.get(f, with_bnd=True)
        scale = float(np.abs(oc).max())
        assert float(np.abs(gc - oc).max()) / scale < TOL_STEP, f
        assert float(np.abs(gb - ob).max()) / scale < TOL_STEP, f"boundary {f}"


def test_runcase_end_to_end_writes_the_oracle_solution(qgd, oracle_mod, tmp_path):
    """python -m qgdsolver_b200.runcase on a Sod-tube case directory: dictionaries -> device solver -> time directories;
    the written 0.02/rho, p, U equal the oracle run from the same set-up."""
    import os
    from qgdsolver_b200 import foamcase as fc
    from qgdsolver_b200 import runcase
    from test_runcase_cpu import _case_from_setup, _write_sod
    _write_sod(tmp_path)
    setup = runcase.load_case(str(tmp_path))
    written = runcase.run(setup, qgd, log=lambda *_: None)
    assert [os.path.basename(w) for w in written] == ["0.01", "0.02"]
    o = _case_from_setup(setup).make_oracle(oracle_mod)
    o.qgd_step(100)
    for name in ("rho", "p", "U", "rhoE"):
        f = fc.read_field(os.path.join(str(tmp_path), "0.02", name), setup.mesh)
        ref = o.get(name)
        assert float(np.abs(f.internal - ref).max()) / float(np.abs(ref).max()) < TOL_STEP, name


def test_least_squares_degenerate_face_set_on_device(qgd, oracle_mod):
    """qgd_mesh_set_degenerate_stencil_faces (faceSet degenerateStencilFaces, leastSquaresStencil.C:63-132): operator and
    60 solver steps against the oracle with the same forced faces."""
    c = cases.case_2d((14, 12), perturb=0.15, bcs="mixed", scheme="leastSquares")
    m = c.mesh
    nI = m.n_internal
    forced = np.arange(2, nI, 5, dtype=np.int32)
    rng = np.random.default_rng(4)
    phi, bnd = rng.random(m.n_cells), rng.random(m.n_bnd)
    bsg = m.deltaCoeffs[nI:] * (bnd - phi[m.owner[nI:]])
    o = oracle_mod.Oracle(m)
    o.set_degenerate_faces(forced)
    dm = qgd.Mesh(m)
    dm.set_degenerate_stencil_faces(forced)
    st = qgd.FvscStencil(dm, "leastSquares")
    ref = o.fvsc_grad(phi, bnd, bsg, scheme=oracle_mod.FVSC_SCHEMES["leastSquares"])
    assert float(np.abs(st.Grad(phi, bnd, bsg) - ref).max()) / float(np.abs(ref).max()) < 1e-12
    oc = c.make_oracle(oracle_mod)
    # the oracle context of the step needs the same set before its first leastSquares evaluation
    oc.set_degenerate_faces(forced)
    s = c.make_solver(qgd, dm)
    c.oracle_step(oc, 60)
    s.step(60)
    for f in ("rho", "rhoU", "rhoE"):
        a, b = s.get(f), oc.get(f)
        assert float(np.abs(a - b).max()) / float(np.abs(b).max()) < TOL_STEP, f


@pytest.mark.parametrize("bcs", ["mixed", "fixed"])
def test_truncated_octahedron_mesh_on_device(qgd, oracle_mod, bcs):
    """14 faces per cell (8 hexagons -> `other` faces, 6 squares): fvsc operators and 60 QGDFoam steps against the oracle; the
    cell->face rows exceed the ELL width 8, so the CSR tails are on the default path here."""
    c = cases.case_truncoct(bcs=bcs)
    m = c.mesh
    nI = m.n_internal
    o = oracle_mod.Oracle(m)
    dm = qgd.Mesh(m)
    st = qgd.FvscStencil(dm, "GaussVolPoint")
    rng = np.random.default_rng(3)
    cell, bnd = rng.random((m.n_cells, 3)), rng.random((m.n_bnd, 3))
    bsg = m.deltaCoeffs[nI:, None] * (bnd - cell[m.owner[nI:]])
    ref = o.fvsc_grad(cell, bnd, bsg)
    assert float(np.abs(st.Grad(cell, bnd, bsg) - ref).max()) / float(np.abs(ref).max()) < 1e-12
    oc = c.make_oracle(oracle_mod)
    s = c.make_solver(qgd, dm)
    c.oracle_step(oc, 60)
    s.step(60)
    for f in ("rho", "rhoU", "rhoE", "p"):
        a, b = s.get(f), oc.get(f)
        assert float(np.abs(a - b).max()) / float(np.abs(b).max()) < TOL_STEP, f


@pytest.mark.parametrize("precond", ["diagonal", "none"])
def test_stepwise_pcg_matches_oracle_and_the_persistent_kernel(qgd, oracle_mod, precond):
    """qgd_pcg_solve_stepwise (one kernel per phase, the building block of the multi-GPU solver) on one GPU: same solution and
    iteration count as the oracle and as the cooperative persistent kernel."""
    mesh = cases.pm.hex_box(12, 10, 8, perturb=0.1, seed=1)
    nI = mesh.n_internal
    upper = -(mesh.magSf[:nI] * mesh.deltaCoeffs[:nI])
    diag = np.zeros(mesh.n_cells)
    np.subtract.at(diag, mesh.owner[:nI], upper); np.subtract.at(diag, mesh.neighbour, upper)
    diag += 1e-3 * mesh.V / mesh.V.mean()
    b = np.random.default_rng(0).standard_normal(mesh.n_cells)
    x0 = np.zeros(mesh.n_cells)
    o = oracle_mod.Oracle(mesh)
    xr, itr, r0r, r1r = o.pcg_solve(diag, upper, b, x0, tol=1e-12, maxIter=3000, precond=oracle_mod.PRECONDS[precond])
    dm = qgd.Mesh(mesh)
    xs, its, r0s, r1s = qgd.pcg_solve(dm, diag, upper, b, x0, tol=1e-12, max_iter=3000, precond=precond, stepwise=True)
    xp, itp, _, _ = qgd.pcg_solve(dm, diag, upper, b, x0, tol=1e-12, max_iter=3000, precond=precond)
    assert abs(its - itr) <= 1 and abs(its - itp) <= 1
    assert abs(r0s - r0r) < 1e-12 * r0r and r1s < 1e-12
    assert float(np.abs(xs - xr).max()) / float(np.abs(xr).max()) < 1e-9
    assert float(np.abs(xs - xp).max()) / float(np.abs(xp).max()) < 1e-9
    # a converged start does not iterate; maxIter is honoured
    _, it0, _, _ = qgd.pcg_solve(dm, diag, upper, b, xr, tol=1e-6, max_iter=3000, precond=precond, stepwise=True)
    assert it0 == 0
    _, it5, _, _ = qgd.pcg_solve(dm, diag, upper, b, x0, tol=1e-30, max_iter=5, precond=precond, stepwise=True)
    assert it5 == 5


@pytest.mark.parametrize("n", [2, 4])
def test_decomposed_pcg_matches_oracle(n):
    """qgd_pcg_solve_multi (stepwise PCG + NCCL exchange / all-reduce) on n GPUs against the oracle's decomposed-run solver
    (skipped on boxes with fewer GPUs)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    # an exchange-list mistake would block in ncclRecv rather than fail: the worker runs under a hard timeout
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", str(29640 + n), os.path.join(root, "tests", "multi_gpu_pcg_worker.py")],
                       capture_output=True, text=True, timeout=120)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert "MPCG_ALL_OK" in r.stdout


def test_negative_state_guard_reports_the_first_bad_step(qgd, oracle_mod):
    """QGDFoam.C:142-147: the reference dumps U, e, rho when an update leaves min(e) <= 0 or min(rho) <= 0.  The device records
    the first such step; a healthy run reports 0, an unstable one (time step far beyond the stability limit) the same step at
    which the oracle's fields first turn non-positive."""
    ok = cases.case_hex3d(perturb=0.1, bcs="mixed")
    s = ok.make_solver(qgd)
    s.step(30)
    assert s.state_guard() == 0
    bad = cases.case_sod(100, dt=2e-2)
    sb = bad.make_solver(qgd)
    o = bad.make_oracle(oracle_mod)
    first = 0
    for k in range(1, 41):
        bad.oracle_step(o, 1)
        with np.errstate(invalid="ignore"):
            if first == 0 and ((o.get("e") <= 0).any() or (o.get("rho") <= 0).any()):
                first = k
    sb.step(40)
    assert first > 0 and sb.state_guard() == first
