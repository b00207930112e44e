"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances: FP64 path; north_star asks <= 1e-10 relative L-inf on conserved fields after 100 steps.  Operators and
set-up quantities are checked much tighter (1e-12) because they involve no time accumulation.
"""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-10      # north_star tolerance, conserved fields after 100 steps
TOL_OP = 1e-12        # single operator application


def rel_linf(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


MESHES = {
    "hex_uniform": lambda: cases.pm.hex_box(7, 6, 5),
    "hex_perturbed": lambda: cases.pm.hex_box(7, 6, 5, perturb=0.25, grading=(2, 1, 0.5), seed=11),
    "prism": lambda: cases.pm.prism_box(4, 4, 3, perturb=0.15, seed=2),
    "poly": lambda: cases.pm.hexprism_poly(5, 4, 3, a=0.1, lz=0.4),
    "2d_z": lambda: cases.case_2d(perturb=0.2).mesh,
    "2d_y": lambda: cases.case_2d(perturb=0.2, axis=1).mesh,
    "2d_x": lambda: cases.case_2d(perturb=0.2, axis=0).mesh,
    "1d": lambda: cases.case_sod(40).mesh,
}


def _fields(mesh, k, seed):
    rng = np.random.default_rng(seed)
    nI = mesh.n_internal
    shape = (mesh.n_cells, k) if k > 1 else (mesh.n_cells,)
    cell = np.sin(3 * mesh.C[:, :1] + np.arange(k)) + 0.3 * rng.random((mesh.n_cells, k))
    bnd = np.cos(2 * mesh.Cf[nI:, 1:2] + np.arange(k)) + 0.3 * rng.random((mesh.n_bnd, k))
    # mixed patch behaviour: generic snGrad on even faces, prescribed gradient on odd faces
    bsg = mesh.deltaCoeffs[nI:, None] * (bnd - cell[mesh.owner[nI:]])
    bsg[1::2] = rng.random((mesh.n_bnd, k))[1::2]
    return cell.reshape(shape), bnd.reshape((mesh.n_bnd,) + shape[1:]), bsg.reshape((mesh.n_bnd,) + shape[1:])


@pytest.mark.parametrize("mesh_name", list(MESHES))
@pytest.mark.parametrize("scheme", ["GaussVolPoint", "reduced", "leastSquares", "leastSquaresOpt"])
def test_fvsc_operators_match_oracle(qgd, oracle_mod, mesh_name, scheme):
    mesh = MESHES[mesh_name]()
    if scheme.startswith("leastSquares") and mesh.n_geometric_d == 3:
        pytest.skip("leastSquares is rejected in 3D (fvsc.C:60-63)")
    o = oracle_mod.Oracle(mesh)
    osch = oracle_mod.FVSC_SCHEMES[scheme]
    dm = qgd.Mesh(mesh)
    st = qgd.FvscStencil(dm, scheme)
    for k in (1, 3):
        cell, bnd, bsg = _fields(mesh, k, 100 + k)
        assert rel_linf(st.Grad(cell, bnd, bsg), o.fvsc_grad(cell, bnd, bsg, scheme=osch)) < TOL_OP
    for k in (3, 9):
        cell, bnd, bsg = _fields(mesh, k, 200 + k)
        assert rel_linf(st.Div(cell, bnd, bsg), o.fvsc_div(cell, bnd, bsg, scheme=osch)) < TOL_OP


@pytest.mark.parametrize("mesh_name", list(MESHES))
def test_qgd_lengths_match_oracle(qgd, oracle_mod, mesh_name):
    mesh = MESHES[mesh_name]()
    o = oracle_mod.Oracle(mesh)
    dm = qgd.Mesh(mesh)
    kind = mesh.patch_kind_per_bface()
    keep = np.ones(mesh.n_faces, bool)
    keep[mesh.n_internal:] = kind != 1
    assert rel_linf(dm.hQGDf()[keep], o.hQGDf()[keep]) < 1e-14
    assert rel_linf(dm.hQGD(), o.hQGD()) < 1e-14


STEP_CASES = {
    "hex_zg": lambda: cases.case_hex3d(),
    "hex_perturbed_mixed": lambda: cases.case_hex3d(perturb=0.2, grading=(2, 1, 0.5), bcs="mixed"),
    "hex_fixed_offsetgas": lambda: cases.case_hex3d(bcs="fixed", gas=dict(cases.GAS, Tref=0.2, Hsref=0.1)),
    "hex_qgdflux_walls": lambda: cases.case_hex3d(perturb=0.1, bcs="qgdflux"),
    "hex_adjust_dt": lambda: cases.case_hex3d(bcs="fixed", adjust_time_step=True, dt=1e-3, max_co=0.1, c_tau=0.3),
    "hex_no_quirk": lambda: cases.case_hex3d(perturb=0.1, bcs="mixed", energy_ddt_rhoE_quirk=False, alpha_eff_gamma_factor=False),
    "hex_reduced": lambda: cases.case_hex3d(perturb=0.1, bcs="mixed", scheme="reduced"),
    "prism_fixed": lambda: cases.case_prism(bcs="fixed"),
    "poly_qgdflux": lambda: cases.case_poly(bcs="qgdflux"),
    "2d_mixed": lambda: cases.case_2d(perturb=0.2, bcs="mixed"),
    "2d_y_fixed": lambda: cases.case_2d(perturb=0.1, bcs="fixed", axis=1),
    "sod_1d": lambda: cases.case_sod(200),
    "hex_model1n": lambda: cases.case_hex3d(perturb=0.15, bcs="mixed", model="constScPrModel1n"),
    "hex_model2": lambda: cases.case_hex3d(perturb=0.15, bcs="mixed", model="constScPrModel2"),
    "2d_model1n_qgdflux": lambda: cases.case_2d(perturb=0.1, bcs="qgdflux", model="constScPrModel1n"),
    "prism_model1n_adjust": lambda: cases.case_prism(bcs="fixed", model="constScPrModel1n", adjust_time_step=True, max_co=0.1),
    # constScPrModel1n with a non-uniform alphaQGD field (first step: I(alphaQGD) hQGDf / I(c), constScPrModel1n.C:104-105)
    "hex_model1n_alpha_field": lambda: _alpha_case(cases.case_hex3d(perturb=0.15, bcs="mixed", model="constScPrModel1n")),
    "2d_model1_alpha_field_qgdflux": lambda: _alpha_case(cases.case_2d(perturb=0.1, bcs="qgdflux")),
    # varScModel6 / varScModel7: ScQGD from the pressure-jump sensor (varScModel6.C:210-269, varScModel7.C:176-254)
    "hex_varSc6_mixed": lambda: cases.case_hex3d(perturb=0.15, bcs="mixed", model="varScModel6"),
    "poly_varSc6_qgdflux": lambda: cases.case_poly(bcs="qgdflux", model="varScModel6"),
    "hex_varSc7_fixed_clamped": lambda: cases.case_hex3d(perturb=0.1, bcs="fixed", model="varScModel7",
                                                         varsc=dict(cSc1=3.0, minSc=0.02, maxSc=0.4, const_sc_cells=np.arange(5, 300, 7))),
    "2d_varSc7_adjust": lambda: cases.case_2d(perturb=0.1, bcs="mixed", model="varScModel7", varsc=dict(cSc1=2.0, minSc=0.05),
                                              adjust_time_step=True, max_co=0.1),
    "sod_varSc7": lambda: cases.case_sod(200, model="varScModel7", varsc=dict(cSc1=1.0, minSc=0.05, maxSc=1.0)),
    "prism_varSc6_implicit": lambda: cases.case_prism(bcs="fixed", model="varScModel6", implicit=True),
    # implicitDiffusion true (the reference's default): QGDUEqn.H:54-75, QGDEEqn.H:53-64
    "hex_implicit": lambda: cases.case_hex3d(perturb=0.15, bcs="mixed", implicit=True),
    "hex_implicit_fixed_diag": lambda: cases.case_hex3d(bcs="fixed", implicit=True, diff_solver=dict(precond="diagonal"),
                                                        gas=dict(cases.GAS, mu=5e-3)),
    "2d_implicit_qgdflux": lambda: cases.case_2d(perturb=0.1, bcs="qgdflux", implicit=True),
    "prism_implicit_adjust": lambda: cases.case_prism(bcs="fixed", implicit=True, adjust_time_step=True, max_co=0.1),
    "sod_implicit": lambda: cases.case_sod(200, implicit=True),
    # fvsc leastSquares (2D / 1D only): extendedFaceStencil*.C
    "2d_leastSquares": lambda: cases.case_2d(perturb=0.2, bcs="mixed", scheme="leastSquares"),
    "2d_y_leastSquares_implicit": lambda: cases.case_2d(perturb=0.1, bcs="fixed", axis=1, scheme="leastSquares", implicit=True),
    "sod_leastSquares": lambda: cases.case_sod(200, scheme="leastSquares"),
    "2d_leastSquaresOpt": lambda: cases.case_2d(perturb=0.2, bcs="qgdflux", scheme="leastSquaresOpt"),
    # aspect ratio 10: det(G) < 1 on the x-normal faces -> degenerate-stencil branch (extendedFaceStencilCalculateWeights.C:136-140)
    "2d_leastSquares_degenerate": lambda: cases.case_2d(n=(6, 60), bcs="mixed", scheme="leastSquares"),
    "2d_leastSquaresOpt_degenerate": lambda: cases.case_2d(n=(6, 60), bcs="mixed", scheme="leastSquaresOpt"),
    # slip / symmetryPlane velocity [OF-v2312 basicSymmetry]: U_b = U_P - n (n . U_P) on skewed boundary faces
    "hex_slip_perturbed": lambda: _slip_case(cases.case_hex3d(perturb=0.2, bcs="zg")),
    "prism_slip_model1n": lambda: _slip_case(cases.case_prism(bcs="zg", model="constScPrModel1n")),
    "2d_slip_adjust": lambda: _slip_case(cases.case_2d(perturb=0.2, bcs="zg", adjust_time_step=True, max_co=0.1)),
    "poly_slip_reduced": lambda: _slip_case(cases.case_poly(bcs="zg", scheme="reduced")),
    # the other thermoType instantiations of psiQGDThermos.C:65-111: powerLaw / sutherland transport, eConst thermo
    "hex_powerLaw": lambda: _thermo_case(cases.case_hex3d(perturb=0.15, bcs="mixed", gas=dict(cases.GAS, mu=3e-3)), power_law=dict(mu0=3e-3, T0=0.7, k=0.76)),
    "prism_sutherland_model2": lambda: _thermo_case(cases.case_prism(bcs="fixed", model="constScPrModel2", gas=dict(cases.GAS, mu=3e-3)),
                                                    sutherland=dict(As=2.5e-3, Ts=0.4)),
    "2d_sutherland_implicit": lambda: _thermo_case(cases.case_2d(perturb=0.1, bcs="fixed", implicit=True), sutherland=dict(As=2.5e-3, Ts=0.4)),
    "hex_eConst": lambda: _thermo_case(cases.case_hex3d(perturb=0.1, bcs="mixed", gas=dict(cases.GAS, Tref=0.2)), e_const=dict(Cv=1.5, Esref=0.05)),
    "sod_eConst_adjust": lambda: _thermo_case(cases.case_sod(200, adjust_time_step=True, max_co=0.2), e_const=dict(Cv=2.5, Esref=0.0)),
    # BASELINE configs[1] in miniature: Mach-3 forward-facing step, slip walls + step (polymesh.forward_step)
    "forward_step_30": lambda: cases.case_forward_step(n=30),
}


def _alpha_case(c):
    """0/alphaQGD present and non-uniform (QGDCoeffs.C:119-143)"""
    x = c.mesh.C
    c.alphaQGD = 0.35 + 0.3 * np.sin(4 * x[:, 0]) ** 2 + 0.1 * np.random.default_rng(9).random(c.mesh.n_cells)
    return c


def _thermo_case(c, power_law=None, sutherland=None, e_const=None):
    c.power_law, c.sutherland, c.e_const = power_law, sutherland, e_const
    return c


def _slip_case(c):
    """every second patch becomes a slip wall (p: qgdFlux on one of them), the others keep zeroGradient"""
    for i in range(1, len(c.mesh.patches), 2):
        if c.mesh.patches[i].kind != 1:
            c.bcU[i] = cases.SLIP
    c.bcP[1] = cases.QF if c.mesh.patches[1].kind != 1 else c.bcP[1]
    return c


@pytest.mark.parametrize("name", list(STEP_CASES))
def test_qgdfoam_100_steps_match_oracle(qgd, oracle_mod, name):
    c = STEP_CASES[name]()
    o = c.make_oracle(oracle_mod)
    s = c.make_solver(qgd)
    # initial state (thermo construction on the device)
    for f in ("rho", "rhoU", "rhoE", "e", "p", "T", "c", "mu"):
        assert rel_linf(s.get(f), o.get(f)) < 1e-13, f"init {f}"
    c.oracle_step(o, 100)
    s.step(100)
    for f in ("rho", "rhoU", "rhoE", "U", "e", "p", "T", "mu", "tauQGD") + (("ScQGD",) if c.model.startswith("varSc") else ()):
        gc, gb = s.get(f, with_bnd=True)
        oc, ob = o.get(f, with_bnd=True)
        assert np.isfinite(gc).all()
        scale = float(np.abs(oc).max())
        assert float(np.abs(gc - oc).max()) / scale < TOL_STEP, f"{name}: cells {f}"
        kind = c.mesh.patch_kind_per_bface()
        if (kind != 1).any():        # boundary values, relative to the magnitude of the cell field
            assert float(np.abs(gb[kind != 1] - ob[kind != 1]).max()) / scale < TOL_STEP, f"{name}: boundary {f}"
    if c.opts["adjust_time_step"]:
        assert abs(s.scalars()["deltaT"] - o.deltaT()) < 1e-10 * o.deltaT()
    if c.implicit:
        assert all(0 <= it < c.diff_solver["max_iter"] for it in s.diffusion_iterations())


@pytest.mark.parametrize("name", ["hex_perturbed_mixed", "2d_mixed", "sod_1d", "hex_implicit", "prism_model1n", "qhd_cavity2d",
                                  "qhd_cavity3d_H2bynu", "hex_varSc7", "poly_varSc6", "hex_sources", "2d_sources_implicit",
                                  "qhd_scalar_transport2d", "qhd_scalar_transport3d_adjust"])
def test_cuda_path_reproduces_committed_golden_fixtures(qgd, name):
    """tests/golden/*.npz (written by tests/golden/make_golden.py from the oracle) against the CUDA path, no oracle run."""
    import os
    from test_oracle_kat import GOLDEN, GOLDEN_CASES, golden_fields
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    c = GOLDEN_CASES[name]()
    s = c.make_solver(qgd)
    s.step(int(z["steps"]))
    for f, v in golden_fields(c, s).items():
        assert float(np.abs(v - z[f]).max()) / float(np.abs(z[f]).max()) < TOL_STEP, f


def test_fluxes_match_oracle_after_one_step(qgd, oracle_mod):
    c = cases.case_hex3d(perturb=0.2, bcs="mixed")
    o = c.make_oracle(oracle_mod)
    s = c.make_solver(qgd)
    s.set_pipeline(1)
    with pytest.raises(qgd.QGDError):          # the pipelined step keeps no flux array in HBM
        s.get_flux(0)
    s.set_pipeline(0)
    c.oracle_step(o, 1)
    s.step(1)
    Fm = o.get_face("phiJm")
    FU = o.get_face("phiJmU") + o.get_face("phiP") - o.get_face("phiPi")
    FE = o.get_face("phiJmH") + o.get_face("phiQ") - o.get_face("phiPiU")
    assert rel_linf(s.get_flux(0), Fm) < 1e-12
    assert rel_linf(s.get_flux(1), FU) < 1e-12
    assert rel_linf(s.get_flux(2), FE) < 1e-12


PIPE_CASES = {
    "hex": lambda: cases.case_hex3d(n=(14, 12, 10), perturb=0.15, bcs="mixed"),
    "prism": lambda: cases.case_prism(n=(7, 6, 6), bcs="fixed"),
    "poly": lambda: cases.case_poly(n=(7, 7, 5), bcs="qgdflux"),
    "2d": lambda: cases.case_2d(n=(40, 36), perturb=0.2, bcs="mixed"),
    "sod": lambda: cases.case_sod(400),
}


@pytest.mark.parametrize("name", list(PIPE_CASES))
def test_pipelined_step_is_bitwise_identical_to_two_kernel_step(qgd, name):
    """k_face_cell_pipeline (flux ring, flag-ordered work queue) against k_face_flux + k_cell_update on the same inputs:
    same arithmetic in the same order per face and per cell -> bit-identical state.  Small chunks / rings force many
    ring wrap-arounds and producer/consumer waits."""
    c = PIPE_CASES[name]()
    ref = c.make_solver(qgd)
    ref.set_pipeline(0)
    ref.step(25)
    nI = c.mesh.n_internal
    for chunk, lag, ring in ((0, -1, 0), (32, 1, 0), (64, 3, 0), (32, 0, max(64, nI // 7)), (96, 2, max(256, nI // 3))):
        s = c.make_solver(qgd)
        try:
            s.set_pipeline(1, chunk, lag, ring)
        except qgd.QGDError as e:              # ring too small for this mesh band width: a legitimate refusal
            assert ring and "ring too small" in e.message
            continue
        assert s.get_pipeline()["mode"] == 1
        s.step(25)
        for f in ("rho", "rhoU", "rhoE", "e", "p", "T", "mu"):
            a, b = s.get(f, with_bnd=True), ref.get(f, with_bnd=True)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (name, chunk, lag, ring, f)


def test_pipeline_wraps_the_ring_at_scale(qgd):
    """64^3 box: default plan vs a deliberately small ring (many wrap-arounds, real waits) vs the two-kernel form."""
    mesh = cases.pm.hex_box(64, 64, 64)
    c = cases._with_bcs(mesh, "mixed", cases.GAS, 2e-4)
    out = []
    for cfg in ((0,), (1, 0, -1, 0), (1, 256, 8, 64 * 64 * 3 * 6)):
        s = c.make_solver(qgd)
        s.set_pipeline(*cfg)
        s.step(30)
        out.append([s.get(f) for f in ("rho", "rhoU", "rhoE")])
        if cfg[0]:
            info = s.get_pipeline()
            assert info["mode"] == 1 and info["n_chunks"] > 100
    for o in out[1:]:
        for a, b in zip(o, out[0]):
            assert np.array_equal(a, b)


def test_step_host_equals_device_resident_loop(qgd):
    c = cases.case_hex3d(perturb=0.1, bcs="mixed")
    s1 = c.make_solver(qgd)
    s2 = c.make_solver(qgd)
    s1.step(10)
    n = c.mesh.n_cells
    st = {k: np.zeros((n, 3) if k in ("U", "rhoU") else n) for k in ("rho", "U", "e", "p", "T", "rhoU", "rhoE", "mu")}
    s2.step_host(0, None, st)                      # download the initial state
    for _ in range(10):
        s2.step_host(1, st, st)                    # upload, one step, download - every step
    for f in ("rho", "rhoU", "rhoE", "e", "p"):
        assert rel_linf(s2.get(f), s1.get(f)) < 1e-12
        assert rel_linf(st[f].reshape(s1.get(f).shape), s1.get(f)) < 1e-12


def test_step_fields_host_is_the_reference_restart(qgd, oracle_mod):
    """qgd_qgdfoam_step_fields_host: U, T, p in -> one step -> U, T, p (+ conserved) out, every step.  Reference semantics:
    a run that is written and restarted each step (createFields.H:3-109 re-creates the state from the three fields), so
    the comparison is an oracle re-initialised from the oracle's own U, T, p before every step."""
    c = cases.case_hex3d(perturb=0.15, bcs="mixed", model="constScPrModel1n")
    s = c.make_solver(qgd)
    n = c.mesh.n_cells
    fl = dict(U=np.array(c.U0, float), T=np.array(c.T0, float), p=np.array(c.p0, float), rho=np.zeros(n), rhoU=np.zeros((n, 3)), rhoE=np.zeros(n))
    Uo, To, po = c.U0, c.T0, c.p0
    for _ in range(6):
        s.step_fields_host(1, fl, fl)
        r = cases.Case(c.mesh, np.ascontiguousarray(Uo), To, po, c.bcU, c.bcT, c.bcP, c.bvU, c.bvT, c.bvP, gas=c.gas, dt=c.dt, model=c.model)
        o = r.make_oracle(oracle_mod)
        r.oracle_step(o, 1)
        Uo, To, po = o.get("U"), o.get("T"), o.get("p")
        for f, ref in (("U", Uo), ("T", To), ("p", po), ("rho", o.get("rho")), ("rhoU", o.get("rhoU")), ("rhoE", o.get("rhoE"))):
            assert rel_linf(fl[f], ref) < 1e-12, f
    # in = None keeps the device state: equals the device-resident loop
    s2 = c.make_solver(qgd)
    s2.step_fields_host(0, dict(U=c.U0, T=c.T0, p=c.p0), None)
    s2.step_fields_host(5, None, fl)
    s3 = c.make_solver(qgd)
    s3.step(5)
    assert np.array_equal(fl["rho"], s3.get("rho")) and np.array_equal(fl["p"], s3.get("p"))


def test_multi_tile_tma_ring_matches_oracle_64cubed(qgd, oracle_mod):
    """64^3 hex box x 100 steps against the oracle: 786 432 internal faces = 3024 tiles of 256, so every CTA of the persistent
    TMA face kernel loops ~10 times over its 2-stage ring (mbarrier phase flips, stage refills) - the small cases above
    never leave the first iteration."""
    import os
    mesh = cases.pm.hex_box(64, 64, 64)
    c = cases._with_bcs(mesh, "mixed", cases.GAS, 2e-4 * 4)
    o = c.make_oracle(oracle_mod, n_threads=os.cpu_count() or 1)
    s = c.make_solver(qgd)
    assert s.face_kernel()[0] == "k_face_flux_tma"
    c.oracle_step(o, 100)
    s.step(100)
    for f in ("rho", "rhoU", "rhoE", "U", "e", "p", "T"):
        assert rel_linf(s.get(f), o.get(f)) < TOL_STEP, f


def test_tma_face_kernel_on_odd_face_counts(qgd):
    """Face counts that are not a multiple of 2 used to fall back to the register-prefetch kernel (the SoA columns of G / Sf
    were then 8-byte aligned only); with the padded column stride the TMA kernel runs on any mesh and stays bit-identical
    to the plain-load kernel."""
    import os
    import subprocess
    import sys
    mesh = cases.pm.hex_box(33, 17, 8)                  # odd x odd x even like the 129x129x256 sub-mesh: odd number of faces
    assert mesh.n_faces % 2 == 1
    c = cases._with_bcs(mesh, "mixed", cases.GAS, 2e-4)
    s = c.make_solver(qgd)
    assert s.face_kernel() == ("k_face_flux_tma", 3)
    s.step(20)
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np, cases; from qgdsolver_b200 import api; api.init(0);"
            "c = cases._with_bcs(cases.pm.hex_box(33, 17, 8), 'mixed', cases.GAS, 2e-4); s = c.make_solver(api);"
            "assert s.face_kernel()[0] == 'k_face_flux'; s.step(20); np.save(sys.argv[1], np.stack([s.get('rho'), s.get('rhoE')]))")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join("/tmp", "qgd_odd_faces.npy")
    r = subprocess.run([sys.executable, "-c", code % (root, os.path.join(root, "tests")), out], env=dict(os.environ, QGD_FACE_TMA="0"),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = np.load(out)
    assert np.array_equal(ref[0], s.get("rho")) and np.array_equal(ref[1], s.get("rhoE"))


def test_forward_step_at_full_size_matches_oracle(qgd, oracle_mod):
    """BASELINE configs[1]: Mach-3 forward-facing step, n = 640 -> 1 032 192 hex cells (1920 x 640 minus the step), slip walls,
    explicit, 100 steps from the impulsive start against the oracle (all host threads)."""
    import os
    c = cases.case_forward_step(n=640)
    assert c.mesh.n_cells == 1920 * 640 - 1536 * 128
    o = c.make_oracle(oracle_mod, n_threads=os.cpu_count() or 1)
    s = c.make_solver(qgd)
    c.oracle_step(o, 100)
    s.step(100)
    for f in ("rho", "rhoU", "rhoE", "p"):
        assert rel_linf(s.get(f), o.get(f)) < TOL_STEP, f
    assert float(np.abs(o.get("p") - 1.0).max()) > 1.0              # the shock in front of the step has formed


def test_slip_is_refused_with_implicit_diffusion(qgd):
    c = _slip_case(cases.case_hex3d(bcs="zg", implicit=True))
    with pytest.raises(qgd.QGDError) as e:
        c.make_solver(qgd)
    assert e.value.code == qgd.ERR_UNSUPPORTED and "slip" in e.value.message


def test_conservation_and_finiteness_at_scale(qgd):
    """Size-independent properties at a size the oracle is not run at: with qgdFlux walls (U=0) the mass flux through
    every boundary face vanishes, so total mass is conserved to round-off; energy/momentum stay finite."""
    mesh = cases.pm.hex_box(96, 96, 96)
    c = cases._with_bcs(mesh, "qgdflux", cases.GAS, 2e-4)
    s = c.make_solver(qgd)
    m0 = float((s.get("rho") * mesh.V).sum())
    s.step(50)
    rho = s.get("rho")
    assert np.isfinite(rho).all() and np.isfinite(s.get("rhoE")).all()
    assert abs(float((rho * mesh.V).sum()) - m0) < 1e-12 * m0


def test_uniform_state_is_preserved(qgd):
    mesh = cases.pm.hex_box(12, 10, 8, perturb=0.2, seed=4)
    n = mesh.n_cells
    bc = cases.uniform_bcs(mesh)
    U0 = np.tile([0.3, -0.2, 0.1], (n, 1))
    c = cases.Case(mesh, U0, np.full(n, 1.0 / 1.4), np.full(n, 1.0 / 1.4), *bc)
    s = c.make_solver(qgd)
    r0, u0 = s.get("rho").copy(), s.get("U").copy()
    s.step(20)
    assert rel_linf(s.get("rho"), r0) < 1e-12
    assert rel_linf(s.get("U"), u0) < 1e-12


def test_error_behaviour_matches_reference_messages(qgd):
    mesh = cases.pm.hex_box(3, 3, 3)
    dm = qgd.Mesh(mesh)
    with pytest.raises(qgd.QGDError) as e:
        qgd.FvscStencil(dm, "noSuchScheme")            # fvscStencil.C:72-78
    assert e.value.code == qgd.ERR_UNKNOWN_MODEL and "Unknown Model type noSuchScheme" in e.value.message
    assert "Valid model types are:" in e.value.message and "GaussVolPoint" in e.value.message
    with pytest.raises(qgd.QGDError) as e:
        qgd.FvscStencil(dm, "leastSquares")            # fvsc.C:60-63
    assert "Can't use leastSquares or leastSquaresOpt in 3D case." in e.value.message
    # fvsc.C:65-82: GaussVolPoint refuses wedge patches combined with prism cells; hexes with a wedge patch are fine
    pr = cases.pm.prism_box(3, 3, 2)
    pr.patches[0].kind = cases.pm.PATCH_WEDGE
    with pytest.raises(qgd.QGDError) as e:
        qgd.FvscStencil(qgd.Mesh(pr), "GaussVolPoint")
    assert "does not support solving axisymmetric cases with wedge BC and prism cells" in e.value.message
    qgd.FvscStencil(qgd.Mesh(pr), "reduced")
    hx = cases.pm.hex_box(3, 3, 2)
    hx.patches[0].kind = cases.pm.PATCH_WEDGE
    qgd.FvscStencil(qgd.Mesh(hx), "GaussVolPoint")
    with pytest.raises(qgd.QGDError) as e:
        qgd.QGDFoam(dm, R=1.0, Cp=3.5, qgd_coeffs="noModel")      # QGDCoeffs.C:72-78
    assert "Unknown QGD coeffs evaluation approach type noModel" in e.value.message
    with pytest.raises(qgd.QGDError) as e:
        s = qgd.QGDFoam(dm, R=1.0, Cp=3.5)
        s.step(1)
    assert e.value.code == qgd.ERR_STATE


def test_case_read_from_disk_runs_like_the_in_memory_case(qgd, oracle_mod, tmp_path):
    """polyMesh + 0/U,T,p written in OpenFOAM ASCII format, read back with foamcase and stepped on the device."""
    from qgdsolver_b200 import foamcase as fc
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="mixed")
    m = c.mesh
    names = [p.name for p in m.patches]
    code = {0: "fixedValue", 1: "zeroGradient", 3: "qgdFlux"}
    fc.write_polymesh(m, str(tmp_path))
    fc.write_field(str(tmp_path / "0" / "U"), m, "U", c.U0, {n: code[int(k)] for n, k in zip(names, c.bcU)}, c.bvU)
    fc.write_field(str(tmp_path / "0" / "T"), m, "T", c.T0, {n: code[int(k)] for n, k in zip(names, c.bcT)}, c.bvT)
    fc.write_field(str(tmp_path / "0" / "p"), m, "p", c.p0, {n: code[int(k)] for n, k in zip(names, c.bcP)}, c.bvP)
    mesh = fc.read_polymesh(str(tmp_path))
    f = fc.read_case_fields(str(tmp_path), mesh)
    (kU, vU), (kT, vT), (kP, vP) = (fc.bc_arrays(mesh, f[n]) for n in ("U", "T", "p"))
    c2 = cases.Case(mesh, f["U"].internal, f["T"].internal, f["p"].internal, kU, kT, kP, vU, vT, vP, gas=c.gas, dt=c.dt)
    s = c2.make_solver(qgd)
    o = c.make_oracle(oracle_mod)
    s.step(30)
    c.oracle_step(o, 30)
    for fld in ("rho", "rhoU", "rhoE"):
        assert rel_linf(s.get(fld), o.get(fld)) < TOL_STEP


def test_ell_tails_and_degenerate_sizes(qgd, oracle_mod):
    """Edge cases of the device data layout: (a) QGD_ELL_MAXW=4 forces every stencil row of a hex / polyhedral mesh through
    the CSR tail path (cells with more than W faces, points with more than W cells, PCG rows), run in a fresh process;
    (b) the smallest meshes: a single cell (no internal face) and a 2-cell line."""
    import os
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ell_tail_worker.py")
    r = subprocess.run([sys.executable, worker], env=dict(os.environ, QGD_ELL_MAXW="4"), capture_output=True, text=True, timeout=600)
    assert "TAILS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    for n in ((1, 1, 1), (2, 1, 1)):
        c = cases.case_hex3d(n=n, bcs="fixed")
        o = c.make_oracle(oracle_mod)
        s = c.make_solver(qgd)
        c.oracle_step(o, 10)
        s.step(10)
        for f in ("rho", "rhoU", "rhoE"):
            assert rel_linf(s.get(f), o.get(f)) < TOL_STEP


@pytest.mark.parametrize("name", ["hex_perturbed_mixed", "poly_qgdflux", "hex_implicit", "hex_varSc7_fixed_clamped"])
def test_source_matrices_match_oracle(qgd, oracle_mod, name):
    """rhoSu / rhoUSu / rhoESu (createZeroSources.H:28-44; QGDRhoEqn.H:46, QGDUEqn.H:62,85, QGDEEqn.H:60,71) through
    qgd_qgdfoam_set_sources: 60 steps with smooth sources, then 20 steps after clearing them."""
    c = STEP_CASES[name]()
    m = c.mesh
    o = c.make_oracle(oracle_mod)
    s = c.make_solver(qgd)
    V, x = m.V, m.C
    act = (m.geometric_d > 0).astype(float)
    suRho = 0.3 * V * np.sin(3 * x[:, 0]) ** 2
    suU = (V * 0.2)[:, None] * np.stack([np.cos(2 * x[:, 1]), np.sin(x[:, 0] + x[:, 2]), 0.5 * np.cos(x[:, 0])], 1) * act
    suE = 0.5 * V * (1.0 + np.cos(2 * x[:, 0] + x[:, 1]))
    o.qgd_set_sources(suRho, suU, suE)
    s.set_sources(suRho, suU, suE)
    c.oracle_step(o, 60)
    s.step(60)
    for f in ("rho", "rhoU", "rhoE", "U", "e", "p"):
        assert rel_linf(s.get(f), o.get(f)) < TOL_STEP, f
    base = STEP_CASES[name]().make_oracle(oracle_mod)
    STEP_CASES[name]().oracle_step(base, 60)
    assert rel_linf(o.get("rho"), base.get("rho")) > 1e-4          # the sources did something
    o.qgd_set_sources(None, suU, None)                              # momentum source alone, then none
    s.set_sources(None, suU, None)
    c.oracle_step(o, 10)
    s.step(10)
    o.qgd_set_sources()
    s.set_sources()
    c.oracle_step(o, 10)
    s.step(10)
    for f in ("rho", "rhoU", "rhoE", "U", "e", "p"):
        assert rel_linf(s.get(f), o.get(f)) < TOL_STEP, f
