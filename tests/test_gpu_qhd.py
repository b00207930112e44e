"""GPU parity tests of the QHDFoam path and the device PCG (through the C-ABI) against the CPU oracle.

Tolerances: FP64.  The PCG solution is compared at the solver's own tolerance (the device reductions sum in a different
order than the sequential oracle, so iteration counts may differ by a few near the stopping threshold); the time loop is
run with a tight pressure tolerance (1e-13, SURVEY 8d "tighten for parity") and compared at the north_star tolerance
1e-10 relative L-inf after 100 steps.
"""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-10


def rel_linf(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _poisson(mesh, seed):
    """SPD LDU matrix of a variable-coefficient Laplacian with a few Dirichlet-like diagonal boosts."""
    rng = np.random.default_rng(seed)
    nI = mesh.n_internal
    k = 0.5 + rng.random(nI)
    upper = -(k * mesh.magSf[:nI] * mesh.nonOrthDeltaCoeffs[:nI])
    diag = np.zeros(mesh.n_cells)
    np.add.at(diag, mesh.owner[:nI], -upper)
    np.add.at(diag, mesh.neighbour, -upper)
    boost = rng.random(mesh.n_cells) < 0.1
    diag[boost] *= 1.5
    diag[0] *= 2.0
    x_true = np.sin(3 * mesh.C[:, 0]) + np.cos(2 * mesh.C[:, 1]) + 0.3 * rng.random(mesh.n_cells)
    b = diag * x_true
    np.add.at(b, mesh.owner[:nI], upper * x_true[mesh.neighbour])
    np.add.at(b, mesh.neighbour, upper * x_true[mesh.owner[:nI]])
    return diag, upper, b, x_true


PCG_MESHES = {
    "hex3d": lambda: cases.pm.hex_box(9, 8, 7, perturb=0.2, seed=3),
    "hex2d": lambda: cases.pm.hex_box(24, 20, 1, lengths=(1.0, 1.0, 0.1), patch_kinds={"zMin": "empty", "zMax": "empty"}),
    "prism": lambda: cases.pm.prism_box(5, 4, 4, perturb=0.1, seed=5),
    "poly": lambda: cases.pm.hexprism_poly(5, 5, 4, a=0.1, lz=0.5),
    "line": lambda: cases.case_sod(64).mesh,
}


@pytest.mark.parametrize("mesh_name", list(PCG_MESHES))
@pytest.mark.parametrize("precond", ["DIC", "diagonal", "none"])
def test_pcg_matches_oracle(qgd, oracle_mod, mesh_name, precond):
    mesh = PCG_MESHES[mesh_name]()
    diag, upper, b, x_true = _poisson(mesh, 17)
    o = oracle_mod.Oracle(mesh)
    x0 = np.zeros(mesh.n_cells)
    xo, ito, r0o, r1o = o.pcg_solve(diag, upper, b, x0, tol=1e-12, relTol=0.0, maxIter=2000,
                                    precond=oracle_mod.PRECONDS[precond])
    dm = qgd.Mesh(mesh)
    xg, itg, r0g, r1g = qgd.pcg_solve(dm, diag, upper, b, x0, tol=1e-12, rel_tol=0.0, max_iter=2000, precond=precond)
    assert abs(r0g - r0o) < 1e-12 * r0o                      # same normFactor / initial residual definition
    assert r1g < 1e-12 and abs(itg - ito) <= max(2, ito // 20)
    assert rel_linf(xg, xo) < 1e-9 and rel_linf(xg, x_true) < 1e-9


def test_pcg_iterates_are_identical_for_a_fixed_iteration_count(qgd, oracle_mod):
    """With maxIter fixed (no convergence race) the k-th iterate agrees with the oracle to rounding: same algorithm,
    same operation order inside the DIC sweeps and the row sums."""
    mesh = PCG_MESHES["hex3d"]()
    diag, upper, b, _ = _poisson(mesh, 23)
    o = oracle_mod.Oracle(mesh)
    dm = qgd.Mesh(mesh)
    x0 = 0.1 * np.cos(mesh.C[:, 2])
    for precond in ("DIC", "diagonal"):
        for k in (1, 2, 7):
            xo, ito, _, r1o = o.pcg_solve(diag, upper, b, x0, tol=0.0, relTol=0.0, maxIter=k, precond=oracle_mod.PRECONDS[precond])
            xg, itg, _, r1g = qgd.pcg_solve(dm, diag, upper, b, x0, tol=0.0, rel_tol=0.0, max_iter=k, precond=precond)
            assert ito == itg == k
            assert rel_linf(xg, xo) < 1e-12 and abs(r1g - r1o) < 1e-10 * r1o


# small tiles are swept by one warp each (8 in flight per CTA), tiles beyond 12 KB of rows by the whole CTA: ("hex2d", 300), ("hex3d", 260)
@pytest.mark.parametrize("mesh_name,target", [("hex3d", 40), ("hex2d", 64), ("prism", 30), ("poly", 25), ("line", 8), ("hex2d", 300), ("hex3d", 260)])
def test_block_local_dic_matches_the_decomposed_run_oracle(qgd, oracle_mod, mesh_name, target):
    """DIC blocks (qgd_mesh_make_pcg_blocks): the preconditioner is factorised and swept per block by one CTA in shared memory;
    the oracle is the decomposed-run solver or_pcg_solve_blocks with the same cell -> block map (what `mpirun -np N` does with
    N = number of blocks).  Fixed iteration counts agree to rounding, the converged solves agree in iterations and solution;
    the stepwise (one kernel per phase) solver takes the same preconditioner."""
    mesh = PCG_MESHES[mesh_name]()
    diag, upper, b, x_true = _poisson(mesh, 31)
    o = oracle_mod.Oracle(mesh)
    dm = qgd.Mesh(mesh)
    blk = dm.make_pcg_blocks(target)
    assert blk.min() == 0 and np.bincount(blk).max() <= target and blk.max() + 1 >= mesh.n_cells // target
    x0 = 0.1 * np.cos(mesh.C[:, 0])
    for k in (1, 3, 8):
        xo, ito, _, r1o = o.pcg_solve(diag, upper, b, x0, tol=0.0, relTol=0.0, maxIter=k, precond=2, cell_block=blk)
        xg, itg, _, r1g = qgd.pcg_solve(dm, diag, upper, b, x0, tol=0.0, rel_tol=0.0, max_iter=k, precond="DIC")
        assert ito == itg == k and rel_linf(xg, xo) < 1e-12 and abs(r1g - r1o) < 1e-10 * r1o
    xo, ito, r0o, _ = o.pcg_solve(diag, upper, b, x0, tol=1e-12, maxIter=3000, precond=2, cell_block=blk)
    xg, itg, r0g, r1g = qgd.pcg_solve(dm, diag, upper, b, x0, tol=1e-12, max_iter=3000, precond="DIC")
    assert abs(itg - ito) <= max(2, ito // 20) and r1g < 1e-12 and rel_linf(xg, xo) < 1e-9 and rel_linf(xg, x_true) < 1e-9
    xs, its, _, r1s = qgd.pcg_solve(dm, diag, upper, b, x0, tol=1e-12, max_iter=3000, precond="DIC", stepwise=True)
    assert abs(its - ito) <= max(2, ito // 20) and r1s < 1e-12 and rel_linf(xs, xo) < 1e-9
    _, itd, _, _ = qgd.pcg_solve(dm, diag, upper, b, x0, tol=1e-12, max_iter=3000, precond="diagonal")
    if mesh_name != "line":
        assert itg < itd                                   # the point of the exercise
    # one block = the serial DIC: identical to the level-scheduled kernel
    dm.set_pcg_blocks(np.zeros(mesh.n_cells, np.int32))
    x1, it1, _, _ = qgd.pcg_solve(dm, diag, upper, b, x0, tol=0.0, max_iter=5, precond="DIC")
    dm.set_pcg_blocks(None)
    x2, it2, _, _ = qgd.pcg_solve(dm, diag, upper, b, x0, tol=0.0, max_iter=5, precond="DIC")
    assert it1 == it2 == 5 and rel_linf(x1, x2) < 1e-13


def test_block_dic_in_the_solvers(qgd, oracle_mod):
    """QHDFoam pressure solve and the implicit U / e solves of QGDFoam with DIC blocks against the oracle with set_pcg_blocks"""
    c = cases.qhd_cavity(n=(24, 20), dt=1e-3, perturb=0.1)
    dm = qgd.Mesh(c.mesh)
    blk = dm.make_pcg_blocks(48)
    o = c.make_oracle(oracle_mod)
    o.set_pcg_blocks(blk)
    s = c.make_solver(qgd, dm)
    c.oracle_step(o, 60)
    s.step(60)
    for f in ("U", "T", "p"):
        assert rel_linf(s.get(f), o.qhd_get(f)) < TOL_STEP, f
    gi, oi = s.solver_info(), o.qhd_solver_info()
    assert abs(gi["iters"] - oi["iters"]) <= max(2, oi["iters"] // 10)
    g = cases.case_hex3d(perturb=0.15, bcs="mixed", implicit=True)
    dg = qgd.Mesh(g.mesh)
    bg = dg.make_pcg_blocks(60)
    og = g.make_oracle(oracle_mod)
    og.set_pcg_blocks(bg)
    sg = g.make_solver(qgd, dg)
    g.oracle_step(og, 60)
    sg.step(60)
    for f in ("rho", "rhoU", "rhoE", "e"):
        assert rel_linf(sg.get(f), og.get(f)) < TOL_STEP, f


def test_pcg_relative_tolerance_and_converged_start(qgd, oracle_mod):
    mesh = PCG_MESHES["hex2d"]()
    diag, upper, b, x_true = _poisson(mesh, 5)
    dm = qgd.Mesh(mesh)
    o = oracle_mod.Oracle(mesh)
    xo, ito, _, _ = o.pcg_solve(diag, upper, b, np.zeros(mesh.n_cells), tol=0.0, relTol=1e-3, maxIter=500, precond=2)
    xg, itg, r0, r1 = qgd.pcg_solve(dm, diag, upper, b, np.zeros(mesh.n_cells), tol=0.0, rel_tol=1e-3, max_iter=500)
    assert abs(itg - ito) <= 1 and r1 < 1e-3 * r0
    xg, itg, r0, r1 = qgd.pcg_solve(dm, diag, upper, b, x_true, tol=1e-6)      # already converged: zero iterations
    assert itg == 0 and r0 < 1e-6 and np.array_equal(xg, x_true)


QHD_CASES = {
    "cavity2d_constTau_DIC": lambda: cases.qhd_cavity(n=(20, 18), dt=1e-3, perturb=0.15),
    "cavity2d_T0byGr_diag": lambda: cases.qhd_cavity(n=(20, 18), dt=1e-3, model="T0byGr", precond="diagonal",
                                                      coeffs=dict(Gr=800.0, T0=1.0)),
    "cavity2d_H2bynu_zg": lambda: cases.qhd_cavity(n=(16, 16), dt=2e-4, model="H2bynuQHD", p_bc="zg", perturb=0.1),
    "cavity2d_HbyU_adjust": lambda: cases.qhd_cavity(n=(18, 14), dt=1e-3, model="HbyUQHD", coeffs=dict(UQHD=4.0),
                                                      adjust_time_step=True, max_co=0.05, c_tau=0.4),
    "cavity3d_constTau": lambda: cases.qhd_cavity(n=(9, 8, 7), dims=3, dt=1e-3, perturb=0.1),
    "cavity3d_reduced": lambda: cases.qhd_cavity(n=(8, 8, 6), dims=3, dt=1e-3, scheme="reduced"),
    "cavity2d_refcell": lambda: cases.qhd_cavity(n=(14, 12), dt=1e-3, p_ref_cell=37, p_ref_value=0.25),
    # implicitDiffusion true (the reference's default): QHDUEqn.H:46-65, QHDTEqn.H:69-80
    "cavity2d_implicit": lambda: cases.qhd_cavity(n=(18, 16), dt=1e-3, perturb=0.1, implicit=True),
    "cavity3d_implicit_diag": lambda: cases.qhd_cavity(n=(8, 7, 6), dims=3, dt=1e-3, implicit=True, diff_solver=dict(precond="diagonal")),
    "cavity2d_implicit_adjust": lambda: cases.qhd_cavity(n=(16, 14), dt=1e-3, implicit=True, adjust_time_step=True, max_co=0.05, c_tau=0.4),
    # fvsc leastSquares / leastSquaresOpt in QHDFoam (2D only, fvsc.C:60-63): cell-stencil gradients of U, T, p
    "cavity2d_leastSquares": lambda: cases.qhd_cavity(n=(18, 16), dt=1e-3, perturb=0.15, scheme="leastSquares"),
    "cavity2d_leastSquaresOpt_implicit": lambda: cases.qhd_cavity(n=(16, 14), dt=1e-3, perturb=0.1, scheme="leastSquaresOpt", implicit=True),
    "cavity2d_leastSquares_degenerate": lambda: cases.qhd_cavity(n=(6, 48), dt=5e-4, scheme="leastSquares", model="H2bynuQHD"),
}


def _qhd_fixed_p_case():
    c = cases.qhd_cavity(n=(12, 12, 5), dims=3, dt=1e-3, perturb=0.1)
    names = [p.name for p in c.mesh.patches]
    c.bcP[names.index("yMax")] = cases.FV          # p fixes value on one patch: no reference cell
    pid = c.mesh.patch_id_per_bface()
    c.bvP[pid == names.index("yMax")] = 0.02
    c.bcT[names.index("zMin")] = cases.FG          # fixedGradient T
    c.bvT[pid == names.index("zMin")] = -0.3
    return c


QHD_CASES["cavity3d_fixed_p_patch"] = _qhd_fixed_p_case


@pytest.mark.parametrize("name", list(QHD_CASES))
def test_qhdfoam_100_steps_match_oracle(qgd, oracle_mod, name):
    c = QHD_CASES[name]()
    o = c.make_oracle(oracle_mod)
    s = c.make_solver(qgd)
    assert rel_linf(s.get("tauQGD"), o.qhd_get("tauQGD")) < 1e-14
    c.oracle_step(o, 100)
    s.step(100)
    kind = c.mesh.patch_kind_per_bface()
    for f in ("U", "T", "p"):
        gc, gb = s.get(f, with_bnd=True)
        oc, ob = o.qhd_get(f, with_bnd=True)
        assert np.isfinite(gc).all()
        scale = float(np.abs(oc).max())
        assert float(np.abs(gc - oc).max()) / scale < TOL_STEP, f"{name}: cells {f}"
        assert float(np.abs(gb[kind != 1] - ob[kind != 1]).max()) / scale < TOL_STEP, f"{name}: boundary {f}"
    assert rel_linf(s.get_flux(), o.qhd_get_face("phi")) < 1e-9
    gi, oi = s.solver_info(), o.qhd_solver_info()
    assert abs(gi["iters"] - oi["iters"]) <= max(2, oi["iters"] // 10) and gi["final_residual"] < 1e-12
    if c.opts["adjust_time_step"]:
        assert abs(s.scalars()["deltaT"] - o.qhd_deltaT()) < 1e-10 * o.qhd_deltaT()


def test_qhd_hydrostatic_rest_state_on_device(qgd):
    """Size-independent property (no oracle): hydrostatic pressure + consistent fixedGradient p keeps the fluid at rest."""
    c = cases.qhd_cavity(n=(64, 64), dt=1e-3, precond="diagonal", tol=1e-14, max_iter=20000)
    m = c.mesh
    c.U0[:] = 0.0
    c.T0[:] = 0.7
    c.bcT[:] = cases.ZG
    f = c.fluid
    bd = f["beta"] * 0.7 * np.asarray(f["g"])
    nI = m.n_internal
    c.bvP = f["rho0"] * ((m.Sf[nI:] / m.magSf[nI:, None]) @ bd)
    c.p0 = f["rho0"] * (m.C @ bd)
    s = c.make_solver(qgd)
    s.step(5)
    assert np.abs(s.get("U")).max() < 1e-10
    assert np.abs(s.get("T") - 0.7).max() < 1e-12
    assert np.abs(s.get_flux()).max() < 1e-12


def test_qhd_projection_is_divergence_free_at_scale(qgd):
    """Size-independent property at ~0.26 M cells: after pEqn every cell but the reference cell balances its fluxes."""
    c = cases.qhd_cavity(n=(512, 512), dt=2e-4, precond="diagonal", tol=1e-11, max_iter=50000)
    m = c.mesh
    s = c.make_solver(qgd)
    s.step(2)
    phi = s.get_flux()
    nI = m.n_internal
    d = np.zeros(m.n_cells)
    np.add.at(d, m.owner[:nI], phi[:nI])
    np.add.at(d, m.neighbour, -phi[:nI])
    np.add.at(d, m.owner[nI:], phi[nI:])
    d[c.p_ref_cell] = 0.0
    assert np.abs(d).max() < 1e-6 * np.abs(phi).max()
    assert np.isfinite(s.get("U")).all() and np.isfinite(s.get("T")).all()
    info = s.solver_info()
    assert 0 < info["iters"] < 50000 and info["final_residual"] < 1e-11


def test_qhd_error_behaviour(qgd):
    mesh = cases.pm.hex_box(4, 4, 1, patch_kinds={"zMin": "empty", "zMax": "empty"})
    dm = qgd.Mesh(mesh)
    kw = dict(rho0=1.0, mu=1e-2, Pr=0.7, beta=1e-3, g=(0, -9.81, 0))
    with pytest.raises(qgd.QGDError) as e:
        qgd.QHDFoam(dm, qgd_coeffs="noModel", **kw)
    assert e.value.code == qgd.ERR_UNKNOWN_MODEL and "Unknown QGD coeffs evaluation approach type noModel" in e.value.message
    with pytest.raises(qgd.QGDError) as e:
        qgd.QHDFoam(dm, qgd_coeffs="constScPrModel1", **kw)
    assert e.value.code == qgd.ERR_UNSUPPORTED
    with pytest.raises(qgd.QGDError) as e:
        qgd.QHDFoam(dm, precond="GAMG", **kw)
    assert "Unknown symmetric matrix preconditioner GAMG" in e.value.message
    qgd.QHDFoam(dm, fvsc_scheme="leastSquares", **kw)                  # 2D: available since round 2
    with pytest.raises(qgd.QGDError) as e:                              # fvsc.C:60-63
        qgd.QHDFoam(qgd.Mesh(cases.pm.hex_box(3, 3, 3)), fvsc_scheme="leastSquares", **kw)
    assert "Can't use leastSquares or leastSquaresOpt in 3D case." in e.value.message
    with pytest.raises(qgd.QGDError) as e:
        qgd.QHDFoam(dm, **kw).step(1)
    assert e.value.code == qgd.ERR_STATE


SCALAR_CASES = {
    # scalarTransportQHDFoam.C:70-135: frozen U (far from solenoidal), T equation with -fvc::Sp(fvc::div(phiu),T), implicit only
    "st2d_implicit": lambda: cases.qhd_cavity(n=(18, 16), dt=1e-3, perturb=0.1, implicit=True, scalar_transport=True),
    "st3d_implicit_diag_H2bynu": lambda: cases.qhd_cavity(n=(8, 7, 6), dims=3, dt=1e-3, implicit=True, scalar_transport=True,
                                                            model="H2bynuQHD", diff_solver=dict(precond="diagonal")),
    "st2d_implicit_adjust": lambda: cases.qhd_cavity(n=(16, 14), dt=1e-3, implicit=True, scalar_transport=True,
                                                      adjust_time_step=True, max_co=0.05, c_tau=0.4),
    "st3d_reduced": lambda: cases.qhd_cavity(n=(8, 8, 6), dims=3, dt=1e-3, scheme="reduced", implicit=True, scalar_transport=True),
    "st2d_explicit_noop": lambda: cases.qhd_cavity(n=(12, 10), dt=1e-3, implicit=False, scalar_transport=True),
}


@pytest.mark.parametrize("name", list(SCALAR_CASES))
def test_scalar_transport_qhdfoam_matches_oracle(qgd, oracle_mod, name):
    c = SCALAR_CASES[name]()
    m = c.mesh
    if c.implicit:      # a velocity field with a strong divergence, so the Sp term matters
        c.U0 = np.stack([0.3 * np.sin(3 * m.C[:, 0]) + 0.1, 0.2 * np.cos(2 * m.C[:, 1]) * m.C[:, 0], 0.1 * m.C[:, 2] * (m.geometric_d[2] > 0)], 1)
        c.bcU[:] = cases.ZG
    o = c.make_oracle(oracle_mod)
    s = c.make_solver(qgd)
    U0, p0, T0 = s.get("U").copy(), s.get("p").copy(), s.get("T").copy()
    c.oracle_step(o, 60)
    s.step(60)
    kind = m.patch_kind_per_bface()
    gc, gb = s.get("T", with_bnd=True)
    oc, ob = o.qhd_get("T", with_bnd=True)
    scale = float(np.abs(oc).max())
    assert float(np.abs(gc - oc).max()) / scale < TOL_STEP and float(np.abs(gb[kind != 1] - ob[kind != 1]).max()) / scale < TOL_STEP
    assert np.array_equal(s.get("U"), U0) and np.array_equal(s.get("p"), p0)          # never touched
    assert rel_linf(s.get_flux(), o.qhd_get_face("phi")) < 1e-12                      # phi == phiu
    if c.implicit:
        assert float(np.abs(gc - T0).max()) > 1e-6                                    # T did move
    else:
        assert np.array_equal(gc, T0)                                                 # :114 nothing is solved
    if c.opts["adjust_time_step"]:
        assert abs(s.scalars()["deltaT"] - o.qhd_deltaT()) < 1e-10 * o.qhd_deltaT()
