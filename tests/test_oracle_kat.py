"""Known-answer tests that pin the CPU oracle (the reference ships no tests or golden vectors for this path):
manufactured checks derived from the reference source itself plus committed golden vectors of the oracle."""
import os

import numpy as np
import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _interior_faces(mesh):
    """internal faces none of whose vertices lies on a non-empty boundary face"""
    nI = mesh.n_internal
    kind = mesh.patch_kind_per_bface()
    isb = np.zeros(mesh.n_points, bool)
    nv = mesh.face_nverts()
    bf = np.nonzero(kind != 1)[0] + nI
    for f in bf:
        isb[mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]] = True
    onb = np.add.reduceat(isb[mesh.face_verts].astype(int), mesh.face_offsets[:-1])[:nI]
    return onb == 0


def _linear(mesh, a, b=0.5):
    nI = mesh.n_internal
    phi = mesh.C @ a + b
    bnd = mesh.Cf[nI:] @ a + b
    bsg = mesh.deltaCoeffs[nI:] * (bnd - phi[mesh.owner[nI:]])
    return phi, bnd, bsg


@pytest.mark.parametrize("mesh_fn", [lambda: cases.pm.hex_box(7, 6, 5), lambda: cases.pm.prism_box(5, 5, 4),
                                     lambda: cases.case_2d((12, 10)).mesh, lambda: cases.case_sod(30).mesh])
def test_linear_exactness_on_uniform_meshes(oracle_mod, mesh_fn):
    """phi = a.x + b: GaussVolPoint returns a on interior faces (the commented check at QHDFoam/createFaceFluxes.H:46-63)"""
    mesh = mesh_fn()
    a = np.array([0.3, -1.2, 0.7]) * (mesh.geometric_d > 0)
    o = oracle_mod.Oracle(mesh)
    g = o.fvsc_grad(*_linear(mesh, a))
    inner = _interior_faces(mesh)
    assert inner.sum() > 0
    assert np.abs(g[:mesh.n_internal][inner] - a).max() < 1e-13


def test_gauss_formula_matches_closed_form_on_perturbed_hex(oracle_mod):
    """SURVEY A.2 closed form (derived from GaussVolPointBase3D.C:346-389), evaluated independently in numpy with the
    oracle's own point values, must equal the oracle's coefficient-table evaluation."""
    mesh = cases.pm.hex_box(6, 5, 4, perturb=0.25, grading=(2, 1, 0.5), seed=9)
    rng = np.random.default_rng(1)
    phi = rng.random(mesh.n_cells)
    nI = mesh.n_internal
    bnd = rng.random(mesh.n_bnd)
    bsg = mesh.deltaCoeffs[nI:] * (bnd - phi[mesh.owner[nI:]])
    o = oracle_mod.Oracle(mesh)
    g = o.fvsc_grad(phi, bnd, bsg)
    pf = o.vol_point_interpolate(phi, bnd)
    fv = mesh.face_verts.reshape(-1, 4)[:nI]
    p = mesh.points[fv]
    d = mesh.C[mesh.neighbour] - mesh.C[mesh.owner[:nI]]
    e1, e2 = p[:, 1] - p[:, 3], p[:, 2] - p[:, 0]
    D = (e2 * np.cross(e1, d)).sum(1)
    ref = (np.cross(d, e1) * (pf[fv[:, 0]] - pf[fv[:, 2]])[:, None] + np.cross(d, e2) * (pf[fv[:, 1]] - pf[fv[:, 3]])[:, None]
           + np.cross(e1, e2) * (phi[mesh.owner[:nI]] - phi[mesh.neighbour])[:, None]) / D[:, None]
    assert np.abs(g[:nI] - ref).max() / np.abs(ref).max() < 1e-12


def test_reduced_equals_nf_sngrad_and_gaussvolpoint_normal_part(oracle_mod):
    mesh = cases.pm.hex_box(6, 5, 4)
    a = np.array([0.0, 0.0, 1.3])
    phi, bnd, bsg = _linear(mesh, a)
    o = oracle_mod.Oracle(mesh)
    gr = o.fvsc_grad(phi, bnd, bsg, scheme=oracle_mod.FVSC_REDUCED)
    nI = mesh.n_internal
    nf = mesh.Sf / mesh.magSf[:, None]
    sn = mesh.nonOrthDeltaCoeffs[:nI] * (phi[mesh.neighbour] - phi[mesh.owner[:nI]])
    assert np.abs(gr[:nI] - nf[:nI] * sn[:, None]).max() < 1e-13
    # a field varying only along z has zero tangential differences on z-faces: both schemes agree there
    gg = o.fvsc_grad(phi, bnd, bsg)
    zf = np.abs(nf[:nI, 2]) > 0.99
    assert np.abs(gg[:nI][zf] - gr[:nI][zf]).max() < 1e-12


def test_div_is_trace_of_grad(oracle_mod):
    mesh = cases.pm.hex_box(6, 5, 4, perturb=0.2, seed=3)
    rng = np.random.default_rng(5)
    U = rng.random((mesh.n_cells, 3))
    nI = mesh.n_internal
    Ub = rng.random((mesh.n_bnd, 3))
    bsg = mesh.deltaCoeffs[nI:, None] * (Ub - U[mesh.owner[nI:]])
    o = oracle_mod.Oracle(mesh)
    G = o.fvsc_grad(U, Ub, bsg)
    dv = o.fvsc_div(U, Ub, bsg)
    assert np.abs(dv - (G[:, 0] + G[:, 4] + G[:, 8])).max() < 1e-12


def test_tri_face_vector_gradient_quirk(oracle_mod):
    """internal triangular faces: off-diagonals follow the reference's index pattern, the trace is right
    (GaussVolPointBase3D.C:844-854)"""
    mesh = cases.pm.prism_box(4, 4, 3, perturb=0.1)
    rng = np.random.default_rng(2)
    U = rng.random((mesh.n_cells, 3)); Ub = rng.random((mesh.n_bnd, 3))
    nI = mesh.n_internal
    bsg = mesh.deltaCoeffs[nI:, None] * (Ub - U[mesh.owner[nI:]])
    o = oracle_mod.Oracle(mesh)
    G = o.fvsc_grad(U, Ub, bsg)
    tri = mesh.face_nverts()[:nI] == 3
    assert tri.any()
    Gt = G[:nI][tri]
    assert np.abs(Gt[:, 0:3] - Gt[:, 3:6]).max() == 0 and np.abs(Gt[:, 0:3] - Gt[:, 6:9]).max() == 0
    comp = [o.fvsc_grad(U[:, j].copy(), Ub[:, j].copy(), bsg[:, j].copy())[:nI][tri][:, j] for j in range(3)]
    assert np.abs(Gt[:, 0] - comp[0]).max() < 1e-13 and np.abs(Gt[:, 4] - comp[1]).max() < 1e-13


def test_qgd_length_scales(oracle_mod):
    mesh = cases.pm.hex_box(8, 4, 2, lengths=(1.0, 1.0, 1.0))
    o = oracle_mod.Oracle(mesh)
    hf, h = o.hQGDf(), o.hQGD()
    nI = mesh.n_internal
    d = np.abs(mesh.C[mesh.neighbour] - mesh.C[mesh.owner[:nI]]).sum(1)
    assert np.abs(hf[:nI] - d).max() < 1e-14                  # 2*min(|C-Cf|) = cell spacing on a uniform mesh
    assert np.abs(hf[nI:] - 2.0 / mesh.deltaCoeffs[nI:]).max() < 1e-14
    # area-weighted mean of the face lengths: 2*(dx*Ax + dy*Ay + dz*Az)/(2*(Ax+Ay+Az))
    dx, dy, dz = 1 / 8, 1 / 4, 1 / 2
    exp = (dx * dy * dz * 3) / (dy * dz + dx * dz + dx * dy)
    assert np.abs(h - exp).max() < 1e-14


def test_sod_tube_against_exact_riemann_solution(oracle_mod):
    """physics sanity with the energy quirk off (the literal QGDEEqn.H:67-72 form does not conserve total energy)"""
    c = cases.case_sod(400, dt=2e-4, energy_ddt_rhoE_quirk=False)
    o = c.make_oracle(oracle_mod, n_threads=2)
    c.oracle_step(o, 1000)
    rho, U, p = o.get("rho"), o.get("U"), o.get("p")
    i, j = int(0.6 * 400), int(0.8 * 400)
    assert abs(rho[i] - 0.4263) < 0.01 and abs(rho[j] - 0.2656) < 0.005
    assert abs(U[i, 0] - 0.9275) < 0.01 and abs(p[i] - 0.3031) < 0.005
    # with the quirk the plateau is measurably different: the flag is live
    c2 = cases.case_sod(400, dt=2e-4)
    o2 = c2.make_oracle(oracle_mod, n_threads=2)
    c2.oracle_step(o2, 1000)
    assert abs(o2.get("p")[i] - 0.3031) > 0.02


def test_mass_conservation_with_qgdflux_walls(oracle_mod):
    """sum rho V changes only through boundary phiJm (telescoping fvc::div); with U=0 walls and the qgdFlux pressure BC
    the regularised wall mass flux vanishes identically (qgdFluxFvPatchScalarField.C:184-192)"""
    c = cases.case_poly(bcs="qgdflux")
    o = c.make_oracle(oracle_mod)
    m0 = (o.get("rho") * c.mesh.V).sum()
    c.oracle_step(o, 20)
    assert abs((o.get("rho") * c.mesh.V).sum() - m0) < 1e-13 * m0
    assert np.abs(o.get_face("phiJm")[c.mesh.n_internal:]).max() < 1e-12


def test_pcg_dic_against_direct_solve(oracle_mod):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    mesh = cases.pm.hex_box(8, 7, 6, perturb=0.1, seed=1)
    nI = mesh.n_internal
    upper = -(mesh.magSf[:nI] * mesh.deltaCoeffs[:nI])
    diag = np.zeros(mesh.n_cells)
    np.subtract.at(diag, mesh.owner[:nI], upper); np.subtract.at(diag, mesh.neighbour, upper)
    diag += 1e-3 * mesh.V / mesh.V.mean()
    b = np.random.default_rng(0).standard_normal(mesh.n_cells)
    o = oracle_mod.Oracle(mesh)
    A = sp.coo_matrix((np.concatenate([diag, upper, upper]),
                       (np.concatenate([np.arange(mesh.n_cells), mesh.owner[:nI], mesh.neighbour]),
                        np.concatenate([np.arange(mesh.n_cells), mesh.neighbour, mesh.owner[:nI]])))).tocsr()
    xref = spl.spsolve(A.tocsc(), b)
    its = {}
    for pc in (0, 1, 2):
        x, it, r0, r1 = o.pcg_solve(diag, upper, b, np.zeros(mesh.n_cells), tol=1e-13, maxIter=2000, precond=pc)
        assert np.abs(x - xref).max() / np.abs(xref).max() < 1e-9
        its[pc] = it
    assert its[2] < its[1] <= its[0]


def test_block_local_dic_pcg_against_an_independent_restatement(oracle_mod):
    """or_pcg_solve_blocks with DIC = the solver of a decomposed reference run: global matrix and reductions, the preconditioner
    M = (D + L_b) D^-1 (D + L_b)^T built per block from the in-block faces only [OF-v2312 DICPreconditioner on each processor's
    lduMatrix].  Independent restatement: the same M assembled as a scipy sparse matrix (D from the DIC recurrence on the block
    graph) and applied by two sparse triangular solves inside a textbook preconditioned CG with OpenFOAM's L1 normFactor residual;
    iterates after a fixed number of iterations and the converged solves must agree.  With one block it is the serial DIC."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    mesh = cases.pm.hex_box(9, 8, 5, perturb=0.15, seed=6)
    n, nI = mesh.n_cells, mesh.n_internal
    l, u = mesh.owner[:nI].astype(int), mesh.neighbour.astype(int)
    rng = np.random.default_rng(3)
    upper = -(0.5 + rng.random(nI)) * mesh.magSf[:nI] * mesh.deltaCoeffs[:nI]
    diag = np.zeros(n)
    np.subtract.at(diag, l, upper); np.subtract.at(diag, u, upper)
    diag += 1e-2 * np.abs(diag).mean() * rng.random(n)
    b = rng.standard_normal(n)
    A = sp.coo_matrix((np.concatenate([diag, upper, upper]), (np.concatenate([np.arange(n), l, u]), np.concatenate([np.arange(n), u, l])))).tocsr()
    o = oracle_mod.Oracle(mesh)
    ix = np.minimum((mesh.C[:, 0] * 3).astype(int), 2)
    iy = np.minimum((mesh.C[:, 1] * 2).astype(int), 1)
    for blocks in (np.zeros(n, np.int32), (ix * 2 + iy).astype(np.int32), (np.arange(n) // 37).astype(np.int32)):
        inb = blocks[l] == blocks[u]
        # DIC recurrence on the block graph: rD = diag; faces ascending: rD[u] -= upper^2 / rD[l]
        rD = diag.copy()
        for f in np.nonzero(inb)[0]:
            rD[u[f]] -= upper[f] ** 2 / rD[l[f]]
        Lb = sp.coo_matrix((upper[inb], (u[inb], l[inb])), shape=(n, n)).tocsr()       # strictly lower part, in-block entries
        DL = (sp.diags(rD) + Lb).tocsr()

        def Minv(r):      # M = (D + L) D^-1 (D + L)^T
            y = spl.spsolve_triangular(DL, r, lower=True)
            return spl.spsolve_triangular(DL.T.tocsr(), rD * y, lower=False)

        def pcg(x, iters, tol):
            x = x.copy()
            r = b - A @ x
            xref = x.mean()
            sumA = np.asarray(A.sum(1)).ravel()
            nf = np.abs(A @ x - xref * sumA).sum() + np.abs(b - xref * sumA).sum() + 1e-20
            p = np.zeros(n); rho_old = 1.0
            for k in range(iters):
                if np.abs(r).sum() / nf < tol:
                    return x, k
                z = Minv(r)
                rho = z @ r
                p = z + (rho / rho_old) * p if k else z.copy()
                w = A @ p
                alpha = rho / (w @ p)
                x += alpha * p; r -= alpha * w
                rho_old = rho
            return x, iters

        x0 = 0.1 * np.cos(3 * mesh.C[:, 0])
        for k in (1, 4, 9):
            xo, ito, _, _ = o.pcg_solve(diag, upper, b, x0, tol=0.0, maxIter=k, precond=2, cell_block=blocks)
            xr, _ = pcg(x0, k, 0.0)
            assert ito == k and np.abs(xo - xr).max() < 1e-11 * np.abs(xr).max()
        xo, ito, _, _ = o.pcg_solve(diag, upper, b, x0, tol=1e-11, maxIter=2000, precond=2, cell_block=blocks)
        xr, itr = pcg(x0, 2000, 1e-11)
        assert abs(ito - itr) <= 1 and np.abs(xo - xr).max() < 1e-8 * np.abs(xr).max()
    # one block == the serial solver
    xa, ita, _, _ = o.pcg_solve(diag, upper, b, x0, tol=1e-11, maxIter=2000, precond=2)
    xb, itb, _, _ = o.pcg_solve(diag, upper, b, x0, tol=1e-11, maxIter=2000, precond=2, cell_block=np.zeros(n, np.int32))
    assert ita == itb and np.array_equal(xa, xb)


GOLDEN_CASES = {"hex_perturbed_mixed": lambda: cases.case_hex3d(perturb=0.2, grading=(2, 1, 0.5), bcs="mixed"),
                "2d_mixed": lambda: cases.case_2d(perturb=0.2, bcs="mixed"),
                "sod_1d": lambda: cases.case_sod(100),
                "hex_implicit": lambda: cases.case_hex3d(perturb=0.15, bcs="mixed", implicit=True),
                "prism_model1n": lambda: cases.case_prism(bcs="fixed", model="constScPrModel1n"),
                "qhd_cavity2d": lambda: cases.qhd_cavity(n=(16, 14), dt=1e-3, perturb=0.1),
                "qhd_cavity3d_H2bynu": lambda: cases.qhd_cavity(n=(8, 7, 6), dims=3, dt=5e-4, model="H2bynuQHD", precond="diagonal"),
                "hex_varSc7": lambda: cases.case_hex3d(perturb=0.1, bcs="fixed", model="varScModel7",
                                                       varsc=dict(cSc1=3.0, minSc=0.02, maxSc=0.4, const_sc_cells=np.arange(5, 300, 7))),
                "poly_varSc6": lambda: cases.case_poly(bcs="qgdflux", model="varScModel6"),
                "hex_sources": lambda: cases.with_sources(cases.case_hex3d(perturb=0.2, bcs="mixed")),
                "2d_sources_implicit": lambda: cases.with_sources(cases.case_2d(perturb=0.1, bcs="fixed", implicit=True)),
                "qhd_scalar_transport2d": lambda: cases.scalar_transport_case(n=(16, 14), dt=1e-3, perturb=0.1, implicit=True),
                "qhd_scalar_transport3d_adjust": lambda: cases.scalar_transport_case(n=(8, 7, 6), dims=3, dt=1e-3, implicit=True,
                                                                                    adjust_time_step=True, max_co=0.05, c_tau=0.4)}


def golden_fields(c, o):
    """the fields a golden fixture holds, from an oracle (or, with the same accessor names, a CUDA solver)"""
    if isinstance(c, cases.QHDCase):
        get = o.qhd_get if hasattr(o, "qhd_get") else o.get
        return {f: get(f) for f in ("U", "T", "p")}
    return {f: o.get(f) for f in ("rho", "rhoU", "rhoE", "e", "p")}


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_oracle_reproduces_committed_golden_vectors(oracle_mod, name):
    """golden vectors generated by tests/golden/make_golden.py (oracle outputs; they pin the oracle against regressions,
    they are NOT reference outputs - parity with the reference itself is unpinned)"""
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    c = GOLDEN_CASES[name]()
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, int(z["steps"]))
    for f, v in golden_fields(c, o).items():
        ref = z[f]
        assert np.abs(v - ref).max() <= 1e-12 * np.abs(ref).max(), f


# ============================================================================ QHDFoam oracle KATs
def _div(mesh, phi):
    nI = mesh.n_internal
    d = np.zeros(mesh.n_cells)
    np.add.at(d, mesh.owner[:nI], phi[:nI])
    np.add.at(d, mesh.neighbour, -phi[:nI])
    np.add.at(d, mesh.owner[nI:], phi[nI:])
    return d


def test_qhd_hydrostatic_balance_keeps_fluid_at_rest(oracle_mod):
    """U = 0, uniform T, uniform body force, p BC = fixedGradient rho*BdFrc.n: the discrete solution is the linear
    hydrostatic pressure, every face flux phi vanishes and U, T do not move (QHDpEqn.H:36-47, QHDUEqn.H:36-84)."""
    import cases
    c = cases.qhd_cavity(n=(10, 12), dt=1e-3)
    m = c.mesh
    c.U0[:] = 0.0
    c.T0[:] = 0.7
    c.bcT[:] = cases.ZG
    f = c.fluid
    bd = f["beta"] * 0.7 * np.asarray(f["g"])
    nI = m.n_internal
    nf = m.Sf[nI:] / m.magSf[nI:, None]
    c.bvP = f["rho0"] * (nf @ bd)
    c.p0 = f["rho0"] * (m.C @ bd)
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 5)
    assert np.abs(o.qhd_get("U")).max() < 1e-11
    assert np.abs(o.qhd_get("T") - 0.7).max() < 1e-12
    assert np.abs(o.qhd_get_face("phi")).max() < 1e-13
    p = o.qhd_get("p")
    pl = f["rho0"] * (m.C @ bd)
    assert np.abs((p - p[0]) - (pl - pl[0])).max() < 1e-10


def test_qhd_temperature_diffusion_matches_discrete_eigenmode(oracle_mod):
    """U = 0, g = 0: T_t = div(Hi grad T) with zeroGradient walls; cos(pi x) is an exact eigenvector of the cell-centred
    discrete Laplacian, so n Euler steps multiply it by (1 - dt*lambda_h)^n (QHDTEqn.H:83-91)."""
    import cases
    n = 16
    c = cases.qhd_cavity(n=(n, 4), dt=2e-3, fluid=dict(cases.QHD_FLUID, g=(0.0, 0.0, 0.0)))
    m = c.mesh
    c.U0[:] = 0.0
    c.bcT[:] = cases.ZG
    c.T0 = 1.0 + 0.3 * np.cos(np.pi * m.C[:, 0])
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 20)
    h = 1.0 / n
    Hi = (c.fluid["mu"] / c.fluid["Pr"]) / c.fluid["rho0"]
    lam = Hi * (2.0 - 2.0 * np.cos(np.pi * h)) / h ** 2
    ref = 1.0 + 0.3 * (1.0 - c.dt * lam) ** 20 * np.cos(np.pi * m.C[:, 0])
    assert np.abs(o.qhd_get("T") - ref).max() < 1e-13
    assert np.abs(o.qhd_get("U")).max() < 1e-14


@pytest.mark.parametrize("model", ["constTau", "H2bynuQHD", "HbyUQHD", "T0byGr"])
def test_qhd_projection_is_divergence_free(oracle_mod, model):
    """After pEqn the flux phi = phiu - phiwo + pEqn.flux() balances in every cell except the reference cell, whose
    equation fvMatrix::setReference relaxes (QHDpEqn.H:43-47); tauQGD follows the selected model."""
    import cases
    c = cases.qhd_cavity(n=(12, 10), dt=1e-3, model=model, perturb=0.15, coeffs=dict(Tau=2e-3, UQHD=5.0, Gr=500.0, T0=1.0))
    c.bcT[:] = cases.ZG
    m = c.mesh
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 3)
    d = _div(m, o.qhd_get_face("phi"))
    scale = np.abs(o.qhd_get_face("phi")).max()
    d[c.p_ref_cell] = 0.0
    assert np.abs(d).max() < 1e-9 * scale
    info = o.qhd_solver_info()
    assert 0 < info["iters"] < 400 and info["final_residual"] < 1e-12
    tau = o.qhd_get("tauQGD")
    h = o.hQGD()
    f, k = c.fluid, c.coeffs
    expect = {"constTau": np.full_like(h, k["Tau"]), "H2bynuQHD": 0.5 * h * h / (f["mu"] / f["rho0"]),
              "HbyUQHD": 0.5 * h / k["UQHD"], "T0byGr": np.full_like(h, k["T0"] / k["Gr"])}[model]
    assert np.allclose(tau, expect, rtol=1e-14, atol=0)


def test_qhd_preconditioners_agree(oracle_mod):
    """DIC / diagonal / none change the PCG iteration count, not the converged step."""
    import cases
    res = {}
    for pc in ("DIC", "diagonal", "none"):
        c = cases.qhd_cavity(n=(10, 9, 8), dims=3, dt=1e-3, precond=pc, perturb=0.1)
        o = c.make_oracle(oracle_mod)
        c.oracle_step(o, 3)
        res[pc] = (o.qhd_get("U"), o.qhd_get("T"), o.qhd_get("p"), o.qhd_solver_info()["iters"])
    for pc in ("diagonal", "none"):
        for a, b in zip(res[pc][:3], res["DIC"][:3]):
            assert np.abs(a - b).max() < 1e-9 * max(np.abs(b).max(), 1e-30)
    assert res["DIC"][3] < res["diagonal"][3] <= res["none"][3]


@pytest.mark.parametrize("implicit", [False, True])
def test_shear_wave_decay_matches_discrete_eigenvalue(oracle_mod, implicit):
    """U = (0, A sin(pi x), 0) between no-slip x-walls in a uniform gas: every QGD regularisation term vanishes (no
    pressure / density gradient, div U = 0, U.grad U = 0), leaving d(rho Uy)/dt = d/dx(mu dUy/dx) with mu = mu_mol + muQGD.
    sin(pi x_c) is an exact eigenvector of the cell-centred Dirichlet Laplacian, so one step multiplies the amplitude by
    (1 - dt nu lambda_h) in the explicit branch (Pi = mu(...) through the fvsc face gradient, QGDFoam/updateFluxes.H:95-106)
    and by 1/(1 + dt nu lambda_h) in the implicit branch (fvm::laplacian(muf,U), QGDUEqn.H:56-62).  (Only the first step
    is exact: the patch-point values on the wall edges then feed a small div U into the wall faces.)"""
    import cases
    n, A, N, dt = 16, 1e-3, 1, 1e-3
    mesh = cases.pm.hex_box(n, 2, 2, lengths=(1.0, 2.0 / n, 2.0 / n))      # cubic cells: hQGD == hQGDf everywhere
    nP, nB = len(mesh.patches), mesh.n_bnd
    names = [p.name for p in mesh.patches]
    kU = np.array([cases.FV if nm in ("xMin", "xMax") else cases.ZG for nm in names], np.int32)
    kT = np.full(nP, cases.ZG, np.int32)
    gas = dict(cases.GAS, mu=0.05)
    g = gas["Cp"] / (gas["Cp"] - gas["R"])
    x = mesh.C[:, 0]
    U0 = np.zeros((mesh.n_cells, 3))
    U0[:, 1] = A * np.sin(np.pi * x)
    p0 = np.full(mesh.n_cells, 1.0 / g)
    T0 = p0 / gas["R"]
    c = cases.Case(mesh, U0, T0, p0, kU, kT, kT, np.zeros((nB, 3)), np.ones(nB), np.ones(nB), gas=gas, dt=dt, implicit=implicit)
    o = c.make_oracle(oracle_mod)
    mu = o.get("mu")
    assert np.ptp(mu) < 1e-14 * mu.mean()
    nu = mu.mean() / 1.0
    c.oracle_step(o, N)
    h = 1.0 / n
    lam = (2.0 - 2.0 * np.cos(np.pi * h)) / h ** 2
    fac = (1.0 + dt * nu * lam) ** (-N) if implicit else (1.0 - dt * nu * lam) ** N
    rhoU = o.get("rhoU")
    assert np.abs(rhoU[:, 1] - fac * A * np.sin(np.pi * x)).max() < 1e-11 * A
    assert abs(fac - 1.0) > 5e-4


# ---------------------------------------------------------------- varScModel6 / varScModel7 (pressure-jump sensor)
def _varsc_numpy(mesh, p, pB, bcP, cSc1=1.0):
    """Independent vectorised restatement of varScModel6.C:210-269 / varScModel7.C:176-235 (zeroGradient / fixedValue p)."""
    nI, nC = mesh.n_internal, mesh.n_cells
    own, nei = mesh.owner, mesh.neighbour
    w = mesh.weights[:nI]
    pf = w * p[own[:nI]] + (1 - w) * p[nei]
    dpf = mesh.nonOrthDeltaCoeffs[:nI] * (p[nei] - p[own[:nI]]) / mesh.deltaCoeffs[:nI]
    kind = mesh.patch_kind_per_bface()
    act = kind != 1                                            # not empty
    fixed = bcP[mesh.patch_id_per_bface()] == cases.FV
    dpb = np.where(fixed, pB - p[own[nI:]], 0.0)
    sumD, sumP, n = np.zeros(nC), np.zeros(nC), np.zeros(nC)
    np.add.at(sumD, own[:nI], dpf); np.add.at(sumD, nei, -dpf); np.add.at(sumD, own[nI:][act], dpb[act])
    np.add.at(sumP, own[:nI], pf); np.add.at(sumP, nei, pf); np.add.at(sumP, own[nI:][act], pB[act])
    np.add.at(n, own[:nI], 1); np.add.at(n, nei, 1); np.add.at(n, own[nI:][act], 1)
    return cSc1 * np.abs(sumD) / (sumP / n)


@pytest.mark.parametrize("mk", [lambda **k: cases.case_hex3d(perturb=0.2, bcs="fixed", **k), lambda **k: cases.case_poly(**k),
                                lambda **k: cases.case_2d((12, 10), perturb=0.1, **k), lambda **k: cases.case_sod(40, **k)])
def test_varsc_sensor_matches_independent_restatement(oracle_mod, mk):
    c6 = mk(model="varScModel6")
    o = c6.make_oracle(oracle_mod)
    p, pB = o.get("p", with_bnd=True)
    sc, scB = o.get("ScQGD", with_bnd=True)
    ref = _varsc_numpy(c6.mesh, p, pB, c6.bcP)
    assert np.abs(sc - ref).max() <= 1e-12 * max(ref.max(), 1e-300)
    act = c6.mesh.patch_kind_per_bface() != 1
    assert np.all(scB[act] == c6.gas["ScQGD"])               # calculated patches keep the dictionary value
    # muQGD = p*ScQGD*tauQGD (varScModel6.C:313-322): mu - mu_molecular
    mu, tau = o.get("mu"), o.get("tauQGD")
    assert np.abs((mu - c6.gas["mu"]) - p * sc * tau).max() < 1e-14 * max(np.abs(mu).max(), 1.0)
    # model 7: cSc1 multiplier, clamps (incl. the boundary field) and the constScCellSet
    cells = np.arange(0, c6.mesh.n_cells, 3, dtype=np.int32)
    lo, hi = 0.25 * ref.max(), 0.6 * ref.max()
    c7 = mk(model="varScModel7", varsc=dict(cSc1=2.5, minSc=lo, maxSc=hi, const_sc_cells=cells))
    o7 = c7.make_oracle(oracle_mod)
    sc7, sc7B = o7.get("ScQGD", with_bnd=True)
    exp = np.clip(2.5 * ref, lo, hi)
    exp[cells] = c7.gas["ScQGD"]
    assert np.abs(sc7 - exp).max() <= 1e-12 * max(exp.max(), 1e-300)
    assert np.all(sc7B[act] == min(max(c7.gas["ScQGD"], lo), hi))


def test_varsc_vanishes_for_linear_pressure_on_a_uniform_line(oracle_mod):
    """On a uniform 1D mesh the sensor is the second difference of p: zero for linear p, 2 h^2 / mean(p_f) for p = x^2."""
    c = cases.case_sod(50, model="varScModel6")
    x = c.mesh.C[:, 0]
    c.p0 = 1.0 + 0.5 * x
    c.T0 = c.p0.copy()
    o = c.make_oracle(oracle_mod)
    sc = o.get("ScQGD")
    assert np.abs(sc[1:-1]).max() < 1e-14
    c.p0 = 1.0 + x * x
    c.T0 = c.p0.copy()
    o = c.make_oracle(oracle_mod)
    sc, p = o.get("ScQGD"), o.get("p")
    h = 1.0 / 50
    exp = 2 * h * h / ((p[2:] + 2 * p[1:-1] + p[:-2]) / 4)
    assert np.abs(sc[1:-1] - exp).max() < 1e-12 * exp.max()


def test_varsc_steps_stay_conservative(oracle_mod):
    c = cases.case_sod(100, model="varScModel7", varsc=dict(cSc1=1.0, minSc=0.05, maxSc=1.0))
    o = c.make_oracle(oracle_mod)
    m0 = (o.get("rho") * c.mesh.V).sum()
    c.oracle_step(o, 100)
    assert abs((o.get("rho") * c.mesh.V).sum() - m0) < 1e-13 * m0
    sc = o.get("ScQGD")
    assert sc.min() >= 0.05 and sc.max() <= 1.0 and sc.max() > 0.06     # the sensor fires at the discontinuities


# ---------------------------------------------------------------- explicit source matrices rhoSu / rhoUSu / rhoESu
def _at_rest(n=(6, 5, 4), implicit=False):
    c = cases.case_hex3d(n=n, bcs="zg", implicit=implicit)
    c.U0 = np.zeros_like(c.U0)
    c.p0 = np.full_like(c.p0, 1.0 / 1.4)
    c.T0 = np.full_like(c.T0, (1.0 / 1.4) / c.gas["R"])
    return c


@pytest.mark.parametrize("implicit", [False, True])
def test_mass_source_adds_exactly_its_integral(oracle_mod, implicit):
    """fvm::ddt(rho) + fvc::div(phiJm) == rhoSu on a closed box: total mass grows by dt * sum(rhoSu) per step."""
    c = _at_rest(implicit=implicit)
    o = c.make_oracle(oracle_mod)
    V = c.mesh.V
    rng = np.random.default_rng(5)
    su = 1e-3 * V * rng.random(c.mesh.n_cells)
    o.qgd_set_sources(suRho=su)
    m0 = (o.get("rho") * V).sum()
    c.oracle_step(o, 20)
    assert abs((o.get("rho") * V).sum() - (m0 + 20 * c.dt * su.sum())) < 1e-13 * m0
    o.qgd_set_sources()                                    # cleared: mass is conserved again
    m1 = (o.get("rho") * V).sum()
    c.oracle_step(o, 5)
    assert abs((o.get("rho") * V).sum() - m1) < 1e-13 * m1


def test_momentum_and_energy_sources_first_step_from_rest(oracle_mod):
    """Explicit branch, uniform gas at rest, one step: U = dt*rhoUSu/(rho V) while rhoU keeps the value of the
    conservative update (QGDUEqn.H:79-89: the source corrects U only); e rises by dt*rhoESu/(rho V) exactly - the
    solve of QGDEEqn.H:65-73 replaces e = rhoE/rho - 0.5|U|^2 (:49) by (rho0 e0 + rhoE - rhoE0 + dt Su/V)/rho."""
    c = _at_rest()
    o = c.make_oracle(oracle_mod)
    V = c.mesh.V
    rho0, e0 = o.get("rho").copy(), o.get("e").copy()
    suU = np.zeros((c.mesh.n_cells, 3))
    suU[:, 0] = 2e-2 * V
    suE = 5e-2 * V
    o.qgd_set_sources(suU=suU, suE=suE)
    c.oracle_step(o, 1)
    U, rhoU, e = o.get("U"), o.get("rhoU"), o.get("e")
    assert np.abs(rhoU).max() < 1e-14
    assert np.abs(U[:, 0] - c.dt * 2e-2 / rho0).max() < 1e-14 and np.abs(U[:, 1:]).max() < 1e-16
    assert np.abs(e - (e0 + c.dt * 5e-2 / rho0)).max() < 1e-14 * e0.max()


# ---------------------------------------------------------------- scalarTransportQHDFoam (frozen U, T equation only)
def test_scalar_transport_keeps_uniform_T_for_any_velocity(oracle_mod):
    """fvc::div(phiu*Tf) - fvc::Sp(fvc::div(phiu),T) vanishes for uniform T whatever div(phiu) is, the regularisation and
    the Laplacian vanish with grad T: a uniform T stays uniform although U is far from solenoidal; U and p never change
    (scalarTransportQHDFoam.C:110-125)."""
    import cases
    c = cases.qhd_cavity(n=(10, 8), dt=1e-3, perturb=0.15, implicit=True, scalar_transport=True)
    c.bcT[:] = cases.ZG
    c.bcU[:] = cases.ZG
    c.T0 = np.full_like(c.T0, 1.25)
    c.U0 = np.stack([np.sin(3 * c.mesh.C[:, 0]), c.mesh.C[:, 1] ** 2, 0 * c.mesh.C[:, 0]], 1)
    o = c.make_oracle(oracle_mod)
    U0, p0 = o.qhd_get("U").copy(), o.qhd_get("p").copy()
    c.oracle_step(o, 10)
    assert np.abs(o.qhd_get("T") - 1.25).max() < 1e-12
    assert np.array_equal(o.qhd_get("U"), U0) and np.array_equal(o.qhd_get("p"), p0)
    assert np.array_equal(o.qhd_get_face("phi"), o.qhd_get_face("phiu"))


def test_scalar_transport_diffusion_eigenmode_and_explicit_noop(oracle_mod):
    """U = 0: backward Euler on the discrete Laplacian eigenvector cos(pi x): factor 1/(1 + dt*lambda_h) per step
    (fvm::laplacian(Hif,T), scalarTransportQHDFoam.C:116-124); with implicitDiffusion false nothing is solved (:114)."""
    import cases
    n = 16
    mk = lambda implicit: cases.qhd_cavity(n=(n, 4), dt=2e-3, implicit=implicit, scalar_transport=True)
    c = mk(True)
    m = c.mesh
    c.U0[:] = 0.0
    c.bcT[:] = cases.ZG
    c.T0 = 1.0 + 0.3 * np.cos(np.pi * m.C[:, 0])
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 20)
    h = 1.0 / n
    Hi = (c.fluid["mu"] / c.fluid["Pr"]) / c.fluid["rho0"]
    lam = Hi * (2.0 - 2.0 * np.cos(np.pi * h)) / h ** 2
    ref = 1.0 + 0.3 * (1.0 + c.dt * lam) ** -20 * np.cos(np.pi * m.C[:, 0])
    assert np.abs(o.qhd_get("T") - ref).max() < 1e-12
    e = mk(False)
    e.T0 = c.T0.copy()
    oe = e.make_oracle(oracle_mod)
    T0 = oe.qhd_get("T").copy()
    e.oracle_step(oe, 5)
    assert np.array_equal(oe.qhd_get("T"), T0)


# ---------------------------------------------------------------- flux algebra against an einsum restatement
@pytest.mark.parametrize("implicit", [False, True])
def test_qgd_flux_algebra_matches_an_independent_einsum_restatement(oracle_mod, implicit):
    """QGDFoam/updateFields.H:45-80 + updateFluxes.H:54-139 written once more, directly from the listing, with numpy
    einsum on whole face fields (OpenFOAM index conventions: grad(U)_ij = d_i U_j, (v & T)_j = v_i T_ij,
    (T & v)_i = T_ij v_j, (A & B)_ij = A_ik B_kj, a*b = outer product).  Inputs: the oracle's cell / boundary state
    before the step and its fvsc face gradients; outputs compared: every flux the equations consume."""
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.2, bcs="fixed", implicit=implicit, gas=dict(cases.GAS, mu=5e-3))
    m = c.mesh
    o = c.make_oracle(oracle_mod)
    st = {f: o.get(f, with_bnd=True) for f in ("rho", "U", "rhoU", "rhoE", "p", "c", "mu", "alpha")}
    tau = o.get_face("tauQGDf")                      # as left by the last thermo.correct(): the tau the step's fluxes use
    c.oracle_step(o, 1)
    I = lambda f: o.linear_interpolate(*st[f])
    rho, U, rhoU, p, cs = I("rho"), I("U"), I("rhoU"), I("p"), I("c")
    UrhoU = o.linear_interpolate(np.einsum("ci,cj->cij", st["U"][0], st["rhoU"][0]).reshape(-1, 9),
                                 np.einsum("ci,cj->cij", st["U"][1], st["rhoU"][1]).reshape(-1, 9)).reshape(-1, 3, 3)
    H = o.linear_interpolate((st["rhoE"][0] + st["p"][0]) / st["rho"][0], (st["rhoE"][1] + st["p"][1]) / st["rho"][1])
    g = c.gas["Cp"] / (c.gas["Cp"] - c.gas["R"])
    muf, alphaf = I("mu"), I("alpha") * (g if c.opts["alpha_eff_gamma_factor"] else 1.0)     # alphaEff for internal energy [OF heThermo]
    gU, ge, gR, gP = o.get_face("gradUf").reshape(-1, 3, 3), o.get_face("gradef"), o.get_face("gradRhof"), o.get_face("gradPf")
    Sf = m.Sf
    divU = np.einsum("fii->f", gU)
    rhoW = tau[:, None] * (U * np.einsum("fj,fj->f", gR, U)[:, None] + rhoU * divU[:, None] + np.einsum("fi,fij->fj", rhoU, gU))
    phiw = np.einsum("fi,fi->f", Sf, rhoW)
    rhoW = rhoW + tau[:, None] * gP
    jm = rhoU - rhoW
    phiJm = np.einsum("fi,fi->f", Sf, jm)
    eye = np.eye(3)[None]
    Pi = tau[:, None, None] * (np.einsum("fik,fkj->fij", UrhoU, gU) + np.einsum("fi,fj->fij", U, gP)) \
        + tau[:, None, None] * (eye * (np.einsum("fi,fi->f", U, gP) + g * p * divU)[:, None, None])
    q = -tau[:, None] * np.einsum("fij,fj->fi", UrhoU, ge - (p / rho / rho)[:, None] * gR)
    if not implicit:
        Pi = Pi + muf[:, None, None] * (gU + np.swapaxes(gU, 1, 2) - (2.0 / 3.0) * eye * divU[:, None, None])
        q = q - alphaf[:, None] * ge
    exp = {"phiwStar": phiw, "phiJm": phiJm, "phiJmU": phiJm[:, None] * U, "phiP": Sf * p[:, None],
           "phiPi": np.einsum("fi,fij->fj", Sf, Pi), "phiJmH": phiJm * H, "phiQ": np.einsum("fi,fi->f", Sf, q),
           "phiPiU": np.einsum("fi,fi->f", Sf, np.einsum("fij,fj->fi", Pi, U))}
    for name, ref in exp.items():
        got = o.get_face(name)
        assert np.abs(got - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-30), name
    assert np.abs(cs).min() > 0


def test_qhd_flux_algebra_matches_an_independent_einsum_restatement(oracle_mod):
    """QHDFoam/updateFluxes.H:33-38 and QHDpEqn.H:47 once more with einsum: phiu = Sf & Uf,
    phiwo = Sf & (tauQGDf*((Uf & gradUf) - BdFrcf)), phi = phiu - phiwo + pEqn.flux() with the uncorrected Laplacian
    flux -(tauQGDf/rhof)|Sf| nonOrthDeltaCoeffs (p_N - p_P); checked on the internal faces after one step."""
    import cases
    c = cases.qhd_cavity(n=(7, 6, 5), dims=3, dt=1e-3, perturb=0.15)
    m = c.mesh
    nI = m.n_internal
    o = c.make_oracle(oracle_mod)
    U0, UB0 = o.qhd_get("U", with_bnd=True)
    T0, TB0 = o.qhd_get("T", with_bnd=True)
    c.oracle_step(o, 1)
    f = c.fluid
    Uf = o.linear_interpolate(U0, UB0)
    Bf = o.linear_interpolate(f["beta"] * T0[:, None] * np.asarray(f["g"])[None], f["beta"] * TB0[:, None] * np.asarray(f["g"])[None])
    gU = o.qhd_get_face("gradUf").reshape(-1, 3, 3)
    tau = o.qhd_get_face("tauQGDf")
    phiu = np.einsum("fi,fi->f", m.Sf, Uf)
    phiwo = np.einsum("fi,fi->f", m.Sf, tau[:, None] * (np.einsum("fi,fij->fj", Uf, gU) - Bf))
    assert np.abs(o.qhd_get_face("phiu") - phiu)[:nI].max() < 1e-14 * np.abs(phiu).max()
    assert np.abs(o.qhd_get_face("phiwo") - phiwo)[:nI].max() < 1e-12 * np.abs(phiwo).max()
    p = o.qhd_get("p")
    shift = 0.0                                       # p was shifted to the reference value after the flux was formed: differences only
    flux = -(tau[:nI] / f["rho0"]) * m.magSf[:nI] * m.nonOrthDeltaCoeffs[:nI] * (p[m.neighbour] - p[m.owner[:nI]] + shift)
    phi = phiu[:nI] - phiwo[:nI] + flux
    assert np.abs(o.qhd_get_face("phi")[:nI] - phi).max() < 1e-11 * np.abs(phiu).max()


@pytest.mark.parametrize("mesh_fn", [lambda: cases.pm.hex_box(5, 4, 3, perturb=0.2, seed=4), lambda: cases.pm.hexprism_poly(4, 3, 3, a=0.1, lz=0.4),
                                     lambda: cases.case_2d((8, 7), perturb=0.15).mesh])
def test_vol_point_interpolation_matches_numpy_restatement(oracle_mod, mesh_fn):
    """[OF-v2312 volPointInterpolation] interior points: inverse-distance weights 1/|x_p - C_c| over the cells around the
    point, normalised; points on a (non-empty) boundary patch: inverse-distance over the boundary faces around the
    point, 1/|x_p - Cf_b|, using boundary values only (volPointInterpolation::makeWeights / makeBoundaryWeights)."""
    mesh = mesh_fn()
    o = oracle_mod.Oracle(mesh)
    nI = mesh.n_internal
    rng = np.random.default_rng(2)
    cell, bnd = rng.random(mesh.n_cells), rng.random(mesh.n_bnd) + 5.0
    got = o.vol_point_interpolate(cell, bnd)
    kind = mesh.patch_kind_per_bface()
    num, den = np.zeros(mesh.n_points), np.zeros(mesh.n_points)
    on_patch = np.zeros(mesh.n_points, bool)
    for b in range(mesh.n_bnd):
        if kind[b] == 1:
            continue
        f = nI + b
        for v in mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]:
            w = 1.0 / np.linalg.norm(mesh.points[v] - mesh.Cf[f])
            num[v] += w * bnd[b]; den[v] += w; on_patch[v] = True
    seen = set()
    numc, denc = np.zeros(mesh.n_points), np.zeros(mesh.n_points)
    for f in range(mesh.n_faces):
        for cidx in ([mesh.owner[f]] + ([mesh.neighbour[f]] if f < nI else [])):
            for v in mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]:
                if (v, cidx) in seen:
                    continue
                seen.add((v, cidx))
                w = 1.0 / np.linalg.norm(mesh.points[v] - mesh.C[cidx])
                numc[v] += w * cell[cidx]; denc[v] += w
    ref = np.where(on_patch, num / np.where(den > 0, den, 1.0), numc / denc)
    assert np.abs(got - ref).max() < 1e-13


@pytest.mark.parametrize("mesh_fn", [lambda: cases.pm.hex_box(6, 5, 4, perturb=0.25, grading=(2, 1, 0.5), seed=6),
                                     lambda: cases.pm.hexprism_poly(4, 4, 3, a=0.1, lz=0.4), lambda: cases.case_2d((9, 7), perturb=0.2).mesh])
def test_qgd_length_scales_on_distorted_meshes_match_numpy_restatement(oracle_mod, mesh_fn):
    """QGDCoeffs::updateQGDLength (QGDCoeffs.C:298-362): hQGDf = 2 min(|C_P - Cf|, |C_N - Cf|) on internal faces,
    2/deltaCoeffs on ordinary patch faces; hQGD = area-weighted mean of hQGDf over the cell's non-empty faces."""
    mesh = mesh_fn()
    o = oracle_mod.Oracle(mesh)
    nI = mesh.n_internal
    hf = np.zeros(mesh.n_faces)
    hf[:nI] = 2.0 * np.minimum(np.linalg.norm(mesh.C[mesh.owner[:nI]] - mesh.Cf[:nI], axis=1),
                               np.linalg.norm(mesh.C[mesh.neighbour] - mesh.Cf[:nI], axis=1))
    hf[nI:] = 2.0 / mesh.deltaCoeffs[nI:]
    keep = np.ones(mesh.n_faces, bool)
    keep[nI:] = mesh.patch_kind_per_bface() != 1
    assert np.abs(o.hQGDf()[keep] - hf[keep]).max() < 1e-14
    num, den = np.zeros(mesh.n_cells), np.zeros(mesh.n_cells)
    for cells, sel in ((mesh.owner, keep), (mesh.neighbour, np.ones(nI, bool))):
        f = np.nonzero(sel)[0]
        np.add.at(num, cells[f], hf[f] * mesh.magSf[f])
        np.add.at(den, cells[f], mesh.magSf[f])
    assert np.abs(o.hQGD() - num / den).max() < 1e-14


def test_thermo_state_identities_with_reference_offsets(oracle_mod):
    """hePsiQGDThermo + perfectGas + hConst + sensibleInternalEnergy [OF-v2312]: e = Cp (T - Tref) + Hsref - R T,
    psi = 1/(R T), rho = psi p, c = sqrt(gamma/psi), rhoE = rho (e + |U|^2/2), mu = mu_mol + p ScQGD tau,
    alpha = mu_mol/Pr + (p ScQGD tau)/PrQGD, tauQGD = alphaQGD hQGD / c (constScPrModel1.C:104-114, QGDThermo.C:91-98)."""
    gas = cases.GAS_OFFSET
    c = cases.case_hex3d(n=(5, 4, 3), perturb=0.1, bcs="zg", gas=gas, dt=1e-7)
    c.p0 = 1.0e5 * (1.0 + 0.05 * np.sin(3 * c.mesh.C[:, 0]))
    c.T0 = 300.0 * (1.0 + 0.05 * np.cos(2 * c.mesh.C[:, 1]))
    c.U0 = 30.0 * c.U0 / max(np.abs(c.U0).max(), 1e-30)
    o = c.make_oracle(oracle_mod)
    T, p, e, rho, U = o.get("T"), o.get("p"), o.get("e"), o.get("rho"), o.get("U")
    R, Cp = gas["R"], gas["Cp"]
    g = Cp / (Cp - R)
    assert np.abs(T - c.T0).max() < 1e-9 * 300 and np.array_equal(p, c.p0)
    assert np.abs(e - (Cp * (T - gas["Tref"]) + gas["Hsref"] - R * T)).max() < 1e-10 * np.abs(e).max()
    assert np.abs(rho - p / (R * T)).max() < 1e-13 * rho.max()
    assert np.abs(o.get("c") - np.sqrt(g * R * T)).max() < 1e-12 * 400
    assert np.abs(o.get("rhoE") - rho * (e + 0.5 * (U ** 2).sum(1))).max() < 1e-12 * np.abs(o.get("rhoE")).max()
    tau = 0.5 * o.hQGD() / o.get("c")
    assert np.abs(o.get("tauQGD") - tau).max() < 1e-14 * tau.max()
    muQ = p * gas["ScQGD"] * tau
    assert np.abs(o.get("mu") - (gas["mu"] + muQ)).max() < 1e-13 * muQ.max()
    assert np.abs(o.get("alpha") - (gas["mu"] / gas["Pr"] + muQ / gas["PrQGD"])).max() < 1e-13 * muQ.max()


def _surface_integrate(mesh, phi):
    """fvc::div(ssf)*V = sum over faces: owner +, neighbour -, boundary + (empty faces carry zero flux)"""
    nI = mesh.n_internal
    shape = (mesh.n_cells,) + phi.shape[1:]
    d = np.zeros(shape)
    np.add.at(d, mesh.owner[:nI], phi[:nI])
    np.add.at(d, mesh.neighbour, -phi[:nI])
    np.add.at(d, mesh.owner[nI:], phi[nI:])
    return d


def test_explicit_conservative_update_matches_numpy_restatement(oracle_mod):
    """QGDRhoEqn.H:40-47, QGDUEqn.H:36-51,79-86, QGDEEqn.H:37-50,65-73, QGDFoam.C:142-156 (explicit branch) restated with
    numpy from the oracle's own face fluxes: Euler update of rho, rhoU, rhoE; U = rhoU/rho; the e solve with the
    fvc::ddt(rhoE) quirk; T from e (hConst: linear), psi, p = rho/psi, c, tauQGD and mu with the OLD pressure."""
    gas = dict(cases.GAS, Tref=0.15, Hsref=0.05)
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.2, bcs="fixed", gas=gas)
    m = c.mesh
    o = c.make_oracle(oracle_mod)
    old = {f: o.get(f).copy() for f in ("rho", "rhoU", "rhoE", "U", "e", "p")}
    c.oracle_step(o, 1)
    dt, V = c.dt, m.V
    F = {f: o.get_face(f) for f in ("phiJm", "phiJmU", "phiP", "phiPi", "phiJmH", "phiQ", "phiPiU")}
    rho = old["rho"] - dt / V * _surface_integrate(m, F["phiJm"])
    rhoU = old["rhoU"] - (dt / V)[:, None] * _surface_integrate(m, F["phiJmU"] + F["phiP"] - F["phiPi"])
    rhoE = old["rhoE"] - dt / V * _surface_integrate(m, F["phiJmH"] + F["phiQ"] - F["phiPiU"])
    assert np.abs(o.get("rho") - rho).max() < 1e-13 * rho.max()
    assert np.abs(o.get("rhoU") - rhoU).max() < 1e-13 * np.abs(rhoU).max()
    assert np.abs(o.get("rhoE") - rhoE).max() < 1e-13 * rhoE.max()
    U = (old["rho"][:, None] * old["U"] + (rhoU - old["rhoU"])) / rho[:, None]          # fvm::ddt(rho,U) - fvc::ddt(rhoU) == 0
    e = (old["rho"] * old["e"] + (rhoE - old["rhoE"])) / rho                            # fvm::ddt(rho,e) - fvc::ddt(rhoE) == 0
    assert np.abs(o.get("U") - U).max() < 1e-12 * np.abs(U).max()
    assert np.abs(o.get("e") - e).max() < 1e-13 * e.max()
    R, Cp = gas["R"], gas["Cp"]
    T = (e + Cp * gas["Tref"] - gas["Hsref"]) / (Cp - R)                                # e = Cp (T - Tref) + Hsref - R T
    assert np.abs(o.get("T") - T).max() < 1e-12 * T.max()
    assert np.abs(o.get("p") - rho * R * T).max() < 1e-12 * old["p"].max()
    cs = np.sqrt(Cp / (Cp - R) * R * T)
    tau = 0.5 * o.hQGD() / cs
    assert np.abs(o.get("tauQGD") - tau).max() < 1e-13 * tau.max()
    assert np.abs(o.get("mu") - (gas["mu"] + old["p"] * gas["ScQGD"] * tau)).max() < 1e-13          # old p: QGDFoam.C:149-154


def test_implicit_diffusion_systems_match_a_sparse_direct_solve(oracle_mod):
    """QGDUEqn.H:54-75 and QGDEEqn.H:53-64: the U and e systems assembled independently with scipy.sparse
    (ddt diagonal rho V/dt, -laplacian(muf|alphauf) with nonOrthDeltaCoeffs, fixedValue patches through
    internalCoeffs / boundaryCoeffs, explicit phiTauMC) and solved directly; the oracle's PCG result agrees."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", implicit=True, gas=dict(cases.GAS, mu=2e-2),
                         diff_solver=dict(tol=1e-15, max_iter=5000))
    m = c.mesh
    nI, nC = m.n_internal, m.n_cells
    o = c.make_oracle(oracle_mod)
    old = {f: o.get(f, with_bnd=True) for f in ("rho", "rhoU", "rhoE", "U", "e", "mu", "alpha")}
    c.oracle_step(o, 1)
    dt, V = c.dt, m.V
    F = {f: o.get_face(f) for f in ("phiJm", "phiJmU", "phiP", "phiPi", "phiJmH", "phiQ", "phiPiU")}
    rho = old["rho"][0] - dt / V * _surface_integrate(m, F["phiJm"])
    rhoU = old["rhoU"][0] - (dt / V)[:, None] * _surface_integrate(m, F["phiJmU"] + F["phiP"] - F["phiPi"])
    Us = rhoU / rho[:, None]
    g = c.gas["Cp"] / (c.gas["Cp"] - c.gas["R"])

    def laplacian(gamma_f):
        a = gamma_f * m.magSf * m.nonOrthDeltaCoeffs
        A = sp.coo_matrix((np.concatenate([a[:nI], a[:nI], -a[:nI], -a[:nI]]),
                           (np.concatenate([m.owner[:nI], m.neighbour, m.owner[:nI], m.neighbour]),
                            np.concatenate([m.owner[:nI], m.neighbour, m.neighbour, m.owner[:nI]]))), shape=(nC, nC)).tocsr()
        return A, a
    muf = o.linear_interpolate(*old["mu"])
    A, a = laplacian(muf)
    Unew = o.get("U")
    nb = m.owner[nI:]
    bI = np.zeros(nC); np.add.at(bI, nb, a[nI:])
    bB = np.zeros((nC, 3)); np.add.at(bB, nb, a[nI:, None] * c.bvU)
    MU = (sp.diags(rho * V / dt + bI) + A).tocsc()
    # fvm::ddt(rho,U) - fvc::ddt(rho,U) - fvm::laplacian(muf,U) - fvc::div(phiTauMC) == 0  with U* = rhoU/rho on the right
    rhsU = (rho * V / dt)[:, None] * Us + _surface_integrate(m, o.get_face("phiTauMC")) + bB
    for j in range(3):
        x = spl.spsolve(MU, rhsU[:, j])
        assert np.abs(x - Unew[:, j]).max() < 1e-11 * np.abs(Unew).max(), j
    # e system: rhoE* from the explicit fluxes incl. phiSigmaDotU (QGDEEqn.H:37-50), then the implicit conduction solve (:53-61)
    rhoEs = old["rhoE"][0] - dt / V * _surface_integrate(m, F["phiJmH"] + F["phiQ"] - F["phiPiU"] - o.get_face("phiSigmaDotU"))
    estar = rhoEs / rho - 0.5 * (Unew ** 2).sum(1)
    alphaf = o.linear_interpolate(*old["alpha"]) * (g if c.opts["alpha_eff_gamma_factor"] else 1.0)
    Ae, ae = laplacian(alphaf)
    bIe = np.zeros(nC); np.add.at(bIe, nb, ae[nI:])
    Th = c.bvT
    eB = c.gas["Cp"] * (Th - c.gas["Tref"]) + c.gas["Hsref"] - c.gas["R"] * Th             # fixedEnergy boundary value
    bBe = np.zeros(nC); np.add.at(bBe, nb, ae[nI:] * eB)
    M = (sp.diags(rho * V / dt + bIe) + Ae).tocsc()
    enew = o.get("e")
    x = spl.spsolve(M, rho * V / dt * estar + bBe)
    assert np.abs(x - enew).max() < 1e-11 * enew.max()
    assert np.abs(o.get("rhoE") - rho * (enew + 0.5 * (Unew ** 2).sum(1))).max() < 1e-13 * np.abs(o.get("rhoE")).max()   # QGDEEqn.H:63
    assert np.abs(o.get("rhoU") - rho[:, None] * Unew).max() < 1e-13                       # QGDUEqn.H:70


def test_qhd_temperature_equation_matches_numpy_restatement(oracle_mod):
    """QHDTEqn.H:64-95 (explicit branch) from the oracle's own phi, phiu, tauQGDf and gradTf:
    T_new = T - dt/V sum_f +-[ phi Tf - Hif snGrad(T) |Sf| - tauQGDf phiu (Uf & gradTf) ]."""
    import cases
    c = cases.qhd_cavity(n=(7, 6, 5), dims=3, dt=1e-3, perturb=0.15)
    m = c.mesh
    nI = m.n_internal
    o = c.make_oracle(oracle_mod)
    (U0, UB0), (T0, TB0) = o.qhd_get("U", with_bnd=True), o.qhd_get("T", with_bnd=True)
    c.oracle_step(o, 1)
    f = c.fluid
    Hi = (f["mu"] / f["Pr"]) / f["rho0"]
    Uf, Tf = o.linear_interpolate(U0, UB0), o.linear_interpolate(T0, TB0)
    phi, phiu, tau, gT = (o.qhd_get_face(k) for k in ("phi", "phiu", "tauQGDf", "gradTf"))
    sn = np.zeros(m.n_faces)
    sn[:nI] = m.nonOrthDeltaCoeffs[:nI] * (T0[m.neighbour] - T0[m.owner[:nI]])
    fixed = c.bcT[m.patch_id_per_bface()] == cases.FV
    sn[nI:] = np.where(fixed, m.deltaCoeffs[nI:] * (TB0 - T0[m.owner[nI:]]), 0.0)
    flux = phi * Tf - Hi * sn * m.magSf - tau * phiu * np.einsum("fi,fi->f", Uf, gT)
    Tnew = T0 - c.dt / m.V * _surface_integrate(m, flux)
    assert np.abs(o.qhd_get("T") - Tnew).max() < 1e-13 * np.abs(Tnew).max()


def test_qhd_momentum_equation_matches_numpy_restatement(oracle_mod):
    """QHDUEqn.H:36-85 (explicit branch) restated with numpy from the oracle's phi, tauQGDf, gradUf, gradPf and the solved p:
    ddt(U) + div(phi Uf - Sf & (Uf*Wf)) - laplacian(nu, U) - div((nu Sf) & interpolate(T(grad U))) == -grad(p)/rho + BdFrc."""
    import cases
    c = cases.qhd_cavity(n=(7, 6, 5), dims=3, dt=1e-3, perturb=0.15)
    m = c.mesh
    nI, nC = m.n_internal, m.n_cells
    o = c.make_oracle(oracle_mod)
    (U0, UB0), (T0, TB0) = o.qhd_get("U", with_bnd=True), o.qhd_get("T", with_bnd=True)
    c.oracle_step(o, 1)
    f = c.fluid
    nu, g, beta = f["mu"] / f["rho0"], np.asarray(f["g"]), f["beta"]
    Sf, V = m.Sf, m.V
    Uf = o.linear_interpolate(U0, UB0)
    Bf = o.linear_interpolate(beta * T0[:, None] * g[None], beta * TB0[:, None] * g[None])
    phi, phiu, tau = (o.qhd_get_face(k) for k in ("phi", "phiu", "tauQGDf"))
    gU, gP = o.qhd_get_face("gradUf").reshape(-1, 3, 3), o.qhd_get_face("gradPf")
    Wf = tau[:, None] * (np.einsum("fi,fij->fj", Uf, gU) + gP / f["rho0"] - Bf)
    phiUf = phi[:, None] * Uf - phiu[:, None] * Wf
    sn = np.zeros((m.n_faces, 3))
    sn[:nI] = m.nonOrthDeltaCoeffs[:nI, None] * (U0[m.neighbour] - U0[m.owner[:nI]])
    fixedU = c.bcU[m.patch_id_per_bface()] == cases.FV
    sn[nI:] = np.where(fixedU[:, None], m.deltaCoeffs[nI:, None] * (UB0 - U0[m.owner[nI:]]), 0.0)
    lap = nu * sn * m.magSf[:, None]
    gradU = _surface_integrate(m, np.einsum("fi,fj->fij", Sf, Uf)) / V[:, None, None]             # Gauss linear, grad(U)_ij = d_i U_j
    nrm = Sf[nI:] / m.magSf[nI:, None]
    gP_ = gradU[m.owner[nI:]]
    gradUB = gP_ + np.einsum("bi,bj->bij", nrm, sn[nI:] - np.einsum("bi,bij->bj", nrm, gP_))      # gaussGrad::correctBoundaryConditions
    GTf = o.linear_interpolate(np.swapaxes(gradU, 1, 2).reshape(nC, 9), np.swapaxes(gradUB, 1, 2).reshape(-1, 9)).reshape(-1, 3, 3)
    flux2 = np.einsum("fi,fij->fj", nu * Sf, GTf)
    p, pB = o.qhd_get("p", with_bnd=True)
    # p as it stood when the U equation was formed: before the reference shift of QHDFoam.C:123-131 (a constant: no effect on grad p
    # except through the boundary values, which were shifted alike)
    pf = o.linear_interpolate(p, pB)
    gradp = _surface_integrate(m, Sf * pf[:, None]) / V[:, None]
    rhs = -_surface_integrate(m, phiUf) / V[:, None] + _surface_integrate(m, lap) / V[:, None] + _surface_integrate(m, flux2) / V[:, None] \
        - gradp / f["rho0"] + beta * T0[:, None] * g[None]
    Unew = U0 + c.dt * rhs
    assert np.abs(o.qhd_get("U") - Unew).max() < 1e-11 * np.abs(Unew).max()


def test_qhd_pressure_equation_matches_a_sparse_direct_solve(oracle_mod):
    """QHDpEqn.H:33-48: -laplacian(tauQGDf/rhof, p) == -div(phiu - phiwo) with a fixedValue p patch (no reference cell),
    assembled independently with scipy.sparse and solved directly; the oracle's PCG pressure agrees."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    import cases
    c = cases.qhd_cavity(n=(8, 7, 5), dims=3, dt=1e-3, perturb=0.1, tol=1e-15)
    m = c.mesh
    names = [p.name for p in m.patches]
    pid = m.patch_id_per_bface()
    c.bcP[names.index("yMax")] = cases.FV
    c.bvP[pid == names.index("yMax")] = 0.02
    c.bcP[names.index("xMin")] = cases.FG                      # fixedGradient p on one wall
    c.bvP[pid == names.index("xMin")] = 0.3
    for nm in ("xMax", "yMin", "zMin", "zMax"):
        c.bcP[names.index(nm)] = cases.ZG
    nI, nC = m.n_internal, m.n_cells
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 1)
    phiu, phiwo, tau = (o.qhd_get_face(k) for k in ("phiu", "phiwo", "tauQGDf"))
    a = (tau / c.fluid["rho0"]) * m.magSf * m.nonOrthDeltaCoeffs
    A = sp.coo_matrix((np.concatenate([a[:nI], a[:nI], -a[:nI], -a[:nI]]),
                       (np.concatenate([m.owner[:nI], m.neighbour, m.owner[:nI], m.neighbour]),
                        np.concatenate([m.owner[:nI], m.neighbour, m.neighbour, m.owner[:nI]]))), shape=(nC, nC)).tocsr()
    nb = m.owner[nI:]
    kind = c.bcP[pid]
    dI, rhs = np.zeros(nC), -_surface_integrate(m, phiu - phiwo)
    fv, fg = kind == cases.FV, kind == cases.FG
    np.add.at(dI, nb[fv], a[nI:][fv])
    np.add.at(rhs, nb[fv], a[nI:][fv] * c.bvP[fv])
    np.add.at(rhs, nb[fg], ((tau / c.fluid["rho0"]) * m.magSf)[nI:][fg] * c.bvP[fg])     # gradient() |Sf| tau/rho
    x = spl.spsolve((A + sp.diags(dI)).tocsc(), rhs)
    assert np.abs(o.qhd_get("p") - x).max() < 1e-10 * np.abs(x).max()


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_gaussvolpoint_2d_matches_numpy_restatement(oracle_mod, axis):
    """GaussVolPointBase2D.C:122-169 (ctor, internal faces) and :315-329 (faceGrad) restated literally with numpy on a
    distorted one-cell-thick mesh: ip1 / ip3 = the first two face points on the upper side of the neighbour centre,
    v42 = C_N - C_P, v13 = x_3 - x_1, the direction cosines / sines w.r.t. e1, e2 and
    g_e1 = dfdn c1 - dfdt c2, g_e2 = dfdt c3 - dfdn c4, g_e3 = 0."""
    mesh = cases.case_2d((9, 7), perturb=0.2, axis=axis).mesh
    o = oracle_mod.Oracle(mesh)
    nI = mesh.n_internal
    ie3 = axis
    ie1, ie2 = [d for d in range(3) if d != ie3]
    rng = np.random.default_rng(9)
    phi = np.sin(3 * mesh.C[:, ie1]) + mesh.C[:, ie2] ** 2 + 0.1 * rng.random(mesh.n_cells)
    bnd = np.cos(2 * mesh.Cf[nI:, ie1]) + 0.1 * rng.random(mesh.n_bnd)
    bsg = mesh.deltaCoeffs[nI:] * (bnd - phi[mesh.owner[nI:]])
    got = o.fvsc_grad(phi, bnd, bsg)
    pF = o.vol_point_interpolate(phi, bnd)
    ref = np.zeros((nI, 3))
    for f in range(nI):
        ic2, ic4 = mesh.neighbour[f], mesh.owner[f]
        fv = mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]
        up = [v for v in fv if mesh.points[v][ie3] >= mesh.C[ic2][ie3]]
        ip1, ip3 = up[0], up[1]
        v42, v13 = mesh.C[ic2] - mesh.C[ic4], mesh.points[ip3] - mesh.points[ip1]
        m42, m13 = np.linalg.norm(v42), np.linalg.norm(v13)
        cosa1, cosa2, sina1, sina2 = v42[ie1] / m42, v13[ie1] / m13, v42[ie2] / m42, v13[ie2] / m13
        den = sina2 * cosa1 - sina1 * cosa2
        dfdn, dfdt = (phi[ic2] - phi[ic4]) / m42, (pF[ip3] - pF[ip1]) / m13
        ref[f, ie1] = dfdn * (sina2 / den) - dfdt * (sina1 / den)
        ref[f, ie2] = dfdt * (cosa1 / den) - dfdn * (cosa2 / den)
    assert np.abs(got[:nI] - ref).max() < 1e-12 * np.abs(ref).max()


def test_gaussvolpoint_3d_boundary_faces_use_the_mirrored_ghost(oracle_mod):
    """GaussVolPointBase3D.C:129-154, 391-415, 780-801: on an ordinary patch the 'neighbour' of a boundary quad face is the
    owner centre mirrored through the face centre, v5 = Cn + 2 (Cf - Cn), carrying psi_n = phi_b + snGrad_b |vO - vN| / 2;
    the same six-point formula as on internal faces then applies (closed form of SURVEY A.2)."""
    mesh = cases.pm.hex_box(5, 4, 4, perturb=0.2, seed=12)
    rng = np.random.default_rng(3)
    nI = mesh.n_internal
    phi, bnd, bsg = rng.random(mesh.n_cells), rng.random(mesh.n_bnd), rng.random(mesh.n_bnd) - 0.5      # any patch snGrad
    o = oracle_mod.Oracle(mesh)
    g = o.fvsc_grad(phi, bnd, bsg)[nI:]
    pf = o.vol_point_interpolate(phi, bnd)
    fv = mesh.face_verts.reshape(-1, 4)[nI:]
    p = mesh.points[fv]
    Cn = mesh.C[mesh.owner[nI:]]
    v5 = Cn + 2.0 * (mesh.Cf[nI:] - Cn)
    d = v5 - Cn
    psin = bnd + bsg * np.linalg.norm(Cn - v5, axis=1) * 0.5
    e1, e2 = p[:, 1] - p[:, 3], p[:, 2] - p[:, 0]
    D = (e2 * np.cross(e1, d)).sum(1)
    ref = (np.cross(d, e1) * (pf[fv[:, 0]] - pf[fv[:, 2]])[:, None] + np.cross(d, e2) * (pf[fv[:, 1]] - pf[fv[:, 3]])[:, None]
           + np.cross(e1, e2) * (phi[mesh.owner[nI:]] - psin)[:, None]) / D[:, None]
    assert np.abs(g - ref).max() / np.abs(ref).max() < 1e-12


def test_gaussvolpoint_3d_triangular_faces_match_literal_coefficients(oracle_mod):
    """GaussVolPointBase3D.C:175-229 (triCalcWeights) typed in literally: five coefficients per direction for the three
    face points, the neighbour (index 3) and the owner (index 4), divided by vt = ((p2-p1)^(p3-p1)) & (C_own - C_nei) / 6
    (the dfdxif macro, :488-513)."""
    mesh = cases.pm.prism_box(4, 3, 3, perturb=0.15, seed=4)
    rng = np.random.default_rng(5)
    nI = mesh.n_internal
    phi, bnd = rng.random(mesh.n_cells), rng.random(mesh.n_bnd)
    bsg = mesh.deltaCoeffs[nI:] * (bnd - phi[mesh.owner[nI:]])
    o = oracle_mod.Oracle(mesh)
    g = o.fvsc_grad(phi, bnd, bsg)
    pf = o.vol_point_interpolate(phi, bnd)
    n = 0
    for f in range(nI):
        fv = mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]
        if fv.size != 3:
            continue
        n += 1
        own, nei = mesh.C[mesh.owner[f]], mesh.C[mesh.neighbour[f]]
        p1, p2, p3 = (mesh.points[v] for v in fv)
        x, y, z = 0, 1, 2
        vt = np.dot(np.cross(p2 - p1, p3 - p1), own - nei) / 6.0
        S = 1.0 / 6.0
        atx = [S * ((own[z] - nei[z]) * (p2[y] - p3[y]) + (nei[y] - own[y]) * (p2[z] - p3[z])),
               S * ((nei[y] - own[y]) * (p3[z] - p1[z]) + (own[z] - nei[z]) * (p3[y] - p1[y])),
               S * ((nei[y] - own[y]) * (p1[z] - p2[z]) + (own[z] - nei[z]) * (p1[y] - p2[y])),
               S * (p1[z] * (p2[y] - p3[y]) + p2[z] * (p3[y] - p1[y]) + p3[z] * (p1[y] - p2[y]))]
        aty = [S * ((own[x] - nei[x]) * (p2[z] - p3[z]) + (nei[z] - own[z]) * (p2[x] - p3[x])),
               S * ((nei[z] - own[z]) * (p3[x] - p1[x]) + (own[x] - nei[x]) * (p3[z] - p1[z])),
               S * ((nei[z] - own[z]) * (p1[x] - p2[x]) + (own[x] - nei[x]) * (p1[z] - p2[z])),
               S * (p1[x] * (p2[z] - p3[z]) + p2[x] * (p3[z] - p1[z]) + p3[x] * (p1[z] - p2[z]))]
        atz = [S * ((own[y] - nei[y]) * (p2[x] - p3[x]) + (nei[x] - own[x]) * (p2[y] - p3[y])),
               S * ((nei[x] - own[x]) * (p3[y] - p1[y]) + (own[y] - nei[y]) * (p3[x] - p1[x])),
               S * ((nei[x] - own[x]) * (p1[y] - p2[y]) + (own[y] - nei[y]) * (p1[x] - p2[x])),
               S * (p1[y] * (p2[x] - p3[x]) + p2[y] * (p3[x] - p1[x]) + p3[y] * (p1[x] - p2[x]))]
        vals = [pf[fv[0]], pf[fv[1]], pf[fv[2]], phi[mesh.neighbour[f]], phi[mesh.owner[f]]]
        ref = np.array([sum(a * v for a, v in zip(A + [-A[3]], vals)) / vt for A in (atx, aty, atz)])
        assert np.abs(g[f] - ref).max() < 1e-11 * max(np.abs(ref).max(), 1.0), f
    assert n > 0


@pytest.mark.parametrize("mesh_fn", [lambda: cases.case_2d((9, 8), perturb=0.2).mesh, lambda: cases.case_2d((6, 40)).mesh,
                                     lambda: cases.case_sod(30).mesh])
def test_least_squares_internal_faces_match_literal_restatement(oracle_mod, mesh_fn):
    """extendedFaceStencilFindNeighbours.C:48-84 (cells around the face's points), extendedFaceStencilCalculateWeights.C:60-154
    (wf2 = 1/|df|^2, G = sum wf2 df df + G0 on the empty directions, inverted unless det(G) < 1, minus G0) and
    extendedFaceStencilScalarGrad.C:52-83 (sum wf2 (G df)(phi_c - phi_f); degenerate faces: nf*snGrad) in plain numpy."""
    mesh = mesh_fn()
    nI = mesh.n_internal
    rng = np.random.default_rng(8)
    d = np.nonzero(mesh.geometric_d > 0)[0]
    phi = np.sin(3 * mesh.C[:, d[0]]) + 0.2 * rng.random(mesh.n_cells)
    bnd = rng.random(mesh.n_bnd)
    bsg = mesh.deltaCoeffs[nI:] * (bnd - phi[mesh.owner[nI:]])
    o = oracle_mod.Oracle(mesh)
    got = o.fvsc_grad(phi, bnd, bsg, scheme=oracle_mod.FVSC_SCHEMES["leastSquares"])
    sF = o.linear_interpolate(phi, bnd)
    point_cells = [[] for _ in range(mesh.n_points)]
    for f in range(mesh.n_faces):
        for cidx in [mesh.owner[f]] + ([mesh.neighbour[f]] if f < nI else []):
            for v in mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]:
                if cidx not in point_cells[v]:
                    point_cells[v].append(cidx)
    ndeg = 0
    for f in range(nI):
        cells = []
        for v in mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]:
            for cidx in point_cells[v]:
                if cidx not in cells:
                    cells.append(cidx)
        df = mesh.C[cells] - mesh.Cf[f]
        wf2 = 1.0 / (df * df).sum(1)
        G = np.einsum("c,ci,cj->ij", wf2, df, df)
        G0 = np.diag([1.0 if abs(G[i, i]) < 1e-15 else 0.0 for i in range(3)])
        G = G + G0
        if np.linalg.det(G) < 1:
            ndeg += 1
            nf = mesh.Sf[f] / mesh.magSf[f]
            ref = mesh.nonOrthDeltaCoeffs[f] * (phi[mesh.neighbour[f]] - phi[mesh.owner[f]]) * nf
        else:
            Gi = np.linalg.inv(G) - G0
            ref = np.einsum("c,ci,c->i", wf2, df @ Gi.T, phi[cells] - sF[f])
        assert np.abs(got[f] - ref).max() < 1e-10 * max(np.abs(ref).max(), 1.0), f
    assert ndeg > 0 or mesh.n_cells != 240          # the 6 x 40 mesh (aspect ratio ~7) has degenerate stencils


def test_qgdflux_boundary_condition_closes_the_step_as_the_listing_says(oracle_mod):
    """qgdFluxFvPatchScalarField.C:159-197 + fixedGradient evaluate [OF]: at the p.correctBoundaryConditions() that ends
    the step (QGDFoam.C:155) gradient = -phiwStar/tauQGDf/|Sf| with this step's phiwStar and the tauQGDf just refreshed by
    thermo.correct(), p_b = p_P + gradient/deltaCoeffs; then rho_b = psi_b p_b (QGDFoam.C:156)."""
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="qgdflux")
    c.bcU[:] = cases.ZG                                                   # a moving boundary fluid: phiwStar != 0 on the patches
    m = c.mesh
    nI = m.n_internal
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 3)
    p, pB = o.get("p", with_bnd=True)
    phiw, tauf = o.get_face("phiwStar"), o.get_face("tauQGDf")
    grad = -(phiw[nI:] / tauf[nI:] / m.magSf[nI:])
    assert np.abs(pB - (p[m.owner[nI:]] + grad / m.deltaCoeffs[nI:])).max() < 1e-14 * p.max()
    assert np.abs(grad).max() > 1e-8                                      # the condition is active
    T, TB = o.get("T", with_bnd=True)
    rhoB = o.get("rho", with_bnd=True)[1]
    assert np.abs(rhoB - pB / (c.gas["R"] * TB)).max() < 1e-13
    # with no-slip walls (U_b = 0) every term of rhoW carries Uf or rhoUf: phiwStar and the gradient vanish identically
    w = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="qgdflux")
    ow = w.make_oracle(oracle_mod)
    w.oracle_step(ow, 3)
    assert np.abs(ow.get_face("phiwStar")[nI:]).max() == 0.0
    assert np.array_equal(ow.get("p", with_bnd=True)[1], ow.get("p")[m.owner[nI:]])


def test_decomposed_pcg_semantics(oracle_mod):
    """PCG as a decomposed run performs it (or_pcg_solve_blocks): global matrix and reductions, DIC local to each
    processor block.  One block == the serial solver bit for bit; diagonal / no preconditioning do not see the blocks at
    all (same iterates); block-DIC needs some more iterations than global DIC and reaches the same solution."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from qgdsolver_b200 import decompose
    mesh = cases.pm.hex_box(10, 9, 8, perturb=0.1, seed=1)
    nI, nC = mesh.n_internal, mesh.n_cells
    upper = -(mesh.magSf[:nI] * mesh.deltaCoeffs[:nI])
    diag = np.zeros(nC)
    np.subtract.at(diag, mesh.owner[:nI], upper); np.subtract.at(diag, mesh.neighbour, upper)
    diag += 1e-3 * mesh.V / mesh.V.mean()
    b = np.random.default_rng(0).standard_normal(nC)
    o = oracle_mod.Oracle(mesh)
    x0 = np.zeros(nC)
    blocks = decompose.geometric_split(mesh, 8)
    for pc in (0, 1, 2):
        xs, its, r0s, r1s = o.pcg_solve(diag, upper, b, x0, tol=1e-13, maxIter=3000, precond=pc)
        x1, it1, _, _ = o.pcg_solve(diag, upper, b, x0, tol=1e-13, maxIter=3000, precond=pc, cell_block=np.zeros(nC, np.int32))
        assert it1 == its and np.array_equal(x1, xs)
        xb, itb, r0b, r1b = o.pcg_solve(diag, upper, b, x0, tol=1e-13, maxIter=3000, precond=pc, cell_block=blocks)
        assert r0b == r0s
        if pc < 2:
            assert itb == its and np.array_equal(xb, xs)
        else:
            assert its < itb < 2 * its
            assert np.abs(xb - xs).max() < 1e-9 * np.abs(xs).max()
    A = sp.coo_matrix((np.concatenate([diag, upper, upper]),
                       (np.concatenate([np.arange(nC), mesh.owner[:nI], mesh.neighbour]),
                        np.concatenate([np.arange(nC), mesh.neighbour, mesh.owner[:nI]])))).tocsc()
    assert np.abs(xb - spl.spsolve(A, b)).max() < 1e-9 * np.abs(xb).max()


def test_decomposed_run_oracle_for_qhdfoam_and_implicit_qgdfoam(oracle_mod):
    """set_pcg_blocks: the steps of a decomposed run differ from the serial ones only through the block-local DIC of the
    linear solvers - same fields to solver tolerance, more pressure iterations.  This is the oracle the multi-GPU QHDFoam /
    implicit QGDFoam runs will be compared with."""
    import cases
    from qgdsolver_b200 import decompose
    q = cases.qhd_cavity(n=(12, 10, 6), dims=3, dt=1e-3, perturb=0.1, tol=1e-13)
    blocks = decompose.geometric_split(q.mesh, 4)
    a, b = q.make_oracle(oracle_mod), q.make_oracle(oracle_mod)
    b.set_pcg_blocks(blocks)
    q.oracle_step(a, 5); q.oracle_step(b, 5)
    for f in ("U", "T", "p"):
        assert np.abs(a.qhd_get(f) - b.qhd_get(f)).max() < 1e-9 * np.abs(a.qhd_get(f)).max(), f
    assert b.qhd_solver_info()["iters"] > a.qhd_solver_info()["iters"]
    c = cases.case_hex3d(n=(8, 7, 6), perturb=0.15, bcs="fixed", implicit=True, gas=dict(cases.GAS, mu=2e-2))
    blocks = decompose.geometric_split(c.mesh, 4)
    a, b = c.make_oracle(oracle_mod), c.make_oracle(oracle_mod)
    b.set_pcg_blocks(blocks)
    c.oracle_step(a, 5); c.oracle_step(b, 5)
    for f in ("rho", "rhoU", "rhoE"):
        assert np.abs(a.get(f) - b.get(f)).max() < 1e-11 * np.abs(a.get(f)).max(), f


def test_implicit_branch_explicit_stress_parts_match_numpy_restatement(oracle_mod):
    """updateFluxes.H:107-111 and QGDUEqn.H:72-74 restated with numpy:
    phiTauMC = Sf & interpolate(muEff dev2(T(fvc::grad(U))))  (Gauss-linear cell gradient of the old U, boundary values by
    gaussGrad::correctBoundaryConditions), phiSigmaDotU = Sf & ((muf interpolate(fvc::grad(U_new)) + tauMC) & Uf_old)."""
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", implicit=True, gas=dict(cases.GAS, mu=2e-2))
    m = c.mesh
    nI, nC = m.n_internal, m.n_cells
    o = c.make_oracle(oracle_mod)
    (U0, UB0), (mu0, muB0) = o.get("U", with_bnd=True), o.get("mu", with_bnd=True)
    c.oracle_step(o, 1)
    Sf, V = m.Sf, m.V
    nrm = Sf[nI:] / m.magSf[nI:, None]

    def gauss_grad(U, UB):
        Uf = o.linear_interpolate(U, UB)
        g = _surface_integrate(m, np.einsum("fi,fj->fij", Sf, Uf)) / V[:, None, None]
        sn = m.deltaCoeffs[nI:, None] * (UB - U[m.owner[nI:]])                       # fixedValue patches
        gP = g[m.owner[nI:]]
        gB = gP + np.einsum("bi,bj->bij", nrm, sn - np.einsum("bi,bij->bj", nrm, gP))
        return g, gB, Uf

    def dev2T(mu, g):
        gt = np.swapaxes(g, 1, 2)
        return mu[:, None, None] * (gt - (2.0 / 3.0) * np.trace(g, axis1=1, axis2=2)[:, None, None] * np.eye(3)[None])

    g0, g0B, Uf0 = gauss_grad(U0, UB0)
    tauMC = o.linear_interpolate(dev2T(mu0, g0).reshape(nC, 9), dev2T(muB0, g0B).reshape(-1, 9)).reshape(-1, 3, 3)
    phiTauMC = np.einsum("fi,fij->fj", Sf, tauMC)
    got = o.get_face("phiTauMC")
    assert np.abs(got - phiTauMC).max() < 1e-12 * np.abs(phiTauMC).max()
    U1, UB1 = o.get("U", with_bnd=True)
    g1, g1B, _ = gauss_grad(U1, UB1)
    g1f = o.linear_interpolate(g1.reshape(nC, 9), g1B.reshape(-1, 9)).reshape(-1, 3, 3)
    muf = o.linear_interpolate(mu0, muB0)
    sig = np.einsum("fij,fj->fi", muf[:, None, None] * g1f + tauMC, Uf0)
    phiS = np.einsum("fi,fi->f", Sf, sig)
    assert np.abs(o.get_face("phiSigmaDotU") - phiS).max() < 1e-11 * np.abs(phiS).max()


def test_least_squares_degenerate_face_set(oracle_mod):
    """faceSet degenerateStencilFaces (leastSquaresStencil.C:63-132): the listed internal faces get nf*snGrad, every other face
    keeps its least-squares gradient."""
    mesh = cases.case_2d((10, 9), perturb=0.15).mesh
    nI = mesh.n_internal
    rng = np.random.default_rng(4)
    phi, bnd = rng.random(mesh.n_cells), rng.random(mesh.n_bnd)
    bsg = mesh.deltaCoeffs[nI:] * (bnd - phi[mesh.owner[nI:]])
    sch = oracle_mod.FVSC_SCHEMES["leastSquares"]
    base = oracle_mod.Oracle(mesh).fvsc_grad(phi, bnd, bsg, scheme=sch)
    forced = np.arange(3, nI, 7, dtype=np.int32)
    o = oracle_mod.Oracle(mesh)
    o.set_degenerate_faces(np.concatenate([forced, [nI + 2]]))            # a boundary face in the set is ignored
    got = o.fvsc_grad(phi, bnd, bsg, scheme=sch)
    keep = np.ones(mesh.n_faces, bool); keep[forced] = False
    assert np.array_equal(got[keep], base[keep])
    nf = mesh.Sf[forced] / mesh.magSf[forced, None]
    sn = mesh.nonOrthDeltaCoeffs[forced] * (phi[mesh.neighbour[forced]] - phi[mesh.owner[forced]])
    assert np.abs(got[forced] - sn[:, None] * nf).max() < 1e-14 * np.abs(got).max()
    assert np.abs(got[forced] - base[forced]).max() > 1e-3


def test_power_law_transport_in_the_oracle(oracle_mod):
    """powerLawTransportI.H:120-150: mu = mu0 (T/T0)^k, alphah = mu (1/Pr); the QGD parts are added on top
    (QGDThermo.C:91-98).  With k = 0 and mu0 = mu the run is the constant-transport run bit for bit."""
    gas = dict(cases.GAS, mu=3e-3)
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", gas=gas)
    c.power_law = dict(mu0=3e-3, T0=0.7, k=0.76)
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 5)
    T, p = o.get("T"), o.get("p")
    tau = o.get("tauQGD")
    mol = 3e-3 * (T / 0.7) ** 0.76
    # mu / alpha were formed in thermo.correct() with the pressure before p = rho/psi: reconstruct muQGD from alpha - mu instead
    mu, alpha = o.get("mu"), o.get("alpha")
    muQ = mu - mol
    assert muQ.min() > 0 and np.abs(alpha - (mol * (1.0 / gas["Pr"]) + muQ / gas["PrQGD"])).max() < 1e-15
    assert np.abs(muQ / (gas["ScQGD"] * tau) - p).max() < 1e-2 * p.max()          # muQGD = p_old ScQGD tau, p_old ~ p
    same = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", gas=gas)
    same.power_law = dict(mu0=3e-3, T0=1.0, k=0.0)
    a, b = same.make_oracle(oracle_mod), cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", gas=gas).make_oracle(oracle_mod)
    same.oracle_step(a, 10); same.oracle_step(b, 10)
    assert np.abs(a.get("rho") - b.get("rho")).max() < 1e-14 and np.abs(a.get("rhoE") - b.get("rhoE")).max() < 1e-13
    assert np.abs(o.get("rhoE") - b.get("rhoE")).max() > 1e-9                      # the temperature dependence matters


def test_sutherland_transport_and_econst_thermo_in_the_oracle(oracle_mod):
    """The other two instantiations of psiQGDThermos.C:65-111.  sutherland [OF-v2312 sutherlandTransportI.H]: mu = As sqrt(T)/(1+Ts/T),
    alphah = kappa/Cp with the modified Eucken kappa = mu Cv (1.32 + 1.77 R/Cv).  eConst [OF-v2312 eConstThermoI.H]: Es = Cv (T - Tref)
    + Esref, Cp = Cv + R, gamma = Cp/Cv; with Cv = Cp - R, Esref = Hsref - R Tref it is the hConst gas again (same e(T) up to
    rounding), so the two runs agree to round-off accumulation."""
    gas = dict(cases.GAS, mu=3e-3)
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", gas=gas)
    c.sutherland = dict(As=2.5e-3, Ts=0.4)
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 5)
    T = o.get("T")
    mol = 2.5e-3 * np.sqrt(T) / (1.0 + 0.4 / T)
    Cv = gas["Cp"] - gas["R"]
    amol = mol * Cv * (1.32 + 1.77 * gas["R"] / Cv) / gas["Cp"]
    mu, alpha = o.get("mu"), o.get("alpha")
    muQ = mu - mol
    assert muQ.min() > 0 and np.abs(alpha - (amol + muQ / gas["PrQGD"])).max() < 1e-15
    base = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", gas=gas).make_oracle(oracle_mod)
    c.oracle_step(base, 5)
    assert np.abs(o.get("rhoE") - base.get("rhoE")).max() > 1e-9
    # eConst with the equivalent coefficients == hConst
    g2 = dict(cases.GAS, Tref=0.2, Hsref=0.1)
    h = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", gas=g2)
    e = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", gas=g2)
    e.e_const = dict(Cv=g2["Cp"] - g2["R"], Esref=g2["Hsref"] - g2["R"] * g2["Tref"])
    oh, oe = h.make_oracle(oracle_mod), e.make_oracle(oracle_mod)
    assert np.abs(oe.get("e") - ((g2["Cp"] - g2["R"]) * (oe.get("T") - 0.2) + (0.1 - g2["R"] * 0.2))).max() < 1e-15
    h.oracle_step(oh, 20); e.oracle_step(oe, 20)
    for f in ("rho", "rhoU", "rhoE", "T", "c"):
        assert np.abs(oh.get(f) - oe.get(f)).max() < 1e-12 * np.abs(oh.get(f)).max(), f
    # a genuinely different gas: Cv chosen freely changes gamma = (Cv + R)/Cv and the sound speed c = sqrt(gamma R T)
    e2 = cases.case_hex3d(n=(6, 5, 4), perturb=0.15, bcs="fixed", gas=g2)
    e2.e_const = dict(Cv=1.5, Esref=0.05)
    o2 = e2.make_oracle(oracle_mod)
    assert np.abs(o2.get("c") - np.sqrt((2.5 / 1.5) * g2["R"] * o2.get("T"))).max() < 1e-14
    e2.oracle_step(o2, 10)
    assert np.isfinite(o2.get("rhoE")).all()


def test_slip_walls_in_the_oracle(oracle_mod):
    """slip / symmetryPlane velocity [OF basicSymmetry]: U_b = U_P - n (n . U_P).  A channel with slip side walls, inflow / outflow
    by zeroGradient: a uniform stream along the channel stays exactly uniform (nothing to regularise, no wall shear), the wall-normal
    boundary velocity is zero for any interior field, and the mass flux through the slip walls vanishes."""
    c = cases.case_hex3d(n=(8, 5, 4), bcs="zg", gas=dict(cases.GAS, mu=1e-2))
    m = c.mesh
    names = [p.name for p in m.patches]
    for nm in ("yMin", "yMax", "zMin", "zMax"):
        c.bcU[names.index(nm)] = cases.SLIP
    c.U0 = np.tile([0.3, 0.0, 0.0], (m.n_cells, 1))
    c.p0 = np.full(m.n_cells, 1.0 / 1.4)
    c.T0 = np.full(m.n_cells, (1.0 / 1.4) / c.gas["R"])
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 20)
    assert np.abs(o.get("U") - [0.3, 0.0, 0.0]).max() < 1e-14 and np.abs(o.get("rho") - 1.0).max() < 1e-14
    # a general interior field: the boundary velocity on slip patches has no normal component
    d = cases.case_hex3d(n=(8, 5, 4), perturb=0.2, bcs="zg")
    for nm in ("yMin", "yMax", "zMin", "zMax"):
        d.bcU[names.index(nm)] = cases.SLIP
    od = d.make_oracle(oracle_mod)
    d.oracle_step(od, 5)
    nI = d.mesh.n_internal
    pid = d.mesh.patch_id_per_bface()
    slip = np.isin(pid, [names.index(nm) for nm in ("yMin", "yMax", "zMin", "zMax")])
    UB = od.get("U", with_bnd=True)[1]
    nrm = d.mesh.Sf[nI:] / d.mesh.magSf[nI:, None]
    assert np.abs((UB * nrm).sum(1)[slip]).max() < 1e-15
    UP = od.get("U")[d.mesh.owner[nI:]]
    assert np.abs(UB - (UP - nrm * (UP * nrm).sum(1)[:, None]))[slip].max() < 1e-15
    assert np.abs(UB - UP)[~slip].max() == 0.0                                    # zeroGradient elsewhere


def test_mach3_forward_step_develops_a_bow_shock(oracle_mod):
    """BASELINE configs[1] geometry (polymesh.forward_step) with the Woodward-Colella inflow: after t = 0.5 the flow upstream of
    the bow shock is still the free stream, the pressure in front of the step has risen to the post-shock / stagnation level
    (normal-shock p2/p1 = 10.33 at Mach 3, stagnation 12.06), everything stays positive and total mass changes only by the in- and
    outflow through xMin / xMax (slip walls and the step are impermeable for the convective flux)."""
    c = cases.case_forward_step(n=30)
    m = c.mesh
    assert m.n_cells == 90 * 30 - 72 * 6 and [p.name for p in m.patches][-1] == "step"
    o = c.make_oracle(oracle_mod)
    nsteps = int(round(0.5 / c.dt))
    c.oracle_step(o, nsteps)
    rho, p, U = o.get("rho"), o.get("p"), o.get("U")
    assert np.isfinite(rho).all() and rho.min() > 0.5 and p.min() > 0.3
    up = m.C[:, 0] < 0.1
    assert np.abs(p[up] - 1.0).max() < 0.02 and np.abs(U[up, 0] - 3.0).max() < 0.01          # free stream (the regularisation reaches a little upstream)
    low = m.C[:, 1] < 0.2
    assert p[low & (m.C[:, 0] < 0.3)].max() < 1.3                                            # the bow shock stands at x ~ 0.37 in front of the step
    front = low & (m.C[:, 0] > 0.45) & (m.C[:, 0] < 0.6)
    assert p[front].min() > 9.0 and 11.0 < p[front].max() < 13.0                             # normal shock 10.33, stagnation 12.06
    mid = np.abs(m.C[:, 1] - 0.5) < 0.02
    assert 4.5 < p[mid].max() < 6.5                                                          # the oblique part of the shock further up
    nI = m.n_internal
    names = [q.name for q in m.patches]
    pid = m.patch_id_per_bface()
    walls = np.isin(pid, [names.index(k) for k in ("yMin", "yMax", "step")])
    UB = o.get("U", with_bnd=True)[1]
    assert np.abs((UB * m.Sf[nI:]).sum(1)[walls]).max() < 1e-14


def test_truncated_octahedron_mesh_free_stream_and_conservation(oracle_mod):
    """QGDFoam on the 14-faced polyhedral cells: a uniform state is preserved exactly (closed cells, consistent face gradients
    on squares and hexagons) and a smooth state conserves mass with qgdFlux walls."""
    c = cases.case_truncoct(bcs="zg")
    m = c.mesh
    c.U0 = np.tile([0.2, -0.1, 0.05], (m.n_cells, 1))
    c.p0 = np.full(m.n_cells, 1.0 / 1.4)
    c.T0 = np.full(m.n_cells, (1.0 / 1.4) / c.gas["R"])
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 20)
    assert np.abs(o.get("rho") - 1.0).max() < 1e-13 and np.abs(o.get("U") - [0.2, -0.1, 0.05]).max() < 1e-13
    w = cases.case_truncoct(bcs="qgdflux")
    ow = w.make_oracle(oracle_mod)
    m0 = (ow.get("rho") * w.mesh.V).sum()
    w.oracle_step(ow, 50)
    assert np.isfinite(ow.get("rho")).all() and abs((ow.get("rho") * w.mesh.V).sum() - m0) < 1e-13 * m0
