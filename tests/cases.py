"""Seeded synthetic cases shared by the oracle KATs and the GPU parity tests."""
from __future__ import annotations

import numpy as np

from qgdsolver_b200 import polymesh as pm

FV, ZG, FG, QF = 0, 1, 2, 3   # bc kinds (fixedValue, zeroGradient, fixedGradient, qgdFlux)
SLIP = 6                      # QGD_BC_SLIP / OR_BC_SLIP: slip / symmetryPlane velocity
WEDGE = 7                     # QGD_BC_WEDGE / OR_BC_WEDGE: wedge velocity (U_b = faceT . U_P) on wedge patches

GAS = dict(R=1.0, Cp=3.5, Hf=0.0, Tref=0.0, Hsref=0.0, mu=1.0e-3, Pr=0.71, ScQGD=1.0, PrQGD=1.0)
GAS_OFFSET = dict(R=287.0, Cp=1004.5, Hf=0.0, Tref=298.15, Hsref=0.0, mu=1.8e-5, Pr=0.71, ScQGD=0.7, PrQGD=0.9)


class Case:
    def __init__(self, mesh, U0, T0, p0, bcU, bcT, bcP, bvU, bvT, bvP, gas=GAS, dt=1e-4, scheme="GaussVolPoint",
                 alphaQGD=None, model="constScPrModel1", implicit=False, diff_solver=None, varsc=None, sources=None, **opts):
        self.model, self.implicit = model, implicit
        self.sources = sources          # (rhoSu, rhoUSu, rhoESu) volume-integrated explicit sources or None (createZeroSources.H:28-44)
        self.power_law = None           # dict(mu0, T0, k): powerLaw transport (powerLawTransportI.H:120-150)
        self.sutherland = None          # dict(As, Ts): sutherland transport [OF-v2312]
        self.e_const = None             # dict(Cv, Esref): eConst thermo instead of hConst [OF-v2312]
        # varScModel7 dictionary entries (cSc1, minSc, maxSc) and the constScCellSet cell list
        self.varsc = dict(cSc1=1.0, minSc=-1.0, maxSc=-1.0, const_sc_cells=None)
        if model == "varScModel5":      # dictionary defaults of varScModel5.C:61-66
            self.varsc.update(minSc=0.05, maxSc=1.0, smoothCoeff=0.1, rC=0.5, badQualitySc=0.05, maxAspectRatio=1.5)
        self.varsc.update(varsc or {})
        self.diff_solver = dict(tol=1e-14, rel_tol=0.0, max_iter=2000, precond="DIC")
        self.diff_solver.update(diff_solver or {})
        self.mesh, self.U0, self.T0, self.p0 = mesh, U0, T0, p0
        self.bcU, self.bcT, self.bcP = [np.asarray(x, np.int32) for x in (bcU, bcT, bcP)]
        self.bvU, self.bvT, self.bvP = bvU, bvT, bvP
        self.gas, self.dt, self.scheme, self.alphaQGD = dict(gas), dt, scheme, alphaQGD
        self.opts = dict(alpha_eff_gamma_factor=True, energy_ddt_rhoE_quirk=True, adjust_time_step=False,
                         max_co=0.3, max_delta_t=1e30, c_tau=0.75)
        self.opts.update(opts)

    # ---- CPU oracle
    def make_oracle(self, O, n_threads=1):
        o = O.Oracle(self.mesh, n_threads=n_threads)
        g = self.gas
        prm = O.QGDParams(R=g["R"], Cp=g["Cp"], Hf=g["Hf"], Tref=g["Tref"], Hsref=g["Hsref"], mu=g["mu"], Pr=g["Pr"],
                          ScQGD=g["ScQGD"], PrQGD=g["PrQGD"], implicitDiffusion=int(self.implicit),
                          diffTol=self.diff_solver["tol"], diffRelTol=self.diff_solver["rel_tol"],
                          diffMaxIter=self.diff_solver["max_iter"], diffPrecond=O.PRECONDS[self.diff_solver["precond"]],
                          alphaEffGammaFactor=int(self.opts["alpha_eff_gamma_factor"]),
                          energyDdtRhoEQuirk=int(self.opts["energy_ddt_rhoE_quirk"]), qgdModel=O.QGD_MODELS[self.model],
                          varScCSc1=self.varsc["cSc1"], varScMinSc=self.varsc["minSc"], varScMaxSc=self.varsc["maxSc"])
        if self.model == "varScModel5":
            v = self.varsc
            prm.varSc5SmoothCoeff, prm.varSc5RC = v["smoothCoeff"], v["rC"]
            prm.varSc5BadQualitySc, prm.varSc5MaxAspectRatio = v["badQualitySc"], v["maxAspectRatio"]
        if self.power_law:
            prm.transportModel, prm.mu0, prm.T0, prm.kExp = 1, self.power_law["mu0"], self.power_law["T0"], self.power_law["k"]
        if self.sutherland:
            prm.transportModel, prm.As, prm.Ts = 2, self.sutherland["As"], self.sutherland["Ts"]
        if self.e_const:
            prm.thermoModel, prm.Cv, prm.Esref = 1, self.e_const["Cv"], self.e_const["Esref"]
        scheme = O.FVSC_SCHEMES[self.scheme]
        o.qgd_init(prm, self.bcU, self.bcT, self.bcP, self.bvU, self.bvT, self.bvP, self.U0, self.T0, self.p0,
                   alphaQGD=self.alphaQGD, deltaT=self.dt, scheme=scheme, const_sc_cells=self.varsc["const_sc_cells"])
        if self.sources is not None:
            o.qgd_set_sources(*self.sources)
        return o

    def oracle_step(self, o, n):
        return o.qgd_step(n, adjust=self.opts["adjust_time_step"], maxCo=self.opts["max_co"],
                          maxDeltaT=self.opts["max_delta_t"], cTau=self.opts["c_tau"])

    # ---- product
    def make_solver(self, api, dmesh=None):
        dmesh = dmesh or api.Mesh(self.mesh)
        ds = self.diff_solver
        tk = {}
        if self.power_law:
            tk.update(transport="powerLaw", mu0=self.power_law["mu0"], T0=self.power_law["T0"], k_exp=self.power_law["k"])
        if self.sutherland:
            tk.update(transport="sutherland", As=self.sutherland["As"], Ts=self.sutherland["Ts"])
        if self.e_const:
            tk.update(thermo="eConst", Cv=self.e_const["Cv"], Esref=self.e_const["Esref"])
        if self.model == "varScModel5":
            v = self.varsc
            tk.update(varsc5_smoothCoeff=v["smoothCoeff"], varsc5_rC=v["rC"], varsc5_badQualitySc=v["badQualitySc"],
                      varsc5_maxAspectRatio=v["maxAspectRatio"])
        s = api.QGDFoam(dmesh, fvsc_scheme=self.scheme, qgd_coeffs=self.model, delta_t=self.dt, implicit_diffusion=self.implicit, **tk,
                        diff_tol=ds["tol"], diff_rel_tol=ds["rel_tol"], diff_max_iter=ds["max_iter"], diff_precond=ds["precond"],
                        varsc_cSc1=self.varsc["cSc1"], varsc_minSc=self.varsc["minSc"], varsc_maxSc=self.varsc["maxSc"],
                        **self.gas, **self.opts)
        if self.varsc["const_sc_cells"] is not None:
            s.set_const_sc_cells(self.varsc["const_sc_cells"])
        s.set_bcs(self.bcU, self.bcT, self.bcP, self.bvU, self.bvT, self.bvP)
        s.init_fields(self.U0, self.T0, self.p0, self.alphaQGD)
        if self.sources is not None:
            s.set_sources(*self.sources)
        return s


def smooth_sources(mesh, amp=1.0):
    """smooth volume-integrated sources for rho, rhoU, rhoE (tests of the rhoSu / rhoUSu / rhoESu coupling)"""
    V, x = mesh.V, mesh.C
    act = (mesh.geometric_d > 0).astype(float)
    suRho = amp * 0.3 * V * np.sin(3 * x[:, 0]) ** 2
    suU = amp * (V * 0.2)[:, None] * np.stack([np.cos(2 * x[:, 1]), np.sin(x[:, 0] + x[:, 2]), 0.5 * np.cos(x[:, 0])], 1) * act
    suE = amp * 0.5 * V * (1.0 + np.cos(2 * x[:, 0] + x[:, 1]))
    return suRho, np.ascontiguousarray(suU), suE


def with_sources(case, amp=1.0):
    case.sources = smooth_sources(case.mesh, amp)
    return case


def scalar_transport_case(**kw):
    """scalarTransportQHDFoam case: cavity mesh, a frozen velocity field with a strong divergence, zeroGradient U"""
    c = qhd_cavity(scalar_transport=True, **kw)
    m = c.mesh
    c.U0 = np.ascontiguousarray(np.stack([0.3 * np.sin(3 * m.C[:, 0]) + 0.1, 0.2 * np.cos(2 * m.C[:, 1]) * m.C[:, 0],
                                          0.1 * m.C[:, 2] * (m.geometric_d[2] > 0)], 1))
    c.bcU[:] = ZG
    return c


def smooth_ic(mesh, gas=GAS, seed=12345, mach=0.1):
    """Taylor-Green-like smooth state, non-trivial in all primitives, plus a seeded perturbation."""
    x, y, z = mesh.C[:, 0], mesh.C[:, 1], mesh.C[:, 2]
    g = gas["Cp"] / (gas["Cp"] - gas["R"])
    tp = 2.0 * np.pi
    act = (mesh.geometric_d > 0).astype(float)
    rho = 1.0 + 0.1 * np.sin(tp * x) * np.cos(tp * y * act[1]) * np.cos(tp * z * act[2])
    p = (1.0 / g) * (1.0 + 0.1 * np.cos(tp * x) * np.cos(tp * y * act[1]))
    U = np.stack([mach * np.sin(tp * x) * np.cos(tp * y) * np.cos(tp * z) * act[0],
                  -mach * np.cos(tp * x) * np.sin(tp * y) * np.cos(tp * z) * act[1],
                  0.05 * mach * np.sin(tp * z) * np.cos(tp * x) * act[2]], axis=1)
    rng = np.random.default_rng(seed)
    T = p / (rho * gas["R"]) * (1.0 + 1e-3 * (rng.random(mesh.n_cells) - 0.5))
    return np.ascontiguousarray(U), T, p


def uniform_bcs(mesh, kU=ZG, kT=ZG, kP=ZG, U=(0.0, 0.0, 0.0), T=1.0, p=1.0):
    nP, nB = len(mesh.patches), mesh.n_bnd
    bvU = np.tile(np.asarray(U, float), (nB, 1))
    return (np.full(nP, kU, np.int32), np.full(nP, kT, np.int32), np.full(nP, kP, np.int32),
            bvU, np.full(nB, float(T)), np.full(nB, float(p)))


def case_hex3d(n=(8, 7, 6), perturb=0.0, grading=(1, 1, 1), bcs="zg", gas=GAS, dt=2e-4, **opts):
    mesh = pm.hex_box(*n, perturb=perturb, grading=grading, seed=3)
    return _with_bcs(mesh, bcs, gas, dt, **opts)


def case_prism(n=(5, 4, 4), perturb=0.1, bcs="zg", gas=GAS, **opts):
    return _with_bcs(pm.prism_box(*n, perturb=perturb, seed=5), bcs, gas, 1e-4, **opts)


def case_poly(n=(5, 5, 4), bcs="zg", **opts):
    mesh = pm.hexprism_poly(*n, a=0.1, lz=0.5)
    return _with_bcs(mesh, bcs, GAS, 1e-4, **opts)


def case_2d(n=(24, 20), perturb=0.0, bcs="zg", axis=2, gas=GAS, **opts):
    kinds = {0: {"xMin": "empty", "xMax": "empty"}, 1: {"yMin": "empty", "yMax": "empty"},
             2: {"zMin": "empty", "zMax": "empty"}}[axis]
    dims = [n[0], n[1]]
    dims.insert(axis, 1)
    lengths = [1.0, 1.0]
    lengths.insert(axis, 0.1)
    mesh = pm.hex_box(*dims, lengths=lengths, patch_kinds=kinds, perturb=perturb, seed=7)
    return _with_bcs(mesh, bcs, gas, 2e-4, **opts)


def case_sod(n=200, dt=2e-4, **opts):
    mesh = pm.hex_box(n, 1, 1, lengths=(1.0, 0.1, 0.1),
                      patch_kinds={"zMin": "empty", "zMax": "empty", "yMin": "empty", "yMax": "empty"})
    x = mesh.C[:, 0]
    rho = np.where(x < 0.5, 1.0, 0.125)
    p = np.where(x < 0.5, 1.0, 0.1)
    gas = dict(GAS, mu=0.0, Pr=1.0)
    bc = uniform_bcs(mesh)
    return Case(mesh, np.zeros((n, 3)), p / rho, p, *bc, gas=gas, dt=dt, **opts)


def _with_bcs(mesh, bcs, gas, dt, **opts):
    U0, T0, p0 = smooth_ic(mesh, gas)
    nP, nB = len(mesh.patches), mesh.n_bnd
    if bcs == "zg":
        bc = uniform_bcs(mesh)
    elif bcs == "fixed":       # everything fixedValue (values from a smooth function of the face centre)
        bc = list(uniform_bcs(mesh, FV, FV, FV))
        cf = mesh.Cf[mesh.n_internal:]
        g = gas["Cp"] / (gas["Cp"] - gas["R"])
        bc[3] = np.stack([0.05 * np.sin(cf[:, 1] * 3), 0.05 * np.cos(cf[:, 0] * 2), 0.01 * np.sin(cf[:, 2])], 1) * (mesh.geometric_d > 0)
        bc[4] = (1.0 / g) / gas["R"] * (1.0 + 0.05 * np.sin(cf[:, 0] * 4))
        bc[5] = (1.0 / g) * (1.0 + 0.05 * np.cos(cf[:, 1] * 3))
    elif bcs == "mixed":       # wall-like: U fixed 0, T zeroGradient, p qgdFlux on odd patches; in/outflow on even
        kU = np.array([FV if i % 2 else ZG for i in range(nP)], np.int32)
        kT = np.array([ZG if i % 2 else FV for i in range(nP)], np.int32)
        kP = np.array([QF if i % 2 else ZG for i in range(nP)], np.int32)
        g = gas["Cp"] / (gas["Cp"] - gas["R"])
        bc = (kU, kT, kP, np.zeros((nB, 3)), np.full(nB, (1.0 / g) / gas["R"]), np.full(nB, 1.0 / g))
    elif bcs == "qgdflux":     # all walls: U fixed 0, T zeroGradient, p qgdFlux
        bc = list(uniform_bcs(mesh, FV, ZG, QF))
    else:
        raise ValueError(bcs)
    return Case(mesh, U0, T0, p0, *bc, gas=gas, dt=dt, **opts)


# ============================================================================ QHDFoam cases
QHD_FLUID = dict(rho0=1.0, mu=1.0e-2, Pr=0.71, beta=3.0e-3, g=(0.0, -9.81, 0.0))


class QHDCase:
    """Seeded QHDFoam case: heRhoQGDThermo(rhoConst, hConst, const) + a QHD tau model + PCG controls for p."""

    def __init__(self, mesh, U0, T0, p0, bcU, bcT, bcP, bvU, bvT, bvP, fluid=QHD_FLUID, model="constTau",
                 coeffs=None, dt=1e-3, scheme="GaussVolPoint", alphaQGD=None, tol=1e-13, rel_tol=0.0, max_iter=5000,
                 precond="DIC", p_ref_cell=0, p_ref_value=0.0, adjust_time_step=False, max_co=0.3, max_delta_t=1e30,
                 c_tau=0.75, implicit=False, diff_solver=None, scalar_transport=False):
        self.implicit = implicit
        self.scalar_transport = scalar_transport          # scalarTransportQHDFoam.C:70-135 instead of QHDFoam.C:83-139
        self.diff_solver = dict(tol=1e-14, rel_tol=0.0, max_iter=2000, precond="DIC")
        self.diff_solver.update(diff_solver or {})
        self.mesh, self.U0, self.T0, self.p0 = mesh, U0, T0, p0
        self.bcU, self.bcT, self.bcP = [np.asarray(x, np.int32) for x in (bcU, bcT, bcP)]
        self.bvU, self.bvT, self.bvP = bvU, bvT, bvP
        self.fluid, self.model, self.dt, self.scheme, self.alphaQGD = dict(fluid), model, dt, scheme, alphaQGD
        self.coeffs = dict(Tau=1e-3, UQHD=1.0, Gr=1.0e4, T0=1.0)
        self.coeffs.update(coeffs or {})
        self.solver = dict(tol=tol, rel_tol=rel_tol, max_iter=max_iter, precond=precond)
        self.p_ref_cell, self.p_ref_value = p_ref_cell, p_ref_value
        self.opts = dict(adjust_time_step=adjust_time_step, max_co=max_co, max_delta_t=max_delta_t, c_tau=c_tau)

    def make_oracle(self, O, n_threads=1):
        o = O.Oracle(self.mesh, n_threads=n_threads)
        f, c, sv = self.fluid, self.coeffs, self.solver
        prm = O.QHDParams(rho0=f["rho0"], mu=f["mu"], Pr=f["Pr"], beta=f["beta"], qgdModel=O.QHD_MODELS[self.model],
                          Tau=c["Tau"], UQHD=c["UQHD"], Gr=c["Gr"], T0=c["T0"], implicitDiffusion=int(self.implicit),
                          pTol=sv["tol"], pRelTol=sv["rel_tol"], pMaxIter=sv["max_iter"],
                          pPrecond=O.PRECONDS[sv["precond"]], pRefCell=self.p_ref_cell, pRefValue=self.p_ref_value,
                          diffTol=self.diff_solver["tol"], diffRelTol=self.diff_solver["rel_tol"],
                          diffMaxIter=self.diff_solver["max_iter"], diffPrecond=O.PRECONDS[self.diff_solver["precond"]],
                          scalarTransport=int(self.scalar_transport))
        for j in range(3):
            prm.g[j] = f["g"][j]
        scheme = O.FVSC_SCHEMES[self.scheme]
        o.qhd_init(prm, self.bcU, self.bcT, self.bcP, self.bvU, self.bvT, self.bvP, self.U0, self.T0, self.p0,
                   alphaQGD=self.alphaQGD, deltaT=self.dt, scheme=scheme)
        return o

    def oracle_step(self, o, n):
        return o.qhd_step(n, adjust=self.opts["adjust_time_step"], maxCo=self.opts["max_co"],
                          maxDeltaT=self.opts["max_delta_t"], cTau=self.opts["c_tau"])

    def make_solver(self, api, dmesh=None):
        dmesh = dmesh or api.Mesh(self.mesh)
        ds = self.diff_solver
        s = api.QHDFoam(dmesh, fvsc_scheme=self.scheme, qgd_coeffs=self.model, delta_t=self.dt, p_ref_cell=self.p_ref_cell,
                        p_ref_value=self.p_ref_value, implicit_diffusion=self.implicit, diff_tol=ds["tol"],
                        diff_rel_tol=ds["rel_tol"], diff_max_iter=ds["max_iter"], diff_precond=ds["precond"],
                        scalar_transport=self.scalar_transport, **self.fluid, **self.coeffs, **self.solver, **self.opts)
        s.set_bcs(self.bcU, self.bcT, self.bcP, self.bvU, self.bvT, self.bvP)
        s.init_fields(self.U0, self.T0, self.p0, self.alphaQGD)
        return s


def qhd_cavity(n=(16, 16), dims=2, perturb=0.0, model="constTau", p_bc="qhdflux", seed=21, **kw):
    """Differentially heated cavity (BASELINE configs[2] in miniature): hot xMin wall, cold xMax wall, adiabatic
    others, no-slip U, qhdFlux (== fixedGradient 0 in QHDFoam) or zeroGradient p; smooth non-trivial initial U,T."""
    if dims == 2:
        mesh = pm.hex_box(n[0], n[1], 1, lengths=(1.0, 1.0, 0.1), patch_kinds={"zMin": "empty", "zMax": "empty"},
                          perturb=perturb, seed=seed)
    else:
        mesh = pm.hex_box(*n, perturb=perturb, seed=seed)
    nP, nB = len(mesh.patches), mesh.n_bnd
    names = [p.name for p in mesh.patches]
    kU = np.full(nP, FV, np.int32)
    kT = np.array([FV if nm in ("xMin", "xMax") else ZG for nm in names], np.int32)
    kP = np.full(nP, FG if p_bc == "qhdflux" else ZG, np.int32)
    bvU = np.zeros((nB, 3))
    bvT = np.zeros(nB)
    bvP = np.zeros(nB)
    pid = mesh.patch_id_per_bface()
    for i, nm in enumerate(names):
        if nm == "xMin":
            bvT[pid == i] = 1.0
        if nm == "xMax":
            bvT[pid == i] = 0.0
    x, y, z = mesh.C[:, 0], mesh.C[:, 1], mesh.C[:, 2]
    act = (mesh.geometric_d > 0).astype(float)
    rng = np.random.default_rng(seed)
    psi = np.sin(np.pi * x) ** 2 * np.sin(np.pi * y) ** 2
    U0 = np.stack([0.05 * np.sin(np.pi * x) ** 2 * np.sin(2 * np.pi * y),
                   -0.05 * np.sin(2 * np.pi * x) * np.sin(np.pi * y) ** 2,
                   0.01 * np.sin(np.pi * z) * psi * act[2]], axis=1)
    T0 = 1.0 - x + 0.05 * np.sin(2 * np.pi * y) + 1e-3 * (rng.random(mesh.n_cells) - 0.5)
    p0 = 1e-3 * np.cos(np.pi * x) * np.cos(np.pi * y)
    return QHDCase(mesh, np.ascontiguousarray(U0), T0, p0, kU, kT, kP, bvU, bvT, bvP, model=model, **kw)


def case_forward_step(n=40, co=0.05, **opts):
    """BASELINE configs[1] in miniature: Mach-3 wind tunnel with a forward-facing step (Woodward & Colella set-up: rho = 1.4,
    p = 1, u = 3, gamma = 1.4), inviscid, QGD regularisation only.  Inlet fixedValue, outlet zeroGradient, walls and step:
    slip velocity, zeroGradient T and p.  co: acoustic Courant number (|U|+c) dt / h of the free stream; the
    explicit QGD step needs about 0.05 here (0.2 blows up at the impulsive start: tau |U|^2 acts as a viscosity ~ 4.5 h)."""
    mesh = pm.forward_step(n)
    gas = dict(GAS, mu=0.0, Pr=1.0)
    names = [p.name for p in mesh.patches]
    nP, nB = len(names), mesh.n_bnd
    kU = np.array([FV if nm == "xMin" else (ZG if nm == "xMax" else SLIP) for nm in names], np.int32)
    kT = np.array([FV if nm == "xMin" else ZG for nm in names], np.int32)
    kP = kT.copy()
    T_in = 1.0 / (1.4 * gas["R"])
    bvU = np.tile([3.0, 0.0, 0.0], (nB, 1))
    U0 = np.tile([3.0, 0.0, 0.0], (mesh.n_cells, 1))
    dt = co * (1.0 / n) / 4.0
    return Case(mesh, U0, np.full(mesh.n_cells, T_in), np.ones(mesh.n_cells), kU, kT, kP, bvU, np.full(nB, T_in), np.ones(nB),
                gas=gas, dt=dt, **opts)


def case_truncoct(n=(5, 4, 4), bcs="zg", **opts):
    """QGDFoam on the truncated-octahedron polyhedral mesh (14 faces per cell: hexagon faces take the nf*snGrad branch of
    GaussVolPoint, square faces the six-point formula; cell->face rows longer than the ELL width exercise the CSR tails)"""
    mesh = pm.truncated_octahedron_box(*n, h=1.0 / max(n))
    return _with_bcs(mesh, bcs, GAS, 1e-4, **opts)


def case_wedge(n=(16, 12), angle_deg=5.0, r0=0.5, perturb=0.0, bcs="fixed", gas=GAS, lengths=(1.0, 1.0), **opts):
    """Axisymmetric QGDFoam case on a wedge mesh (polymesh.wedge_box): `wedge` velocity on the two wedge patches (scalars there are
    zeroGradient), the other four patches as `bcs` says; a smooth in-plane initial state (no circumferential velocity)."""
    mesh = pm.wedge_box(n[0], n[1], lengths=lengths, r0=r0, angle_deg=angle_deg, perturb=perturb, seed=13)
    c = _with_bcs(mesh, bcs, gas, 1e-4, **opts)
    for i, p in enumerate(mesh.patches):
        if p.kind == pm.PATCH_WEDGE:
            c.bcU[i], c.bcT[i], c.bcP[i] = WEDGE, ZG, ZG
    x, y = mesh.C[:, 0], mesh.C[:, 1]
    c.U0 = np.ascontiguousarray(np.stack([0.1 * np.sin(2 * np.pi * x) * np.cos(3 * y), 0.08 * np.cos(2 * np.pi * x) * np.sin(3 * y),
                                          np.zeros(mesh.n_cells)], 1))
    return c


def with_symmetry_planes(c, names):
    """the named patches become polyPatch type symmetryPlane (vertex constraint of volPointInterpolation [OF-v2312 pointConstraints],
    leastSquares leaves their faces at zero) with the symmetryPlane field conditions: slip velocity, zeroGradient scalars"""
    m = c.mesh
    m.patches = [pm.Patch(p.name, pm.PATCH_SYMMETRY_PLANE if p.name in names else p.kind, p.start, p.size) for p in m.patches]
    for i, p in enumerate(m.patches):
        if p.name in names:
            c.bcU[i], c.bcT[i], c.bcP[i] = SLIP, ZG, ZG
    return c
