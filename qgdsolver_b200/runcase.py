"""Run an OpenFOAM case directory prepared for QGDFoam / QHDFoam / scalarTransportQHDFoam on the device:
read the dictionaries the reference solvers read, the mesh and the start-time fields, step, write time
directories back.  Host-side harness (the compiled product is libqgd_b200.so; this file only maps a case onto its C ABI).

    python -m qgdsolver_b200.runcase <caseDir> [-solver QGDFoam|QHDFoam|scalarTransportQHDFoam] [-device 0]
    torchrun --nproc-per-node N -m qgdsolver_b200.runcase <caseDir> -parallel     (a decomposePar'd case, QGDFoam explicit)

What is read, key by key (the reference's defaults are kept; a missing mandatory key fails like dictionary::lookup):
  system/controlDict     application, startFrom/startTime, endTime, deltaT, writeControl, writeInterval,
                         adjustTimeStep, maxCo, maxDeltaT (readTimeControls.H [OF]), cTau (setDeltaT-QGDQHD.H:45)
  system/fvSchemes       fvsc { default <scheme>; }  (fvsc.C:47-58); per-term fvsc keys must agree with `default`
                         (the fused step evaluates all gradients with one stencil); interpolationSchemes / divSchemes
                         must leave qgdInterpolate / qgdFlux on their linear branch (QGDInterpolate.H:38-118)
  system/fvSolution      solvers { p; "(U|e|T)" ... } : solver PCG, preconditioner, tolerance, relTol, maxIter
  constant/thermophysicalProperties   thermoType (hePsiQGDThermo | heRhoQGDThermo, pureMixture, const, hConst,
                         perfectGas | rhoConst, sensibleInternalEnergy), mixture {...}, QGD { QGDCoeffs, implicitDiffusion
                         (default true, QGDThermo.C:61), <model>Dict {...}, pRefCell, pRefValue }
  constant/gravitationalProperties (or constant/g)   g      (QHDFoam/createFields.H:96-108)
  <startTime>/U, T, p, alphaQGD
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import foamcase, foamdict
from .foamdict import FoamDictError

RR = 8314.47          # J/(kmol K)  [OF-v2312 thermodynamicConstants::RR]
TSTD = 298.15         # hConstThermo: Tref defaults to Tstd, Hsref to 0  [OF-v2312, unverified here]
QGD_MODELS = ("constScPrModel1", "constScPrModel1n", "constScPrModel2", "varScModel5", "varScModel6", "varScModel7")
QHD_MODELS = ("constTau", "H2bynuQHD", "HbyUQHD", "T0byGr")
FVSC = ("GaussVolPoint", "reduced", "leastSquares", "leastSquaresOpt")


@dataclass
class CaseSetup:
    case_dir: str
    solver: str                                   # QGDFoam | QHDFoam | scalarTransportQHDFoam
    mesh: object
    fields: Dict[str, foamcase.VolField]
    start_time: str
    end_time: float
    delta_t: float
    write_control: str
    write_interval: float
    solver_kwargs: Dict[str, object] = field(default_factory=dict)     # keyword arguments of api.QGDFoam / api.QHDFoam
    bc: Dict[str, tuple] = field(default_factory=dict)                 # name -> (kinds per patch, values per boundary face)
    const_sc_cell_set: Optional[str] = None
    write_binary: bool = False                    # controlDict::writeFormat binary
    time_precision: int = 6                       # controlDict::timePrecision [OF Time]


def _time_dirs(case_dir: str) -> List[str]:
    out = []
    for n in os.listdir(case_dir):
        if os.path.isdir(os.path.join(case_dir, n)):
            try:
                float(n)
                out.append(n)
            except ValueError:
                pass
    return sorted(out, key=float)


def _check_schemes(schemes: foamdict.FoamDict) -> str:
    fvsc = schemes.sub_dict("fvsc")
    scheme = fvsc.word("default")                                     # fvsc.C:57 (lookup fails fatally when absent)
    if scheme not in FVSC:
        raise FoamDictError(f"Unknown Model type {scheme}\n\nValid model types are:\n{len(FVSC)}\n(\n" + "\n".join(FVSC) + "\n)\n")
    for k in fvsc:
        if k != "default" and fvsc.word(k) != scheme:
            raise FoamDictError(f"fvSchemes::fvsc::{k} = {fvsc.word(k)} differs from default = {scheme}: the device step evaluates "
                                "every face gradient with one stencil")
    if schemes.found("interpolationSchemes"):                         # QGDInterpolate.H:38-67
        d = schemes.sub_dict("interpolationSchemes")
        for k in d:
            if d.word(k) not in ("linear", "none"):
                raise FoamDictError(f"fvSchemes::interpolationSchemes::{k} = {d.word(k)}: only linear interpolation is device-native")
    if schemes.found("divSchemes"):                                   # QGDInterpolate.H:76-118
        d = schemes.sub_dict("divSchemes")
        for k in d:
            v = d._tokens(k)
            if v not in (["none"], ["Gauss", "linear"]):
                raise FoamDictError(f"fvSchemes::divSchemes::{k}: qgdFlux with a convection scheme other than central (flux*field_f) "
                                    "is not device-native")
    # the device step (and the oracle) hard-code the discretisation the QGD tutorials use; anything else would run silently
    # with different numerics than the reference, so it is refused like the entries above
    def every(name, ok, what):
        if not schemes.found(name):
            return
        d = schemes.sub_dict(name)
        for k in d:
            v = [t for t in d._tokens(k) if isinstance(t, str)]
            if not ok(v):
                raise FoamDictError(f"fvSchemes::{name}::{k} = {' '.join(v)}: {what}")
    every("ddtSchemes", lambda v: v == ["Euler"], "only Euler time integration is device-native (QGDUEqn.H:36-87 are written for it)")
    every("gradSchemes", lambda v: v in (["Gauss", "linear"], ["none"]), "only Gauss linear is device-native for fvc::grad")
    return scheme


def _check_laplacian_schemes(schemes: foamdict.FoamDict, mesh) -> None:
    """laplacianSchemes / snGradSchemes: the device assembles the uncorrected Laplacian (|Sf| nonOrthDeltaCoeffs, no
    faceFluxCorrection).  `corrected` is the same thing on an orthogonal mesh only, so it is accepted there and refused otherwise."""
    import numpy as np
    nI = mesh.n_internal
    d = mesh.C[mesh.neighbour] - mesh.C[mesh.owner[:nI]]
    cosang = (d * mesh.Sf[:nI]).sum(1) / (np.linalg.norm(d, axis=1) * mesh.magSf[:nI])
    orthogonal = nI == 0 or float(np.abs(1.0 - cosang).max()) < 1e-10
    for name in ("laplacianSchemes", "snGradSchemes"):
        if not schemes.found(name):
            continue
        dct = schemes.sub_dict(name)
        for k in dct:
            v = [t for t in dct._tokens(k) if isinstance(t, str)]
            tail = v[2:] if name == "laplacianSchemes" and v[:2] == ["Gauss", "linear"] else (v if name == "snGradSchemes" else None)
            if v == ["none"]:
                continue
            if tail is None or tail not in (["uncorrected"], ["orthogonal"], ["corrected"]):
                raise FoamDictError(f"fvSchemes::{name}::{k} = {' '.join(v)}: only Gauss linear uncorrected | orthogonal | corrected "
                                    "(orthogonal meshes) is device-native")
            if tail == ["corrected"] and not orthogonal:
                raise FoamDictError(f"fvSchemes::{name}::{k} = {' '.join(v)}: the non-orthogonal correction is not device-native and "
                                    "this mesh is non-orthogonal; use uncorrected (what the device assembles) or an orthogonal mesh")


def _linear_solver(fvsolution: foamdict.FoamDict, name: str, defaults=(1e-6, 0.0, 1000, "DIC")):
    solvers = fvsolution.sub_dict("solvers")
    if not solvers.found(name):
        raise FoamDictError(f'keyword {name} is undefined in dictionary "fvSolution/solvers"')
    d = solvers.sub_dict(name)
    s = d.word("solver")
    if s != "PCG":
        raise FoamDictError(f"fvSolution::solvers::{name}::solver {s}: the device solver is PCG (symmetric matrices)")
    return dict(tol=d.scalar("tolerance", defaults[0]), rel_tol=d.scalar("relTol", defaults[1]),
                max_iter=d.label("maxIter", defaults[2]), precond=d.word("preconditioner", defaults[3]))


def load_case(case_dir: str, solver: Optional[str] = None) -> CaseSetup:
    sysd = os.path.join(case_dir, "system")
    control = foamdict.read(os.path.join(sysd, "controlDict"))
    schemes = foamdict.read(os.path.join(sysd, "fvSchemes"))
    thermo = foamdict.read(os.path.join(case_dir, "constant", "thermophysicalProperties"))
    solver = solver or control.word("application")
    if solver not in ("QGDFoam", "QHDFoam", "scalarTransportQHDFoam"):
        raise FoamDictError(f"application {solver}: only QGDFoam, QHDFoam and scalarTransportQHDFoam run on the device")
    qhd = solver != "QGDFoam"
    # ---- time controls
    times = _time_dirs(case_dir)
    start_from = control.word("startFrom", "startTime")
    if start_from == "latestTime":
        start = times[-1]
    elif start_from == "firstTime":
        start = times[0]
    else:
        st = control.scalar("startTime", 0.0)
        match = [t for t in times if float(t) == st]
        if not match:
            raise FoamDictError(f"no time directory for startTime {st} in {case_dir}")
        start = match[0]
    mesh = foamcase.read_polymesh(case_dir)
    fields = foamcase.read_case_fields(case_dir, mesh, start)
    scheme = _check_schemes(schemes)
    _check_laplacian_schemes(schemes, mesh)
    # ---- thermophysicalProperties
    tt = thermo.sub_dict("thermoType")
    want = {"type": "heRhoQGDThermo" if qhd else "hePsiQGDThermo", "mixture": "pureMixture", "transport": "const", "thermo": "hConst",
            "equationOfState": "rhoConst" if qhd else "perfectGas", "energy": "sensibleInternalEnergy"}
    # psiQGDThermos.C:65-111 instantiates const | sutherland | powerLaw transport with hConst, and const transport with eConst
    psi_combos = {("const", "hConst"), ("sutherland", "hConst"), ("powerLaw", "hConst"), ("const", "eConst")}
    for k, v in want.items():
        if not qhd and k in ("transport", "thermo"):
            continue
        if tt.word(k) != v:
            raise FoamDictError(f"thermoType::{k} {tt.word(k)}: the device-native combination for {solver} is {v}")
    if not qhd and (tt.word("transport"), tt.word("thermo")) not in psi_combos:
        raise FoamDictError(f"thermoType transport {tt.word('transport')} + thermo {tt.word('thermo')}: not one of the hePsiQGDThermo "
                            "instantiations (psiQGDThermos.C:65-111)")
    mix = thermo.sub_dict("mixture")
    tr, th = mix.sub_dict("transport"), mix.sub_dict("thermodynamics")
    qgd = thermo.sub_dict("QGD")
    model = qgd.word("QGDCoeffs")                                     # QGDThermo.C:56
    table = QHD_MODELS if qhd else QGD_MODELS
    if model not in QGD_MODELS + QHD_MODELS:                          # QGDCoeffs.C:70-79
        toc = sorted(QGD_MODELS + QHD_MODELS)
        raise FoamDictError(f"Unknown QGD coeffs evaluation approach type {model}\n\nValid model types are:\n{len(toc)}\n(\n" + "\n".join(toc) + "\n)\n")
    if model not in table:
        raise FoamDictError(f"QGDCoeffs {model} is not a model of {solver}")
    coeffs = qgd.sub_or_self(model + "Dict")                          # QGDCoeffs.C:81-116
    implicit = qgd.switch("implicitDiffusion", True)                  # QGDThermo.C:61
    kw: Dict[str, object] = dict(fvsc_scheme=scheme, qgd_coeffs=model, implicit_diffusion=implicit,
                                 adjust_time_step=control.switch("adjustTimeStep", False), max_co=control.scalar("maxCo", 1.0),
                                 max_delta_t=control.scalar("maxDeltaT", 1e30), c_tau=control.scalar("cTau", 0.75),
                                 delta_t=control.scalar("deltaT"))
    setup = CaseSetup(case_dir, solver, mesh, fields, start, control.scalar("endTime"), control.scalar("deltaT"),
                      control.word("writeControl", "timeStep"), control.scalar("writeInterval", 1.0), kw)
    setup.write_binary = control.word("writeFormat", "ascii") == "binary"
    setup.time_precision = control.label("timePrecision", 6)
    fvsolution = foamdict.read(os.path.join(sysd, "fvSolution")) if os.path.exists(os.path.join(sysd, "fvSolution")) else None
    if not qhd:
        R = RR / mix.sub_dict("specie").scalar("molWeight")
        kw.update(R=R, Hf=th.scalar("Hf", 0.0), Tref=th.scalar("Tref", TSTD))
        if tt.word("thermo") == "eConst":                             # eConstThermo: Cv, Hf, Tref, Esref [OF-v2312]
            kw.update(thermo="eConst", Cv=th.scalar("Cv"), Esref=th.scalar("Esref", 0.0), Cp=th.scalar("Cv") + R, Hsref=0.0)
        else:
            kw.update(Cp=th.scalar("Cp"), Hsref=th.scalar("Hsref", 0.0))
        trm = tt.word("transport")
        if trm == "sutherland":                                       # sutherlandTransport: As, Ts [OF-v2312]
            kw.update(transport="sutherland", As=tr.scalar("As"), Ts=tr.scalar("Ts"), mu=0.0, Pr=1.0)
        elif trm == "powerLaw":                                       # powerLawTransport.C:53-60: mu0, T0, k, Pr
            kw.update(transport="powerLaw", mu0=tr.scalar("mu0"), T0=tr.scalar("T0"), k_exp=tr.scalar("k"), mu=0.0, Pr=tr.scalar("Pr"))
        else:
            kw.update(mu=tr.scalar("mu"), Pr=tr.scalar("Pr"))
        if model in ("constScPrModel1", "constScPrModel1n"):          # constScPrModel1.C:58-89: optional, default 1
            kw.update(ScQGD=coeffs.scalar("ScQGD", 1.0), PrQGD=coeffs.scalar("PrQGD", 1.0))
        else:                                                         # constScPrModel2.C:60-61, varScModel5.C:73-74: mandatory
            kw.update(ScQGD=coeffs.scalar("ScQGD"), PrQGD=coeffs.scalar("PrQGD"))
        if model == "varScModel7":                                    # varScModel7.C:64-73,96-119
            kw.update(varsc_cSc1=coeffs.scalar("cSc1", 1.0), varsc_minSc=coeffs.scalar("minSc", -1.0), varsc_maxSc=coeffs.scalar("maxSc", -1.0))
            if coeffs.found("constScCellSet"):
                setup.const_sc_cell_set = coeffs.word("constScCellSet")
        if model == "varScModel5":                                    # varScModel5.C:61-110,134-149
            kw.update(varsc_minSc=coeffs.scalar("minSc", 0.05), varsc_maxSc=coeffs.scalar("maxSc", 1.0),
                      varsc5_smoothCoeff=coeffs.scalar("smoothCoeff", 0.1), varsc5_rC=coeffs.scalar("rC", 0.5),
                      varsc5_badQualitySc=coeffs.scalar("badQualitySc", 0.05), varsc5_maxAspectRatio=coeffs.scalar("maxAspectRatio", 1.5))
            if coeffs.found("constScCellSet"):
                setup.const_sc_cell_set = coeffs.word("constScCellSet")
        if implicit:
            if fvsolution is None:
                raise FoamDictError("implicitDiffusion true needs system/fvSolution (solvers for U and e)")
            su, se = _linear_solver(fvsolution, "U", (1e-9, 0.0, 1000, "DIC")), _linear_solver(fvsolution, "e", (1e-9, 0.0, 1000, "DIC"))
            if su != se:
                raise FoamDictError("fvSolution::solvers: U and e must share their PCG controls on the device")
            kw.update(diff_tol=su["tol"], diff_rel_tol=su["rel_tol"], diff_max_iter=su["max_iter"], diff_precond=su["precond"])
    else:
        eos = mix.sub_dict("equationOfState")
        gpath = [p for p in (os.path.join(case_dir, "constant", "gravitationalProperties"), os.path.join(case_dir, "constant", "g"))
                 if os.path.exists(p)]
        if not gpath:
            raise FoamDictError("constant/gravitationalProperties (g) is missing")              # QHDFoam/createFields.H:96-108
        gd = foamdict.read(gpath[0])
        g = gd.vector("g") if gd.found("g") else gd.vector("value")
        kw.update(rho0=eos.scalar("rho"), mu=tr.scalar("mu"), Pr=tr.scalar("Pr"), beta=tr.scalar("beta"), g=tuple(g),
                  Tau=coeffs.scalar("Tau") if model == "constTau" else 0.0,                   # constTau.C:71 (mandatory)
                  UQHD=coeffs.scalar("UQHD") if model == "HbyUQHD" else 1.0,                  # HbyUQHD.C:61
                  Gr=coeffs.scalar("Gr") if model == "T0byGr" else 1.0, T0=coeffs.scalar("T0") if model == "T0byGr" else 1.0,
                  p_ref_cell=qgd.label("pRefCell", 0), p_ref_value=qgd.scalar("pRefValue", 0.0),
                  scalar_transport=(solver == "scalarTransportQHDFoam"))
        if fvsolution is None:
            raise FoamDictError("QHDFoam needs system/fvSolution (solver for p)")
        if solver == "QHDFoam":
            sp = _linear_solver(fvsolution, "p")
            kw.update(tol=sp["tol"], rel_tol=sp["rel_tol"], max_iter=sp["max_iter"], precond=sp["precond"])
        if implicit:
            names = ("T",) if solver == "scalarTransportQHDFoam" else ("U", "T")
            ss = [_linear_solver(fvsolution, n, (1e-9, 0.0, 1000, "DIC")) for n in names]
            if any(s != ss[0] for s in ss):
                raise FoamDictError("fvSolution::solvers: U and T must share their PCG controls on the device")
            kw.update(diff_tol=ss[0]["tol"], diff_rel_tol=ss[0]["rel_tol"], diff_max_iter=ss[0]["max_iter"], diff_precond=ss[0]["precond"])
    for n in ("U", "T", "p"):
        setup.bc[n] = foamcase.bc_arrays(mesh, fields[n])
    return setup


def make_solver(setup: CaseSetup, api, dmesh=None):
    """api.QGDFoam / api.QHDFoam initialised from the case (the calls a shim makes through the C ABI)"""
    dmesh = dmesh or api.Mesh(setup.mesh)
    deg = os.path.join(setup.case_dir, "constant", "polyMesh", "sets", "degenerateStencilFaces")     # leastSquaresStencil.C:63-70 READ_IF_PRESENT
    if os.path.exists(deg) and setup.solver_kwargs["fvsc_scheme"] == "leastSquares":
        dmesh.set_degenerate_stencil_faces(foamcase.read_labels(deg))
    cls = api.QGDFoam if setup.solver == "QGDFoam" else api.QHDFoam
    s = cls(dmesh, **setup.solver_kwargs)
    if setup.const_sc_cell_set:
        path = os.path.join(setup.case_dir, "constant", "polyMesh", "sets", setup.const_sc_cell_set)
        s.set_const_sc_cells(foamcase.read_labels(path))
    (kU, vU), (kT, vT), (kP, vP) = (setup.bc[n] for n in ("U", "T", "p"))
    s.set_bcs(kU, kT, kP, vU, vT, vP)
    f = setup.fields
    s.init_fields(f["U"].internal, f["T"].internal, f["p"].internal, f["alphaQGD"].internal if "alphaQGD" in f else None)
    return s


def time_name(t: float, precision: int = 6) -> str:
    """Time::timeName: general format with controlDict::timePrecision significant digits (default 6) [OF]"""
    return f"{t:.{precision}g}"


def write_time(setup: CaseSetup, s, t: float) -> str:
    """<time>/ fields as the reference's AUTO_WRITE objects (createFields.H): U, T, p (+ rho, e, rhoU, rhoE for QGDFoam)"""
    d = os.path.join(setup.case_dir, time_name(t, setup.time_precision))
    names = ("U", "T", "p", "rho", "e", "rhoU", "rhoE") if setup.solver == "QGDFoam" else ("U", "T", "p")
    for n in names:
        cells, bnd = s.get(n, with_bnd=True)
        src = setup.fields.get(n)
        types = dict(src.patch_types) if src is not None else {}
        foamcase.write_field(os.path.join(d, n), setup.mesh, n, cells, types, bnd, src.dimensions if src is not None else "[0 0 0 0 0 0 0]",
                             binary=setup.write_binary, gradients=src.patch_gradients if src is not None else None)
    return d


def run(setup: CaseSetup, api, log=print, solver=None, writer=None) -> List[str]:
    """the time loop: fixed deltaT -> whole write intervals per C call (no host sync inside); adjustTimeStep -> chunks of
    at most 50 steps, the time being read back after each (write times are not snapped to multiples like
    Time::adjustDeltaT does: the first step at or past a write time is written).
    solver / writer: a ready solver object and a `writer(setup, solver, t) -> path` (the parallel driver passes its own)."""
    s = solver if solver is not None else make_solver(setup, api)
    writer = writer or write_time
    t, dt = float(setup.start_time), setup.delta_t
    written = []
    adjust = bool(setup.solver_kwargs["adjust_time_step"])
    if setup.write_control == "timeStep":
        every_t = None
        every_n = max(1, int(round(setup.write_interval)))
    elif setup.write_control in ("runTime", "adjustableRunTime", "adjustable"):
        every_t, every_n = setup.write_interval, None
    else:
        raise FoamDictError(f"controlDict::writeControl {setup.write_control} is not supported")
    next_write = t + every_t if every_t else None
    steps_since_write = 0
    eps = 1e-9 * max(abs(setup.end_time), dt)
    while t < setup.end_time - eps:
        if adjust:
            n = 1 if every_n is None else min(50, every_n - steps_since_write)
            n = max(1, min(n, 50))
        else:
            to_end = int(round((setup.end_time - t) / dt))
            if every_n is not None:
                n = min(every_n - steps_since_write, to_end)
            else:
                n = min(max(1, int(round((next_write - t) / dt))), to_end)
            n = max(1, n)
        s.step(n)
        sc = s.scalars()
        t = float(setup.start_time) + sc["time"]              # the device clock starts at 0 when the solver is created
        steps_since_write += n
        due = (every_n is not None and steps_since_write >= every_n) or (every_t is not None and t >= next_write - eps)
        if due or t >= setup.end_time - eps:
            written.append(writer(setup, s, t))
            log(f"Time = {time_name(t, setup.time_precision)}  deltaT = {sc['deltaT']:.6g}  Courant = {sc['CoNum']:.4g}")
            steps_since_write = 0
            if every_t is not None:
                while next_write <= t + eps:
                    next_write += every_t
    return written


class _RankCase:
    """the attributes multigpu.make_rank_solver reads, filled from a CaseSetup (QGDFoam, explicit branch)"""

    def __init__(self, setup: CaseSetup):
        k = setup.solver_kwargs
        if setup.solver != "QGDFoam" or k["implicit_diffusion"]:
            raise FoamDictError("-parallel: only QGDFoam with implicitDiffusion false runs on several GPUs yet")
        self.mesh, self.scheme, self.model, self.dt = setup.mesh, k["fvsc_scheme"], k["qgd_coeffs"], k["delta_t"]
        self.gas = {n: k[n] for n in ("R", "Cp", "Hf", "Tref", "Hsref", "mu", "Pr", "ScQGD", "PrQGD")}
        self.gas.update({n: k[n] for n in ("transport", "As", "Ts", "mu0", "T0", "k_exp", "thermo", "Cv", "Esref") if n in k})
        self.opts = {n: k[n] for n in ("adjust_time_step", "max_co", "max_delta_t", "c_tau")}
        self.varsc = dict(cSc1=k.get("varsc_cSc1", 1.0), minSc=k.get("varsc_minSc", -1.0), maxSc=k.get("varsc_maxSc", -1.0), const_sc_cells=None)
        if setup.const_sc_cell_set:
            self.varsc["const_sc_cells"] = foamcase.read_labels(os.path.join(setup.case_dir, "constant", "polyMesh", "sets", setup.const_sc_cell_set))
        (self.bcU, self.bvU), (self.bcT, self.bvT), (self.bcP, self.bvP) = (setup.bc[n] for n in ("U", "T", "p"))
        f = setup.fields
        self.U0, self.T0, self.p0 = f["U"].internal, f["T"].internal, f["p"].internal
        self.alphaQGD = f["alphaQGD"].internal if "alphaQGD" in f else None


def processor_writer(setup: CaseSetup, sub, proc):
    """writer for one rank of a decomposed run: processorN/<time>/ fields of the rank's owned cells.  `sub` is the rank's
    decompose.SubDomain (device numbering), `proc` its decompose.ProcMesh (the decomposePar layout on disk): owned cells are
    in cellProcAddressing order on both sides; boundary values are matched through the global face ids."""
    nIl = proc.mesh.n_internal
    gf_proc = np.abs(proc.face_addr[nIl:].astype(np.int64)) - 1                  # global face of every processor-mesh boundary face
    gf_sub = sub.face_global[sub.mesh.n_internal:]
    where = {int(g): i for i, g in enumerate(gf_sub)}
    phys = gf_proc >= setup.mesh.n_internal
    idx = np.array([where.get(int(g), -1) if ph else -1 for g, ph in zip(gf_proc, phys)], np.int64)
    if (idx[phys] < 0).any():
        raise FoamDictError("processor mesh and device sub-mesh disagree on the physical boundary faces")
    names = ("U", "T", "p", "rho", "e", "rhoU", "rhoE")

    def write(setup_, s, t):
        d = os.path.join(setup.case_dir, f"processor{proc.rank}", time_name(t, setup.time_precision))
        for n in names:
            cells, bnd = s.get(n, with_bnd=True)
            own = cells[:sub.n_owned]
            pb = np.where(phys.reshape((-1,) + (1,) * (bnd.ndim - 1)), bnd[np.maximum(idx, 0)], own[proc.mesh.owner[nIl:]])
            src = setup.fields.get(n)
            types = dict(src.patch_types) if src is not None else {}
            for patch in proc.mesh.patches:
                if patch.kind == foamcase.PATCH_PROCESSOR:
                    types[patch.name] = "processor"
            foamcase.write_field(os.path.join(d, n), proc.mesh, n, own, types, pb, src.dimensions if src is not None else "[0 0 0 0 0 0 0]",
                                 binary=setup.write_binary,
                                 gradients=foamcase.proc_patch_gradients(setup.mesh, src.patch_gradients, proc) if src is not None else None)
        return d
    return write


def run_parallel(setup: CaseSetup, api, rank: int, world: int, log=print) -> List[str]:
    """one rank of `mpirun -np N QGDFoam -parallel`: the cell -> processor map comes from the decomposePar'd case
    (constant/cellDecomposition or the cellProcAddressing lists), results go to processor<rank>/<time>/"""
    from . import decompose, multigpu
    cell_rank = foamcase.read_cell_decomposition(setup.case_dir, setup.mesh.n_cells)
    if int(cell_rank.max()) + 1 != world:
        raise FoamDictError(f"case is decomposed into {int(cell_rank.max()) + 1} processors, launched with {world} ranks")
    s, sub, _dm = multigpu.make_rank_solver(_RankCase(setup), rank, world, cell_rank)
    proc = decompose.processor_meshes(setup.mesh, cell_rank, ranks=[rank])[0]
    return run(setup, api, log if rank == 0 else (lambda *_: None), solver=s, writer=processor_writer(setup, sub, proc))


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "-help", "--help"):
        print(__doc__)
        return 0
    case_dir, solver, device, parallel = argv[0], None, 0, False
    i = 1
    while i < len(argv):
        if argv[i] == "-solver":
            solver = argv[i + 1]; i += 2
        elif argv[i] == "-device":
            device = int(argv[i + 1]); i += 2
        elif argv[i] == "-parallel":      # under torchrun: one rank per GPU (RANK / LOCAL_RANK / WORLD_SIZE from the environment)
            parallel = True; i += 1
        else:
            raise SystemExit(f"unknown option {argv[i]}")
    from . import api
    if parallel:
        import torch
        import torch.distributed as dist
        from . import multigpu
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        api.init(local)
        multigpu.init_comm(dist.get_rank(), dist.get_world_size())
        run_parallel(load_case(case_dir, solver), api, dist.get_rank(), dist.get_world_size())
        api.synchronize()
        dist.barrier()
        api.comm_finalize()
        dist.destroy_process_group()
        return 0
    api.init(device)                      # fails loudly without a CUDA device: there is no CPU fallback
    setup = load_case(case_dir, solver)
    run(setup, api)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
