"""Case-directory front end (qgdsolver_b200/foamdict.py, runcase.py) on CPU: the dictionaries the reference solvers read
 are parsed with OpenFOAM's lookup rules and mapped onto the C-ABI descriptors; the mapping is checked by
running the ORACLE from the parsed set-up and comparing with the same case built in memory."""
import os

import numpy as np
import pytest

import cases
from qgdsolver_b200 import foamcase as fc
from qgdsolver_b200 import foamdict, runcase

HDR = "FoamFile\n{\n    version 2.0;\n    format ascii;\n    class dictionary;\n    object %s;\n}\n"

CONTROL = HDR % "controlDict" + """
application     QGDFoam;
startFrom       startTime;
startTime       0;
stopAt          endTime;
endTime         0.02;
deltaT          2e-4;          // fixed
writeControl    runTime;
writeInterval   0.01;
adjustTimeStep  no;
maxCo           0.3;
cTau            0.6;
"""
SCHEMES = HDR % "fvSchemes" + """
ddtSchemes { default Euler; }
gradSchemes { default Gauss linear; }
divSchemes { default none; }
laplacianSchemes { default Gauss linear corrected; }
interpolationSchemes { default none; }
snGradSchemes { default corrected; }
fvsc { default GaussVolPoint; }
"""
THERMO = HDR % "thermophysicalProperties" + """
thermoType
{
    type            hePsiQGDThermo;
    mixture         pureMixture;
    transport       const;
    thermo          hConst;
    equationOfState perfectGas;
    specie          specie;
    energy          sensibleInternalEnergy;
}
mixture
{
    specie { molWeight 8314.47; }        /* R = 1 */
    thermodynamics { Cp 3.5; Hf 0; Tref 0; }
    transport { mu 0; Pr 1; }
}
QGD
{
    implicitDiffusion false;
    QGDCoeffs constScPrModel1;
    constScPrModel1Dict { ScQGD 1; PrQGD 1; }
}
"""


def _write_sod(tmp, n=60):
    c = cases.case_sod(n)
    m = c.mesh
    fc.write_polymesh(m, str(tmp))
    os.makedirs(tmp / "system")
    (tmp / "system" / "controlDict").write_text(CONTROL)
    (tmp / "system" / "fvSchemes").write_text(SCHEMES)
    (tmp / "constant" / "thermophysicalProperties").write_text(THERMO)
    zg = {p.name: ("empty" if p.kind == 1 else "zeroGradient") for p in m.patches}
    fc.write_field(str(tmp / "0" / "U"), m, "U", c.U0, zg)
    fc.write_field(str(tmp / "0" / "T"), m, "T", c.T0, zg)
    fc.write_field(str(tmp / "0" / "p"), m, "p", c.p0, zg)
    return c


def _case_from_setup(s):
    """tests-side stand-in for runcase.make_solver: the same set-up handed to the oracle through cases.Case"""
    k = s.solver_kwargs
    (kU, vU), (kT, vT), (kP, vP) = (s.bc[n] for n in ("U", "T", "p"))
    gas = dict(R=k["R"], Cp=k["Cp"], Hf=k["Hf"], Tref=k["Tref"], Hsref=k["Hsref"], mu=k["mu"], Pr=k["Pr"], ScQGD=k["ScQGD"], PrQGD=k["PrQGD"])
    return cases.Case(s.mesh, s.fields["U"].internal, s.fields["T"].internal, s.fields["p"].internal, kU, kT, kP, vU, vT, vP, gas=gas,
                      dt=k["delta_t"], scheme=k["fvsc_scheme"], model=k["qgd_coeffs"], implicit=k["implicit_diffusion"],
                      adjust_time_step=k["adjust_time_step"], max_co=k["max_co"], max_delta_t=k["max_delta_t"], c_tau=k["c_tau"])


def test_dictionary_parser_follows_openfoam_lookup_rules():
    d = foamdict.parse("""
        a 1.5; // comment
        name word; sw on; v (0 -9.81 0); g [0 1 -2 0 0 0 0] (0 0 -1);
        nu nu [0 2 -1 0 0 0 0] 1e-5;
        lst 3 ( x y z );
        sub { b $a; inner { c $b; } }
        solvers { p { solver PCG; tolerance 1e-8; } "(U|e)" { solver PCG; relTol 0.1; } "(U|T)Final" { $p; } }
        /* block
           comment */
        copy $sub;
    """, "test")
    assert d.scalar("a") == 1.5 and d.word("name") == "word" and d.switch("sw") is True
    assert d.vector("v") == [0.0, -9.81, 0.0] and d.vector("g") == [0.0, 0.0, -1.0] and d.scalar("nu") == 1e-5
    assert d["lst"] == [["x", "y", "z"]]
    assert d.sub_dict("sub").scalar("b") == 1.5 and d.sub_dict("sub").sub_dict("inner").scalar("c") == 1.5
    s = d.sub_dict("solvers")
    assert s.sub_dict("U").scalar("relTol") == 0.1 and s.sub_dict("e").word("solver") == "PCG" and s.found("p") and not s.found("T")
    assert d.sub_dict("copy").scalar("b") == 1.5
    assert d.scalar("missing", 7.0) == 7.0 and d.sub_or_self("noDict") is d
    with pytest.raises(foamdict.FoamDictError) as e:
        d.scalar("missing")
    assert "keyword missing is undefined in dictionary" in str(e.value)
    for bad in ("a 1", "a { b 1;", "#include \"x\"\n", "a $nope;"):
        with pytest.raises(foamdict.FoamDictError):
            foamdict.parse(bad)


def test_sod_case_directory_maps_onto_the_solver_descriptor_and_runs_like_the_in_memory_case(tmp_path, oracle_mod):
    c = _write_sod(tmp_path)
    s = runcase.load_case(str(tmp_path))
    k = s.solver_kwargs
    assert s.solver == "QGDFoam" and s.start_time == "0" and s.end_time == 0.02 and s.write_control == "runTime"
    assert k["fvsc_scheme"] == "GaussVolPoint" and k["qgd_coeffs"] == "constScPrModel1" and k["implicit_diffusion"] is False
    assert k["R"] == 1.0 and k["Cp"] == 3.5 and k["Tref"] == 0.0 and k["Hsref"] == 0.0 and k["mu"] == 0.0 and k["c_tau"] == 0.6
    assert k["adjust_time_step"] is False and k["delta_t"] == 2e-4
    assert list(s.mesh.geometric_d) == [1, -1, -1]
    a, b = _case_from_setup(s).make_oracle(oracle_mod), c.make_oracle(oracle_mod)
    a.qgd_step(100); b.qgd_step(100)
    for f in ("rho", "rhoU", "rhoE", "p"):
        assert np.array_equal(a.get(f), b.get(f)), f


def test_defaults_and_fatal_lookups_follow_the_reference(tmp_path):
    _write_sod(tmp_path)
    th = tmp_path / "constant" / "thermophysicalProperties"
    th.write_text(THERMO.replace("implicitDiffusion false;", "").replace("Tref 0;", ""))
    with pytest.raises(foamdict.FoamDictError) as e:                 # implicitDiffusion defaults to true (QGDThermo.C:61) -> needs solvers
        runcase.load_case(str(tmp_path))
    assert "implicitDiffusion true needs system/fvSolution" in str(e.value)
    (tmp_path / "system" / "fvSolution").write_text(HDR % "fvSolution" + 'solvers { "(U|e)" { solver PCG; preconditioner DIC; tolerance 1e-10; relTol 0; } }')
    s = runcase.load_case(str(tmp_path))
    k = s.solver_kwargs
    assert k["implicit_diffusion"] is True and k["diff_tol"] == 1e-10 and k["diff_precond"] == "DIC" and k["Tref"] == runcase.TSTD
    th.write_text(THERMO.replace("QGDCoeffs constScPrModel1;", "QGDCoeffs noSuchModel;"))
    with pytest.raises(foamdict.FoamDictError) as e:
        runcase.load_case(str(tmp_path))
    assert "Unknown QGD coeffs evaluation approach type noSuchModel" in str(e.value)      # QGDCoeffs.C:72-78
    th.write_text(THERMO.replace("QGDCoeffs constScPrModel1;", ""))
    with pytest.raises(foamdict.FoamDictError) as e:
        runcase.load_case(str(tmp_path))
    assert "keyword QGDCoeffs is undefined" in str(e.value)                               # QGDThermo.C:56 mandatory
    th.write_text(THERMO.replace("constScPrModel1;", "constScPrModel2;").replace("constScPrModel1Dict", "constScPrModel2Dict").replace("ScQGD 1;", ""))
    with pytest.raises(foamdict.FoamDictError) as e:
        runcase.load_case(str(tmp_path))
    assert "keyword ScQGD is undefined" in str(e.value)                                   # constScPrModel2.C:60-61 mandatory
    th.write_text(THERMO)
    (tmp_path / "system" / "fvSchemes").write_text(SCHEMES.replace("default GaussVolPoint;", 'default GaussVolPoint; "grad(p)" reduced;'))
    with pytest.raises(foamdict.FoamDictError):
        runcase.load_case(str(tmp_path))
    (tmp_path / "system" / "fvSchemes").write_text(SCHEMES.replace("fvsc { default GaussVolPoint; }", "fvsc { }"))
    with pytest.raises(foamdict.FoamDictError) as e:
        runcase.load_case(str(tmp_path))
    assert "keyword default is undefined" in str(e.value)                                  # fvsc.C:57


QHD_THERMO = HDR % "thermophysicalProperties" + """
thermoType { type heRhoQGDThermo; mixture pureMixture; transport const; thermo hConst; equationOfState rhoConst; specie specie;
             energy sensibleInternalEnergy; }
mixture
{
    specie { molWeight 28.9; }
    equationOfState { rho 1.0; }
    thermodynamics { Cp 1000; Hf 0; }
    transport { mu 1e-2; Pr 0.71; beta 3e-3; }
}
QGD { implicitDiffusion false; QGDCoeffs constTau; constTauDict { Tau 1e-3; } pRefCell 3; pRefValue 0.25; }
"""


def test_qhd_cavity_case_directory(tmp_path, oracle_mod):
    c = cases.qhd_cavity(n=(10, 8), dt=1e-3, p_ref_cell=3, p_ref_value=0.25, tol=1e-12)
    m = c.mesh
    fc.write_polymesh(m, str(tmp_path))
    os.makedirs(tmp_path / "system")
    (tmp_path / "system" / "controlDict").write_text(CONTROL.replace("QGDFoam", "QHDFoam").replace("2e-4", "1e-3"))
    (tmp_path / "system" / "fvSchemes").write_text(SCHEMES)
    (tmp_path / "system" / "fvSolution").write_text(HDR % "fvSolution" + "solvers { p { solver PCG; preconditioner DIC; tolerance 1e-12; relTol 0; maxIter 5000; } }")
    (tmp_path / "constant" / "thermophysicalProperties").write_text(QHD_THERMO)
    (tmp_path / "constant" / "gravitationalProperties").write_text(HDR % "gravitationalProperties" + "g g [0 1 -2 0 0 0 0] (0 -9.81 0);")
    code = {0: "fixedValue", 1: "zeroGradient", 2: "fixedGradient"}
    names = [p.name for p in m.patches]
    for nm, arr, kinds, vals in (("U", c.U0, c.bcU, c.bvU), ("T", c.T0, c.bcT, c.bvT), ("p", c.p0, c.bcP, c.bvP)):
        types = {n: ("empty" if p.kind == 1 else code[int(k)]) for n, k, p in zip(names, kinds, m.patches)}
        grads = {n: np.zeros((p.size, 3) if nm == "U" else p.size) for n, p in zip(names, m.patches) if types[n] == "fixedGradient"}
        fc.write_field(str(tmp_path / "0" / nm), m, nm, arr, types, vals, gradients=grads)
    s = runcase.load_case(str(tmp_path))
    k = s.solver_kwargs
    assert s.solver == "QHDFoam" and k["qgd_coeffs"] == "constTau" and k["Tau"] == 1e-3 and k["g"] == (0.0, -9.81, 0.0)
    assert k["rho0"] == 1.0 and k["beta"] == 3e-3 and k["p_ref_cell"] == 3 and k["p_ref_value"] == 0.25 and k["precond"] == "DIC"
    assert k["scalar_transport"] is False and k["tol"] == 1e-12 and k["max_iter"] == 5000
    (kU, vU), (kT, vT), (kP, vP) = (s.bc[n] for n in ("U", "T", "p"))
    q = cases.QHDCase(s.mesh, s.fields["U"].internal, s.fields["T"].internal, s.fields["p"].internal, kU, kT, kP, vU, vT, vP,
                      fluid=dict(rho0=k["rho0"], mu=k["mu"], Pr=k["Pr"], beta=k["beta"], g=k["g"]), model=k["qgd_coeffs"],
                      coeffs=dict(Tau=k["Tau"], UQHD=k["UQHD"], Gr=k["Gr"], T0=k["T0"]), dt=k["delta_t"], tol=k["tol"], rel_tol=k["rel_tol"],
                      max_iter=k["max_iter"], precond=k["precond"], p_ref_cell=k["p_ref_cell"], p_ref_value=k["p_ref_value"])
    a, b = q.make_oracle(oracle_mod), c.make_oracle(oracle_mod)
    q.oracle_step(a, 10); c.oracle_step(b, 10)
    for f in ("U", "T", "p"):
        assert np.abs(a.qhd_get(f) - b.qhd_get(f)).max() <= 1e-13 * max(np.abs(b.qhd_get(f)).max(), 1e-30), f
    s2 = runcase.load_case(str(tmp_path), "scalarTransportQHDFoam")
    assert s2.solver_kwargs["scalar_transport"] is True


def test_time_loop_plans_whole_write_intervals():
    """runcase.run with a recording stand-in for the device solver: fixed deltaT -> one C call per write interval"""
    class Fake:
        def __init__(self):
            self.calls, self.t = [], 0.0
        def step(self, n):
            self.calls.append(n); self.t += n * 2e-4
        def scalars(self):
            return dict(deltaT=2e-4, CoNum=0.1, time=self.t)
    fake = Fake()
    setup = runcase.CaseSetup("/nonexistent", "QGDFoam", None, {}, "0", 0.02, 2e-4, "runTime", 0.01, dict(adjust_time_step=False))
    assert runcase.time_name(0.0123456789) == "0.0123457" and runcase.time_name(0.0123456789, 3) == "0.0123" and runcase.time_name(2.0) == "2"
    written = []
    orig_make, orig_write = runcase.make_solver, runcase.write_time
    runcase.make_solver = lambda s, api, dmesh=None: fake
    runcase.write_time = lambda s, sol, t: written.append(runcase.time_name(t)) or runcase.time_name(t)
    try:
        runcase.run(setup, api=None, log=lambda *_: None)
    finally:
        runcase.make_solver, runcase.write_time = orig_make, orig_write
    assert fake.calls == [50, 50] and written == ["0.01", "0.02"]


def test_parallel_writer_puts_owned_results_into_processor_directories(tmp_path):
    """runcase.processor_writer with a stand-in rank solver that returns a known global field in the device sub-mesh
    numbering: processorN/<time>/T read back through cellProcAddressing reassembles the global field, physical boundary
    values land on the right processor patch faces."""
    from qgdsolver_b200 import decompose
    _write_sod(tmp_path)                                   # dictionaries + 0/ fields (mesh replaced below by a 3D box)
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.1, bcs="fixed")
    m = c.mesh
    fc.write_polymesh(m, str(tmp_path))
    types = {p.name: "fixedValue" for p in m.patches}
    for nm, arr, vals in (("U", c.U0, c.bvU), ("T", c.T0, c.bvT), ("p", c.p0, c.bvP)):
        fc.write_field(str(tmp_path / "0" / nm), m, nm, arr, types, vals)
    rank = decompose.geometric_split(m, 2)
    fc.write_decomposed_case(m, rank, str(tmp_path))
    import pytest
    with pytest.raises(runcase.FoamDictError, match="non-orthogonal"):     # `corrected` schemes on a perturbed mesh are refused
        runcase.load_case(str(tmp_path))
    (tmp_path / "system" / "fvSchemes").write_text(SCHEMES.replace("corrected", "uncorrected"))
    setup = runcase.load_case(str(tmp_path))
    rc = runcase._RankCase(setup)
    assert rc.model == "constScPrModel1" and rc.opts["c_tau"] == 0.6 and rc.gas["R"] == 1.0 and rc.U0.shape == (m.n_cells, 3)
    gT = np.sin(5 * m.C[:, 0]) + m.C[:, 1]
    gTb = np.cos(3 * m.Cf[m.n_internal:, 2])
    back = np.full(m.n_cells, np.nan)
    for r in range(2):
        sub = decompose.extended_submeshes(m, rank, ranks=[r])[0]
        proc = decompose.processor_meshes(m, rank, ranks=[r])[0]
        gb = sub.face_global[sub.mesh.n_internal:]

        class Fake:
            def get(self, name, with_bnd=False):
                k3 = name in ("U", "rhoU")
                cells = gT[sub.cell_global]
                bnd = np.where(gb >= m.n_internal, gTb[np.maximum(gb - m.n_internal, 0)], -7.0)
                if k3:
                    cells, bnd = np.stack([cells] * 3, 1), np.stack([bnd] * 3, 1)
                return cells, bnd
        d = runcase.processor_writer(setup, sub, proc)(setup, Fake(), 0.01)
        assert d.endswith(os.path.join(f"processor{r}", "0.01"))
        pm = fc.read_polymesh(str(tmp_path / f"processor{r}"))
        f = fc.read_field(os.path.join(d, "T"), pm)
        back[proc.cell_addr] = f.internal
        for patch in pm.patches:
            if patch.kind == cases.pm.PATCH_PROCESSOR:
                assert f.patch_types[patch.name] == "processor"
            else:
                gfaces = np.abs(proc.face_addr[patch.start:patch.start + patch.size]) - 1 - m.n_internal
                assert np.array_equal(f.patch_values[patch.name], gTb[gfaces])
        fU = fc.read_field(os.path.join(d, "U"), pm)
        assert fU.internal.shape == (proc.cell_addr.size, 3)
    assert np.array_equal(back, gT)
    # reconstructPar: processor time directories -> the undecomposed case
    rec = fc.reconstruct_fields(str(tmp_path), "0.01", ("T", "U"), m)
    assert np.array_equal(rec["T"], gT)
    fT = fc.read_field(str(tmp_path / "0.01" / "T"), m)
    assert np.array_equal(fT.internal, gT)
    nI = m.n_internal
    for p in m.patches:
        assert np.array_equal(fT.patch_values[p.name], gTb[p.start - nI:p.start - nI + p.size]) and fT.patch_types[p.name] == "fixedValue"
    # wrong world size / unsupported solver are refused before any device call
    with pytest.raises(foamdict.FoamDictError):
        runcase.run_parallel(setup, None, 0, 3)
    setup.solver_kwargs["implicit_diffusion"] = True
    with pytest.raises(foamdict.FoamDictError):
        runcase._RankCase(setup)


def test_realistic_fvschemes_keywords_and_scheme_validation(tmp_path):
    """keywords with argument lists (`div(phiJm,U)`, `interpolate(rho)`, `laplacian(taubyrhof,p)`) are single keywords, as in
    OpenFOAM's keyType; ddt / grad / laplacian / snGrad schemes other than what the device assembles are refused"""
    import pytest
    from qgdsolver_b200 import foamdict, runcase
    realistic = HDR % "fvSchemes" + """
ddtSchemes { default Euler; }
gradSchemes { default Gauss linear; grad(p) Gauss linear; }
divSchemes { default none; div(phiJm) Gauss linear; div(phiJm,U) Gauss linear; div((muf*dev2(T(grad(U))))) Gauss linear; }
laplacianSchemes { default Gauss linear corrected; laplacian(taubyrhof,p) Gauss linear uncorrected; }
interpolationSchemes { default none; interpolate(rho) linear; interpolate(U) linear; interpolate((U*rhoU)) linear; }
snGradSchemes { default corrected; }
fvsc { default GaussVolPoint; }
"""
    d = foamdict.parse(realistic, "fvSchemes")
    assert d.sub_dict("divSchemes")._tokens("div(phiJm,U)") == ["Gauss", "linear"]
    assert d.sub_dict("divSchemes")._tokens("div((muf*dev2(T(grad(U)))))") == ["Gauss", "linear"]
    assert d.sub_dict("interpolationSchemes").word("interpolate(rho)") == "linear"
    assert d.sub_dict("laplacianSchemes")._tokens("laplacian(taubyrhof,p)") == ["Gauss", "linear", "uncorrected"]
    assert runcase._check_schemes(d) == "GaussVolPoint"
    import cases
    ortho = cases.pm.hex_box(4, 3, 2)
    skew = cases.pm.hex_box(4, 3, 2, perturb=0.2, seed=1)
    runcase._check_laplacian_schemes(d, ortho)
    with pytest.raises(foamdict.FoamDictError, match="non-orthogonal"):
        runcase._check_laplacian_schemes(d, skew)
    runcase._check_laplacian_schemes(foamdict.parse(realistic.replace(" corrected;", " uncorrected;"), "fvSchemes"), skew)
    for bad, what in (("default Euler;", "default backward;"), ("grad(p) Gauss linear;", "grad(p) leastSquares;"),
                      ("div(phiJm,U) Gauss linear;", "div(phiJm,U) Gauss upwind;"), ("interpolate(U) linear;", "interpolate(U) vanLeer;")):
        with pytest.raises(foamdict.FoamDictError):
            runcase._check_schemes(foamdict.parse(realistic.replace(bad, what), "fvSchemes"))
    with pytest.raises(foamdict.FoamDictError):
        runcase._check_laplacian_schemes(foamdict.parse(realistic.replace("laplacian(taubyrhof,p) Gauss linear uncorrected;",
                                                                           "laplacian(taubyrhof,p) Gauss harmonic corrected;"), "fvSchemes"), ortho)


def test_thermo_type_instantiations_and_slip_patches_from_the_dictionaries(tmp_path):
    """thermoType transport / thermo entries map to the four hePsiQGDThermo instantiations of psiQGDThermos.C:65-111 (anything else
    is refused like the reference's `Unknown psiQGDThermo type`); slip / symmetryPlane velocity patches map to QGD_BC_SLIP, on
    scalars to zeroGradient."""
    _write_sod(tmp_path)
    th = tmp_path / "constant" / "thermophysicalProperties"
    s = runcase.load_case(str(tmp_path))
    assert "transport" not in s.solver_kwargs and s.solver_kwargs["mu"] == 0.0
    th.write_text(THERMO.replace("transport       const;", "transport       sutherland;").replace("transport { mu 0; Pr 1; }", "transport { As 1.4792e-06; Ts 116; }"))
    k = runcase.load_case(str(tmp_path)).solver_kwargs
    assert k["transport"] == "sutherland" and k["As"] == 1.4792e-06 and k["Ts"] == 116.0
    th.write_text(THERMO.replace("transport       const;", "transport       powerLaw;").replace("transport { mu 0; Pr 1; }", "transport { mu0 1.8e-5; T0 300; k 0.76; Pr 0.71; }"))
    k = runcase.load_case(str(tmp_path)).solver_kwargs
    assert (k["transport"], k["mu0"], k["T0"], k["k_exp"], k["Pr"]) == ("powerLaw", 1.8e-5, 300.0, 0.76, 0.71)        # powerLawTransport.C:53-60
    th.write_text(THERMO.replace("thermo          hConst;", "thermo          eConst;").replace("thermodynamics { Cp 3.5; Hf 0; Tref 0; }", "thermodynamics { Cv 2.5; Hf 0; Tref 0; Esref 0.1; }"))
    k = runcase.load_case(str(tmp_path)).solver_kwargs
    assert (k["thermo"], k["Cv"], k["Esref"], k["Cp"]) == ("eConst", 2.5, 0.1, 3.5)
    th.write_text(THERMO.replace("thermo          hConst;", "thermo          eConst;").replace("transport       const;", "transport       sutherland;"))
    with pytest.raises(foamdict.FoamDictError, match="psiQGDThermos"):
        runcase.load_case(str(tmp_path))
    th.write_text(THERMO.replace("thermo          hConst;", "thermo          janaf;"))
    with pytest.raises(foamdict.FoamDictError):
        runcase.load_case(str(tmp_path))
    # slip patches
    c = cases.case_hex3d(n=(4, 3, 3))
    m = c.mesh
    types = {p.name: "zeroGradient" for p in m.patches}
    types[m.patches[1].name], types[m.patches[2].name], types[m.patches[3].name] = "slip", "symmetryPlane", "symmetry"
    fc.write_field(str(tmp_path / "s" / "U"), m, "U", c.U0, types)
    fc.write_field(str(tmp_path / "s" / "p"), m, "p", c.p0, types)
    kU, _ = fc.bc_arrays(m, fc.read_field(str(tmp_path / "s" / "U"), m))
    kP, _ = fc.bc_arrays(m, fc.read_field(str(tmp_path / "s" / "p"), m))
    assert list(kU) == [1, 6, 6, 6, 1, 1] and list(kP) == [1, 1, 1, 1, 1, 1]
