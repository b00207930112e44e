"""varScModel5, product side, without a GPU: the device functors and launch sequences of qgdsolver_b200/csrc/qgd_varsc5.h are
compiled with g++ and run under a serial host executor (tests/host_setup/host_setup_shim.cpp) - the very code the CUDA executor
launches kernel by kernel - and compared with the oracle: the order-exact parallel form of fvc::smooth bit for bit, the
mesh-quality floor, and one whole varScModel5::correct on a state the oracle produced.  The thread order inside every emulated
launch is varied (ascending, descending, strided): a data race between the items of one launch would show as a difference."""
import ctypes as C

import numpy as np
import pytest

import cases
from test_host_setup_cpu import Host, shim  # noqa: F401  (fixture)
from test_oracle_varsc5 import MESHES

_dp = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(_dp)


def _bind(L):
    L.hs_varsc5_quality.restype = C.c_int
    L.hs_varsc5_quality.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, _dp]
    L.hs_varsc5_smooth.restype = C.c_int
    L.hs_varsc5_smooth.argtypes = [C.c_void_p, _dp, C.c_double, C.c_int, C.POINTER(C.c_longlong)]
    L.hs_varsc5_correct.restype = C.c_int
    L.hs_varsc5_correct.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_char_p, _dp, _dp, _dp, C.c_int]
    return L


@pytest.mark.parametrize("name", sorted(MESHES))
def test_device_form_of_fvc_smooth_equals_the_sequential_oracle(shim, oracle_mod, name):
    L = _bind(shim)
    mesh = MESHES[name]()
    host = Host(L, mesh)
    o = oracle_mod.Oracle(mesh)
    rng = np.random.default_rng(17)
    for trial in range(4):
        f = 0.05 + 0.02 * rng.random(mesh.n_cells)
        spikes = rng.choice(mesh.n_cells, size=max(2, mesh.n_cells // 20), replace=False)
        f[spikes] = 0.3 + 0.7 * rng.random(spikes.size)
        if trial == 3:
            f[rng.choice(mesh.n_cells, size=3, replace=False)] = 0.0
        coeff = [0.1, 0.3, 0.02, 0.1][trial]
        want, it_o = o.fvc_smooth(f, coeff)
        for order in (0, 1, 2):
            got = f.copy()
            n_launch = C.c_longlong()
            it = L.hs_varsc5_smooth(host.h, _d(got), coeff, order, C.byref(n_launch))
            assert np.array_equal(got, want), (name, trial, order)
            assert it_o - 1 <= it <= it_o          # the reference counts one more sweep when only boundary faces were left
            assert n_launch.value > 4


def test_device_form_of_fvc_smooth_on_a_larger_mesh_with_long_waves(shim, oracle_mod):
    """30 x 30 x 30 hex, one strong spike in a corner: the wave crosses the mesh (tens of sweeps, thousands of changed items per sweep,
    compaction over more than one chunk per thread)"""
    L = _bind(shim)
    mesh = cases.pm.hex_box(30, 30, 30, perturb=0.1, seed=1)
    host = Host(L, mesh)
    o = oracle_mod.Oracle(mesh)
    rng = np.random.default_rng(3)
    f = 0.05 * (1.0 + 0.05 * rng.random(mesh.n_cells))
    f[0] = 1.0
    f[mesh.n_cells // 2 + 7] = 0.9
    want, it_o = o.fvc_smooth(f, 0.1)
    got = f.copy()
    it = L.hs_varsc5_smooth(host.h, _d(got), 0.1, 2, None)
    assert it_o > 20 and it_o - 1 <= it <= it_o
    assert np.array_equal(got, want)
    assert (want > f).sum() > 5000


@pytest.mark.parametrize("name", ["hex", "truncoct", "2d", "prism"])
def test_quality_floor_matches_the_oracle(shim, oracle_mod, name):
    L = _bind(shim)
    mesh = MESHES[name]()
    host = Host(L, mesh)
    q, ar = np.zeros(mesh.n_cells), np.zeros(mesh.n_cells)
    mcf = L.hs_varsc5_quality(host.h, 0.05, 1.2, _d(q), _d(ar))
    qo, aro = oracle_mod.Oracle(mesh).varsc5_cell_quality(0.05, 1.2)
    assert np.array_equal(ar, aro) and np.array_equal(q, qo)
    counts = np.bincount(mesh.owner, minlength=mesh.n_cells) + np.bincount(mesh.neighbour, minlength=mesh.n_cells)
    assert mcf == counts.max()


@pytest.mark.parametrize("case_fn", [
    lambda: cases.case_hex3d(n=(7, 6, 5), perturb=0.2, bcs="fixed", model="varScModel5", varsc=dict(rC=0.35, smoothCoeff=0.15, maxAspectRatio=1.2)),
    lambda: cases.case_hex3d(n=(6, 6, 5), perturb=0.1, bcs="mixed", model="varScModel5", gas=cases.GAS_OFFSET, dt=1e-6),
    lambda: cases.case_2d((12, 10), perturb=0.2, bcs="mixed", model="varScModel5", varsc=dict(const_sc_cells=np.array([3, 50, 77], np.int32))),
    lambda: cases.case_truncoct(n=(4, 3, 3), bcs="zg", model="varScModel5", varsc=dict(smoothCoeff=0.05)),
    lambda: cases.case_prism(bcs="fixed", model="varScModel5", implicit=True),
], ids=["hex_fixed", "hex_mixed_offsets", "2d_mixed_cellset", "truncoct", "prism_implicit"])
def test_device_form_of_the_whole_correct_matches_the_oracle(shim, oracle_mod, case_fn):
    """state before / after one oracle step -> the inputs the device pass sees (T, c and boundary state closed by the ordinary
    kernels, old p and p_b, old ScQGD) -> v5Correct -> ScQGD, mu, alphaEff and the tau slot against the oracle's fields"""
    L = _bind(shim)
    c = case_fn()
    mesh = c.mesh
    host = Host(L, mesh)
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 2)
    nC, nB = mesh.n_cells, mesh.n_bnd
    R, g = c.gas["R"], c.gas["Cp"] / (c.gas["Cp"] - c.gas["R"])
    v = c.varsc
    for order in (0, 2):
        sc0, sc0b = o.get("ScQGD", with_bnd=True)
        p0, _ = o.get("p", with_bnd=True)
        tauf0 = o.get_face("tauQGDf")                   # the tauQGDf this step's fluxes (and the qgdFlux patches) use
        c.oracle_step(o, 1)
        T1, T1b = o.get("T", with_bnd=True)
        cs, csb = o.get("c", with_bnd=True)
        pid = mesh.patch_id_per_bface()
        kind = mesh.patch_kind_per_bface()
        nI = mesh.n_internal
        P = mesh.owner[nI:]
        # p_b "as left by the last correctBoundaryConditions() before the closing one": the value for fixedValue patches, the cell
        # value for zeroGradient ones, and for qgdFlux patches the mid-step re-evaluation p_P - phiwStar / tauQGDf / |Sf| / deltaCoeffs
        # (qgdFluxFvPatchScalarField.C:184-192 + fixedGradient::evaluate) with the step's phiwStar and the old p_P
        phiw = o.get_face("phiwStar")[nI:]
        with np.errstate(divide="ignore", invalid="ignore"):
            p_mid = p0[P] - phiw / tauf0[nI:] / mesh.magSf[nI:] / mesh.deltaCoeffs[nI:]
        p0b = np.where(c.bcP[pid] == cases.FV, c.bvP, np.where(c.bcP[pid] == cases.QF, p_mid, p0[P]))
        p0b = np.where(kind == 1, 0.0, p0b)
        S = np.zeros((16, nC))
        S[6], S[12] = T1, cs
        S[13:16] = -7.0
        psiB = np.where(kind == 1, 0.0, 1.0 / (R * np.where(T1b == 0, 1.0, T1b)))
        aQ = np.full(nC, 0.5)
        sc, scb = sc0.copy(), sc0b.copy()
        mask = None
        if v["const_sc_cells"] is not None:
            m = np.zeros(nC, np.uint8)
            m[v["const_sc_cells"]] = 1
            mask = m.tobytes()
        prm = np.array([R, c.gas["Cp"], c.gas["mu"], c.gas["Pr"], c.gas["PrQGD"], 1.0, v["rC"], v["minSc"], v["maxSc"], c.gas["ScQGD"],
                        v["smoothCoeff"], v["badQualitySc"], v["maxAspectRatio"]])
        bMu, bAl, bSl = np.zeros(nB), np.zeros(nB), np.zeros(nB)
        L.hs_varsc5_correct(host.h, _d(prm), _d(S), _d(np.ascontiguousarray(T1b)), _d(np.ascontiguousarray(csb)), _d(psiB), _d(aQ),
                            _d(np.ascontiguousarray(p0)), _d(np.ascontiguousarray(p0b)), _d(sc), _d(scb), mask, _d(bMu), _d(bAl), _d(bSl), order)
        want, wantb = o.get("ScQGD", with_bnd=True)
        live = kind != 1
        assert np.abs(sc - want).max() < 1e-12
        assert np.abs(scb[live] - wantb[live]).max() < 1e-12
        mu, mub = o.get("mu", with_bnd=True)
        al, alb = o.get("alpha", with_bnd=True)
        assert np.abs(S[13] - mu).max() < 1e-13 * np.abs(mu).max()
        assert np.abs(S[14] - g * al).max() < 1e-13 * np.abs(g * al).max()
        assert np.abs(bMu[live] - mub[live]).max() < 1e-13 * np.abs(mub).max()
        assert np.abs(bAl[live] - g * alb[live]).max() < 1e-13 * np.abs(g * alb).max()
        assert np.array_equal(S[15], aQ) and np.array_equal(bSl[kind != 1], aQ[mesh.owner[mesh.n_internal:]][kind != 1])
        assert (bSl[kind == 1] == -1.0).all()          # empty patch faces are left alone
