"""Known-answer tests for the oracle's varScModel5 (varScModel5.C:52-269): the FaceCellWave restatement of fvc::smooth, the
mesh-quality floor from primitiveMeshTools::cellClosedness, and one correct() against an independent numpy restatement.
Like every [OF-v2312] item the OpenFOAM semantics are restated from the OpenFOAM sources as remembered (parity unpinned)."""
import numpy as np
import pytest

import cases


# ---------------------------------------------------------------- an independent, literal FaceCellWave<smoothData> in Python
class _SmoothData:
    """smoothData.H / smoothDataI.H"""
    __slots__ = ("v",)

    def __init__(self, v=-1.0e15):      # value_(-GREAT)
        self.v = v

    def valid(self):
        return self.v > -1.0e-15        # value_ > -SMALL

    def update(self, other, scale, tol):
        if (not self.valid()) or self.v < 1.0e-300:      # VSMALL
            self.v = other.v
            return True
        if other.v > (1 + tol) * scale * self.v:
            self.v = other.v / scale
            return True
        return False


def _mesh_cells(mesh):
    """primitiveMesh::calcCells: the faces of each cell, owner loop first, neighbour loop second"""
    cells = [[] for _ in range(mesh.n_cells)]
    for f in range(mesh.n_faces):
        cells[mesh.owner[f]].append(f)
    for f in range(mesh.n_internal):
        cells[mesh.neighbour[f]].append(f)
    return cells


def py_fvc_smooth(mesh, field, coeff):
    """smooth.C: fvc::smooth(field, coeff) through FaceCellWave (faceToCell / cellToFace with changed lists and bit sets)"""
    max_ratio, tol = 1.0 + coeff, 0.01
    nI = mesh.n_internal
    cell_info = [_SmoothData(float(v)) for v in field]
    face_info = [_SmoothData() for _ in range(mesh.n_faces)]
    changed_faces, changed_face = [], np.zeros(mesh.n_faces, bool)
    changed_cells, changed_cell = [], np.zeros(mesh.n_cells, bool)
    for f in range(nI):
        o, n = mesh.owner[f], mesh.neighbour[f]
        if field[o] > max_ratio * field[n]:
            info = _SmoothData(float(field[o]))
        elif field[n] > max_ratio * field[o]:
            info = _SmoothData(float(field[n]))
        else:
            continue
        face_info[f] = info             # setFaceInfo
        changed_face[f] = True
        changed_faces.append(f)
    cells = _mesh_cells(mesh)
    iters = 0
    while True:
        # faceToCell
        for f in changed_faces:
            assert changed_face[f]
            new = face_info[f]
            for c in ([mesh.owner[f], mesh.neighbour[f]] if f < nI else [mesh.owner[f]]):
                cur = cell_info[c]
                if cur.v != new.v:      # !equal
                    if cur.update(new, max_ratio, tol) and not changed_cell[c]:
                        changed_cell[c] = True
                        changed_cells.append(c)
            changed_face[f] = False
        changed_faces = []
        if not changed_cells:
            break
        # cellToFace
        for c in changed_cells:
            new = cell_info[c]
            for f in cells[c]:
                cur = face_info[f]
                if cur.v != new.v:
                    if cur.update(new, 1.0, tol) and not changed_face[f]:
                        changed_face[f] = True
                        changed_faces.append(f)
            changed_cell[c] = False
        changed_cells = []
        if not changed_faces:
            break
        iters += 1
    return np.array([ci.v for ci in cell_info]), iters


MESHES = {
    "hex": lambda: cases.pm.hex_box(7, 6, 5, perturb=0.2, seed=2),
    "prism": lambda: cases.pm.prism_box(5, 4, 3, perturb=0.1, seed=3),
    "truncoct": lambda: cases.pm.truncated_octahedron_box(4, 3, 3, h=0.25),
    "2d": lambda: cases.case_2d((14, 11), perturb=0.2).mesh,
    "1d": lambda: cases.case_sod(40).mesh,
}


@pytest.mark.parametrize("name", sorted(MESHES))
@pytest.mark.parametrize("coeff", [0.1, 0.5])
def test_fvc_smooth_equals_the_literal_facecellwave_restatement(oracle_mod, name, coeff):
    """random fields with a few spikes (many near-ties inside the 1 % tolerance): bit-equal values and iteration counts"""
    mesh = MESHES[name]()
    o = oracle_mod.Oracle(mesh)
    rng = np.random.default_rng(11)
    for trial in range(3):
        f = 0.05 + 0.02 * rng.random(mesh.n_cells)
        spikes = rng.choice(mesh.n_cells, size=max(2, mesh.n_cells // 25), replace=False)
        f[spikes] = 0.3 + 0.7 * rng.random(spikes.size)
        if trial == 2:
            f[rng.choice(mesh.n_cells, size=3, replace=False)] = 0.0      # "value < VSMALL": the cell copies the face value
        got, it = o.fvc_smooth(f, coeff)
        want, it_py = py_fvc_smooth(mesh, f, coeff)
        assert np.array_equal(got, want)
        assert it == it_py and it > 0


def test_fvc_smooth_closed_form_on_a_chain(oracle_mod):
    """1D chain, one spike v in a floor s: neighbour j cells away is raised to v/maxRatio^j while the offered face value
    exceeds (1 + tol) maxRatio times its own value; everything else keeps s"""
    mesh = cases.case_sod(60).mesh
    o = oracle_mod.Oracle(mesh)
    s, v, coeff = 0.05, 1.0, 0.25
    r = 1.0 + coeff
    f = np.full(mesh.n_cells, s)
    k = 23
    f[k] = v
    got, it = o.fvc_smooth(f, coeff)
    want = f.copy()
    for sgn in (-1, 1):
        val = v
        for j in range(1, 40):
            if val > 1.01 * r * s:
                val = val / r
                want[k + sgn * j] = val
            else:
                break
    assert np.array_equal(got, want)
    assert it >= 10
    # a field that is already smooth is returned untouched without a single sweep
    g = s * r ** (0.9 * (np.arange(mesh.n_cells) % 2))
    got2, it2 = o.fvc_smooth(g, coeff)
    assert np.array_equal(got2, g) and it2 == 0


@pytest.mark.parametrize("name", ["hex", "truncoct", "2d"])
def test_fvc_smooth_invariants(oracle_mod, name):
    mesh = MESHES[name]()
    o = oracle_mod.Oracle(mesh)
    rng = np.random.default_rng(5)
    f = 0.05 + rng.random(mesh.n_cells) ** 6
    coeff = 0.1
    got, _ = o.fvc_smooth(f, coeff)
    assert (got >= f).all()                                   # values are only ever raised
    assert got.max() == f.max()
    a, b = got[mesh.owner[:mesh.n_internal]], got[mesh.neighbour]
    ratio = np.maximum(a, b) / np.minimum(a, b)
    # a face offers the larger value; the smaller cell is raised unless within (1 + tol) maxRatio of it (the face value may
    # itself lag the cell by the same 1 %)
    assert ratio.max() <= (1 + coeff) * 1.01 * 1.01 + 1e-12
    again, it = o.fvc_smooth(got, coeff)
    assert np.array_equal(again, got)                         # idempotent


def test_cell_aspect_ratio_closed_forms(oracle_mod):
    """primitiveMeshTools::cellClosedness on boxes a x b x c: sums of |Sf| components (2bc, 2ac, 2ab)"""
    a, b, c = 1.0 / 4, 2.0 / 5, 3.0 / 2
    mesh = cases.pm.hex_box(4, 5, 2, lengths=(1.0, 2.0, 3.0))
    o = oracle_mod.Oracle(mesh)
    q, ar = o.varsc5_cell_quality(0.05, 1.5)
    sums = np.array([2 * b * c, 2 * a * c, 2 * a * b])
    want = max(sums.max() / sums.min(), sums.sum() / 6.0 / (a * b * c) ** (2.0 / 3.0))
    assert np.allclose(ar, want, rtol=1e-13)
    assert np.allclose(q, 0.05 * want / 1.5, rtol=1e-13)
    cube = cases.pm.hex_box(3, 3, 3)
    q, ar = oracle_mod.Oracle(cube).varsc5_cell_quality(0.05, 1.5)
    assert np.allclose(ar, 1.0, rtol=1e-13) and (q == 0).all()
    # 2D (empty z): only the solved directions count, no hydraulic term
    m2 = cases.pm.hex_box(6, 3, 1, lengths=(1.0, 1.0, 0.1), patch_kinds={"zMin": "empty", "zMax": "empty"})
    q, ar = oracle_mod.Oracle(m2).varsc5_cell_quality(0.05, 1.5)
    dx, dy, dz = 1.0 / 6, 1.0 / 3, 0.1
    assert np.allclose(ar, (2 * dy * dz) / (2 * dx * dz), rtol=1e-13)      # = 2 > 1.5
    assert np.allclose(q, 0.05 * 2.0 / 1.5, rtol=1e-13)


def _gauss_grad_scalar(mesh, cell, bnd):
    """[OF-v2312] fvc::grad, Gauss linear: cell values and the corrected boundary values"""
    nI = mesh.n_internal
    o, n = mesh.owner[:nI], mesh.neighbour
    w = mesh.weights[:nI]
    ff = np.concatenate([w * (cell[o] - cell[n]) + cell[n], bnd])
    kind = mesh.patch_kind_per_bface()
    flux = mesh.Sf * ff[:, None]
    flux[nI:][kind == 1] = 0.0
    g = np.zeros((mesh.n_cells, 3))
    np.add.at(g, mesh.owner, flux)
    np.subtract.at(g, n, flux[:nI])
    g /= mesh.V[:, None]
    P = mesh.owner[nI:]
    nf = mesh.Sf[nI:] / mesh.magSf[nI:, None]
    sn = mesh.deltaCoeffs[nI:] * (bnd - cell[P])
    gb = g[P] + nf * (sn - np.einsum("bi,bi->b", nf, g[P]))[:, None]
    return g, gb


@pytest.mark.parametrize("bcs", ["zg", "fixed"])
def test_varsc5_correct_matches_a_numpy_restatement(oracle_mod, bcs):
    """one QGDFoam step with varScModel5: ScQGD (cells and boundary), then mu = mu_mol + p_old ScQGD tauQGD, from the fields the
    oracle reports before and after the step"""
    c = cases.case_hex3d(n=(7, 6, 5), perturb=0.2, bcs=bcs, model="varScModel5",
                         varsc=dict(rC=0.35, smoothCoeff=0.15, maxAspectRatio=1.2))
    mesh = c.mesh
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 3)
    R = c.gas["R"]
    sc0, sc0b = o.get("ScQGD", with_bnd=True)
    p0, p0b = o.get("p", with_bnd=True)
    c.oracle_step(o, 1)
    T1, T1b = o.get("T", with_bnd=True)
    rho, rhob = p0 * (1.0 / (R * T1)), p0b * (1.0 / (R * T1b))
    g, gb = _gauss_grad_scalar(mesh, rho, rhob)
    h, hf = o.hQGD(), o.hQGDf()
    v = c.varsc
    sc = v["rC"] * (np.linalg.norm(g, axis=1) * h / rho) + (1 - v["rC"]) * sc0
    scb = v["rC"] * (np.linalg.norm(gb, axis=1) * hf[mesh.n_internal:] / rhob) + (1 - v["rC"]) * sc0b
    sc, scb = np.clip(sc, v["minSc"], v["maxSc"]), np.clip(scb, v["minSc"], v["maxSc"])
    q, ar = o.varsc5_cell_quality(v["badQualitySc"], v["maxAspectRatio"])
    assert (q > 0).any() and (q == 0).any()
    sc = np.maximum(sc, q)
    sc_s, it = py_fvc_smooth(mesh, sc, v["smoothCoeff"])
    got, gotb = o.get("ScQGD", with_bnd=True)
    # the gradient sums differ in rounding from the oracle's loop order: compare the unsmoothed values where smoothing left them
    untouched = sc_s == sc
    assert untouched.sum() > mesh.n_cells // 8
    assert np.abs(got[untouched] - sc[untouched]).max() < 1e-13
    assert np.abs(got - sc_s).max() < 1e-12
    assert np.abs(gotb - scb).max() < 1e-13
    assert got.min() >= v["minSc"] and got.max() <= max(v["maxSc"], q.max())
    # QGD viscosity with the smoothed ScQGD and the OLD pressure (varScModel5.C:244-253), tauQGD = alphaQGD hQGD / c (:207)
    cs = o.get("c")
    tau = 0.5 * h / cs
    assert np.abs(o.get("tauQGD") - tau).max() < 1e-15
    mu = c.gas["mu"] + p0 * got * tau
    assert np.abs(o.get("mu") - mu).max() < 1e-15
    # tauQGDf = I(alphaQGD) / I(c) * hQGDf (:204-205)
    nI = mesh.n_internal
    w = mesh.weights[:nI]
    cf = w * (cs[mesh.owner[:nI]] - cs[mesh.neighbour]) + cs[mesh.neighbour]
    assert np.abs(o.get_face("tauQGDf")[:nI] - 0.5 / cf * hf[:nI]).max() < 1e-15


def test_varsc5_const_sc_cell_set_and_relaxation_memory(oracle_mod):
    """cells of constScCellSet are reset to the dictionary ScQGD before smoothing (varScModel5.C:222-230, :137); with rC = 0 the
    sensor is switched off and ScQGD only ever grows by smoothing from those cells"""
    cells = np.array([5, 17, 40], np.int32)
    c = cases.case_hex3d(n=(6, 5, 4), bcs="zg", model="varScModel5", gas=dict(cases.GAS, ScQGD=0.8),
                         varsc=dict(rC=0.0, const_sc_cells=cells, minSc=0.05, maxSc=1.0, smoothCoeff=0.2))
    o = c.make_oracle(oracle_mod)
    c.oracle_step(o, 2)
    sc = o.get("ScQGD")
    assert np.allclose(sc[cells], 0.8)
    # rC = 0: the field starts at the dictionary value everywhere (0.8) and stays: nothing to smooth, nothing to relax
    assert np.allclose(sc, 0.8)
    c2 = cases.case_hex3d(n=(6, 5, 4), bcs="zg", model="varScModel5", gas=dict(cases.GAS, ScQGD=0.8),
                          varsc=dict(rC=1.0, const_sc_cells=cells, minSc=0.05, maxSc=1.0, smoothCoeff=0.2))
    o2 = c2.make_oracle(oracle_mod)
    c2.oracle_step(o2, 2)
    sc2 = o2.get("ScQGD")
    assert np.allclose(sc2[cells], 0.8)
    mesh = c2.mesh
    nb = np.concatenate([mesh.neighbour[mesh.owner[:mesh.n_internal] == 17], mesh.owner[:mesh.n_internal][mesh.neighbour == 17]])
    assert (sc2[nb] >= 0.8 / 1.2 / 1.0101).all() and sc2.min() < 0.3      # the spike is spread geometrically, far cells stay low
