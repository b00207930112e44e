"""The product's host-side set-up code (qgdsolver_b200/csrc/qgd_host_setup.cpp) checked on CPU against the oracle: the file is
compiled with g++ into a test-only shared library (tests/host_setup/host_setup_shim.cpp adds a C interface); the compact face
gradient record `G`, the point weights, the QGD length scales and the least-squares stencils it produces are applied to random
fields in numpy and compared with the oracle's operators.  No GPU, no device call."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
from qgdsolver_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
FF_POINTS, FF_TRI_QUIRK, FF_NORMAL_ONLY = 1, 2, 4


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(ROOT, "tests", "host_setup", "libhost_setup_shim.so")
    srcs = [os.path.join(ROOT, "tests", "host_setup", "host_setup_shim.cpp"), os.path.join(ROOT, "qgdsolver_b200", "csrc", "qgd_host_setup.cpp")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "qgdsolver_b200", "csrc"),
                               "-I", "/usr/local/cuda/include", "-o", out] + srcs + ["-L/usr/local/cuda/lib64", "-lcudart"])
    L = C.CDLL(out)
    L.hs_create.restype = C.c_void_p
    L.hs_create.argtypes = [C.POINTER(api._MeshDesc)]
    L.hs_destroy.argtypes = [C.c_void_p]
    L.hs_lengths.argtypes = [C.c_void_p, _dp, _dp]
    L.hs_face_records.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp]
    L.hs_point_interpolate.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.hs_lsq_width.argtypes = [C.c_void_p, C.c_int]
    L.hs_lsq.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_char_p]
    L.hs_cell_faces.argtypes = [C.c_void_p, _ip, _ip]
    L.hs_pcg_blocks.argtypes = [C.c_void_p, C.c_int, _ip]
    return L


class Host:
    def __init__(self, L, mesh, n_owned=0, coupled_face=None):
        self.L, self.mesh = L, mesh
        d = api._MeshDesc()
        d.n_owned_cells = int(n_owned)
        self._coupled = None if coupled_face is None else np.ascontiguousarray(coupled_face, np.int32)
        if self._coupled is not None:
            d.coupled_internal_face = self._coupled.ctypes.data_as(_ip)
        d.n_cells, d.n_faces, d.n_internal_faces, d.n_points, d.n_patches = mesh.n_cells, mesh.n_faces, mesh.n_internal, mesh.n_points, len(mesh.patches)
        f64 = lambda a: np.ascontiguousarray(a, np.float64)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        self.keep = dict(points=f64(mesh.points), C=f64(mesh.C), V=f64(mesh.V), Cf=f64(mesh.Cf), Sf=f64(mesh.Sf), magSf=f64(mesh.magSf),
                         weights=f64(mesh.weights), deltaCoeffs=f64(mesh.deltaCoeffs), nonOrthDeltaCoeffs=f64(mesh.nonOrthDeltaCoeffs),
                         neighb_cell_centres=f64(np.nan_to_num(mesh.neighb_cell_centres)))
        self.keepi = dict(face_offsets=i32(mesh.face_offsets), face_verts=i32(mesh.face_verts), owner=i32(mesh.owner), neighbour=i32(mesh.neighbour),
                          patch_start=i32([p.start for p in mesh.patches]), patch_size=i32([p.size for p in mesh.patches]),
                          patch_kind=i32([p.kind for p in mesh.patches]))
        for n, a in self.keep.items():
            setattr(d, n, a.ctypes.data_as(_dp))
        for n, a in self.keepi.items():
            setattr(d, n, a.ctypes.data_as(_ip))
        for j in range(3):
            d.geometric_d[j] = int(mesh.geometric_d[j])
        self.h = L.hs_create(C.byref(d))
        assert self.h, "HostMesh::build failed"

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hs_destroy(self.h)

    def records(self, reduced=False):
        m = self.mesh
        vtx, flags = np.zeros((m.n_faces, 4), np.int32), np.zeros(m.n_faces, np.int32)
        G, hd = np.zeros((9, m.n_faces)), np.zeros(max(m.n_bnd, 1))
        self.L.hs_face_records(self.h, int(reduced), vtx.ctypes.data_as(_ip), flags.ctypes.data_as(_ip), G.ctypes.data_as(_dp), hd.ctypes.data_as(_dp))
        return vtx, flags, G, hd[:m.n_bnd]

    def points(self, cell, bnd):
        out = np.zeros(self.mesh.n_points)
        self.L.hs_point_interpolate(self.h, np.ascontiguousarray(cell).ctypes.data_as(_dp), np.ascontiguousarray(bnd).ctypes.data_as(_dp), out.ctypes.data_as(_dp))
        return out


MESHES = {
    "hex_perturbed": lambda: cases.pm.hex_box(6, 5, 4, perturb=0.25, grading=(2, 1, 0.5), seed=11),
    "prism": lambda: cases.pm.prism_box(4, 4, 3, perturb=0.15, seed=2),
    "poly": lambda: cases.pm.hexprism_poly(5, 4, 3, a=0.1, lz=0.4),
    "truncoct": lambda: cases.pm.truncated_octahedron_box(4, 3, 3, h=0.3),
    "2d_z": lambda: cases.case_2d((9, 8), perturb=0.2).mesh,
    "2d_x": lambda: cases.case_2d((9, 8), perturb=0.2, axis=0).mesh,
    "1d": lambda: cases.case_sod(30).mesh,
    "forward_step": lambda: cases.pm.forward_step(10),
    "wedge": lambda: cases.pm.wedge_box(8, 6, angle_deg=6.0, perturb=0.15, seed=4),      # every vertex is a patch point, wedge faces get no derivative
}


def _apply_records(host, cell, bnd, bsg, reduced=False):
    """grad(phi)_f = G1 (phi_v1 - phi_v3) + G2 (phi_v2 - phi_v4) + GP (phi_P - phi_N) with the boundary ghost
    phi_N := phi_b + snGrad_b * halfDist (FF_NORMAL_ONLY faces: GP * snGrad_b): what k_fvsc_grad does with the records"""
    m = host.mesh
    nI = m.n_internal
    vtx, flags, G, hd = host.records(reduced)
    P = host.points(cell, bnd) if not reduced else np.zeros(m.n_points)
    d1 = np.where(flags & FF_POINTS, P[vtx[:, 0]] - P[vtx[:, 2]], 0.0)
    d2 = np.where(flags & FF_POINTS, P[vtx[:, 1]] - P[vtx[:, 3]], 0.0)
    dP = np.zeros(m.n_faces)
    dP[:nI] = cell[m.owner[:nI]] - cell[m.neighbour]
    normal_only = (flags[nI:] & FF_NORMAL_ONLY) != 0
    dP[nI:] = np.where(normal_only, bsg, cell[m.owner[nI:]] - (bnd + bsg * hd))
    d1[nI:] = np.where(normal_only, 0.0, d1[nI:])
    d2[nI:] = np.where(normal_only, 0.0, d2[nI:])
    return (G[0:3] * d1 + G[3:6] * d2 + G[6:9] * dP).T


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("reduced", [False, True])
def test_face_gradient_records_reproduce_the_oracle_operator(shim, oracle_mod, name, reduced):
    mesh = MESHES[name]()
    nI = mesh.n_internal
    rng = np.random.default_rng(5)
    cell = np.sin(3 * mesh.C[:, 0]) + 0.3 * rng.random(mesh.n_cells)
    bnd = np.cos(2 * mesh.Cf[nI:, 1]) + 0.3 * rng.random(mesh.n_bnd)
    bsg = mesh.deltaCoeffs[nI:] * (bnd - cell[mesh.owner[nI:]])
    bsg[1::2] = rng.random(mesh.n_bnd)[1::2]
    host = Host(shim, mesh)
    got = _apply_records(host, cell, bnd, bsg, reduced)
    ref = oracle_mod.Oracle(mesh).fvsc_grad(cell, bnd, bsg, scheme=oracle_mod.FVSC_SCHEMES["reduced" if reduced else "GaussVolPoint"])
    keep = np.ones(mesh.n_faces, bool)
    keep[nI:] = mesh.patch_kind_per_bface() != 1
    assert np.abs(got[keep] - ref[keep]).max() < 1e-11 * np.abs(ref[keep]).max()


@pytest.mark.parametrize("name", list(MESHES))
def test_point_weights_and_length_scales_match_the_oracle(shim, oracle_mod, name):
    mesh = MESHES[name]()
    host = Host(shim, mesh)
    o = oracle_mod.Oracle(mesh)
    rng = np.random.default_rng(6)
    cell, bnd = rng.random(mesh.n_cells), rng.random(mesh.n_bnd) + 3.0
    assert np.abs(host.points(cell, bnd) - o.vol_point_interpolate(cell, bnd)).max() < 1e-13
    hf, hc = np.zeros(mesh.n_faces), np.zeros(mesh.n_cells)
    shim.hs_lengths(host.h, hf.ctypes.data_as(_dp), hc.ctypes.data_as(_dp))
    keep = np.ones(mesh.n_faces, bool)
    keep[mesh.n_internal:] = mesh.patch_kind_per_bface() != 1
    assert np.abs(hf[keep] - o.hQGDf()[keep]).max() < 1e-14 and np.abs(hc - o.hQGD()).max() < 1e-14
    # cell -> face rows: ascending polyMesh face order, side bit = the cell is the face's neighbour
    n = shim.hs_cell_faces(host.h, None, None)
    off, enc = np.zeros(mesh.n_cells + 1, np.int32), np.zeros(n, np.int32)
    shim.hs_cell_faces(host.h, off.ctypes.data_as(_ip), enc.ctypes.data_as(_ip))
    ro, rf = mesh.cell_faces_csr()
    assert np.array_equal(off, ro) and np.array_equal(enc >> 1, rf)
    cells_of_rows = np.repeat(np.arange(mesh.n_cells), np.diff(off))
    assert np.array_equal((enc & 1) == 1, mesh.owner[enc >> 1] != cells_of_rows)


@pytest.mark.parametrize("name", ["2d_z", "2d_x", "1d"])
@pytest.mark.parametrize("opt", [False, True])
def test_least_squares_stencils_reproduce_the_oracle_operator(shim, oracle_mod, name, opt):
    mesh = MESHES[name]() if name != "2d_z" else cases.case_2d((6, 40)).mesh        # aspect ratio ~7: degenerate stencils
    nI = mesh.n_internal
    host = Host(shim, mesh)
    W = shim.hs_lsq_width(host.h, int(opt))
    cells, coef, deg = np.zeros((W, nI), np.int32), np.zeros((W, 3, nI)), np.zeros(nI, np.int8)
    shim.hs_lsq(host.h, int(opt), cells.ctypes.data_as(_ip), coef.ctypes.data_as(_dp), deg.ctypes.data_as(C.c_char_p))
    rng = np.random.default_rng(7)
    cell, bnd = rng.random(mesh.n_cells), rng.random(mesh.n_bnd)
    bsg = mesh.deltaCoeffs[nI:] * (bnd - cell[mesh.owner[nI:]])
    o = oracle_mod.Oracle(mesh)
    sF = o.linear_interpolate(cell, bnd)[:nI]
    lsq = np.einsum("wif,wf->fi", coef, cell[cells] - sF[None])
    nf = mesh.Sf[:nI] / mesh.magSf[:nI, None]
    fallback = (mesh.nonOrthDeltaCoeffs[:nI] * (cell[mesh.neighbour] - cell[mesh.owner[:nI]]))[:, None] * nf
    got = np.where((deg != 0)[:, None], fallback, lsq)
    ref = o.fvsc_grad(cell, bnd, bsg, scheme=oracle_mod.FVSC_SCHEMES["leastSquaresOpt" if opt else "leastSquares"])[:nI]
    assert np.abs(got - ref).max() < 1e-10 * np.abs(ref).max()
    if name == "2d_z" and not opt:
        assert deg.any()


def test_face_flags_mark_the_reference_quirks(shim):
    """FF_TRI_QUIRK exactly on the internal triangular faces of a 3D mesh (GaussVolPointBase3D.C:844-854), FF_POINTS on tri / quad
    faces only (polygons take nf*snGrad), FF_NORMAL_ONLY on boundary polygons and on every boundary face of `reduced`."""
    m = cases.pm.prism_box(4, 3, 3, perturb=0.1)
    nI = m.n_internal
    _, flags, _, _ = Host(shim, m).records()
    nv = m.face_nverts()
    assert np.array_equal((flags[:nI] & FF_TRI_QUIRK) != 0, nv[:nI] == 3) and not (flags[nI:] & FF_TRI_QUIRK).any()
    assert ((flags & FF_POINTS) != 0).all()
    t = cases.pm.truncated_octahedron_box(3, 3, 3)
    _, flags, G, _ = Host(shim, t).records()
    nv = t.face_nverts()
    assert np.array_equal((flags & FF_POINTS) != 0, nv == 4)
    assert np.array_equal((flags[t.n_internal:] & FF_NORMAL_ONLY) != 0, nv[t.n_internal:] > 4)
    assert np.abs(G[0:6][:, nv > 4]).max() == 0.0                     # G1 = G2 = 0 on polygons
    _, flags, _, _ = Host(shim, t).records(reduced=True)
    assert ((flags[t.n_internal:] & FF_NORMAL_ONLY) != 0).all() and not (flags & FF_POINTS).any()


def test_pcg_blocks_are_compact_tiles_and_precondition_like_dic(shim, oracle_mod):
    """HostMesh::makePcgBlocks (recursive coordinate bisection): every cell in exactly one block, block sizes within
    (target/2, target], tiles compact (a 64 x 64 mesh cut into 256-cell tiles: 16 x 16 squares); with these blocks the oracle's
    block-local DIC (or_pcg_solve_blocks = what `mpirun -np N` does) needs far fewer iterations than the diagonal preconditioner."""
    import cases
    n = 64
    mesh = cases.pm.hex_box(n, n, 1, lengths=(1.0, 1.0, 0.1), patch_kinds={"zMin": "empty", "zMax": "empty"})
    host = Host(shim, mesh)
    blk = np.empty(mesh.n_cells, np.int32)
    nb = shim.hs_pcg_blocks(host.h, 256, blk.ctypes.data_as(_ip))
    assert nb == 16 and blk.min() == 0 and blk.max() == nb - 1
    cnt = np.bincount(blk)
    assert cnt.max() <= 256 and cnt.min() > 128
    for b in range(nb):                                  # 16 x 16 squares
        c = mesh.C[blk == b]
        assert np.ptp(c[:, 0]) < 16.0 / n and np.ptp(c[:, 1]) < 16.0 / n
    nb2 = shim.hs_pcg_blocks(host.h, 100, blk.ctypes.data_as(_ip))
    cnt = np.bincount(blk)
    assert nb2 == 64 and cnt.max() <= 100 and cnt.min() >= 50
    nI = mesh.n_internal
    upper = -(mesh.magSf[:nI] * mesh.deltaCoeffs[:nI])
    diag = np.zeros(mesh.n_cells)
    np.subtract.at(diag, mesh.owner[:nI], upper); np.subtract.at(diag, mesh.neighbour, upper)
    diag[0] += diag[0]
    b = np.random.default_rng(0).standard_normal(mesh.n_cells)
    o = oracle_mod.Oracle(mesh)
    its = {}
    for name, pc, cb in (("diagonal", 1, None), ("DIC", 2, None), ("blockDIC", 2, blk)):
        _, its[name], _, res = o.pcg_solve(diag, upper, b, np.zeros(mesh.n_cells), tol=1e-8, maxIter=5000, precond=pc, cell_block=cb)
        assert res < 1e-8
    assert its["DIC"] <= its["blockDIC"] < 0.55 * its["diagonal"]


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("reduced", [False, True])
def test_cell_difference_vector_of_every_record_is_parallel_to_sf(shim, name, reduced):
    """The device keeps GP as one scalar times Sf (7 doubles per face instead of 9): in every fvsc scheme and on every face type
    GP is parallel to the face area vector - quad: (e1 x e2)/D with the two diagonals (half their cross product is the vector area
    of ANY quadrilateral, planar or warped); tri: A_f / (3 vt); 2D: the in-plane normal of the edge v13; 1D / reduced / polygon
    faces: -nf delta.  Also checked on a strongly warped hex mesh."""
    import cases
    meshes = [MESHES[name]()]
    if name == "hex_perturbed":
        meshes.append(cases.pm.hex_box(6, 5, 4, perturb=0.35, grading=(3, 1, 0.3), seed=23))
    for mesh in meshes:
        _, flags, G, _ = Host(shim, mesh).records(reduced)
        keep = np.ones(mesh.n_faces, bool)
        kind = mesh.patch_kind_per_bface()
        keep[mesh.n_internal:] = (kind != 1) & (kind != 3)          # empty faces carry nothing, wedge faces get no derivative (record = 0)
        if not reduced:     # `reduced` keeps nf*snGrad on wedge faces (reducedFaceNormalStencil.C:69-108), GaussVolPoint nothing (GaussVolPointBase2D.C:175-179)
            assert np.abs(G[:, mesh.n_internal:][:, kind == 3]).max(initial=0.0) == 0.0
        GP, S = G[6:9].T[keep], mesh.Sf[keep]
        s = (GP * S).sum(1) / (S * S).sum(1)
        assert np.abs(GP - s[:, None] * S).max() <= 1e-12 * np.abs(GP).max()
        assert np.abs(s).min() > 0


@pytest.mark.parametrize("mesh_fn,parts", [
    (lambda: cases.pm.hex_box(8, 6, 5, perturb=0.25, grading=(2, 1, 0.5), seed=11), 2),
    (lambda: cases.pm.hex_box(8, 8, 6, perturb=0.2, seed=4), 8),
    (lambda: cases.pm.prism_box(4, 4, 3, perturb=0.15, seed=2), 4),
    (lambda: cases.case_2d((12, 10), perturb=0.2).mesh, 4),
])
def test_decomposed_length_scales_match_the_n_subdomain_oracle(shim, oracle_mod, mesh_fn, parts):
    """north_star: an N-GPU run must equal the N-SUBDOMAIN reference, whose QGD length scales differ from the serial run on
    non-uniform meshes: hQGDf = 1/deltaCoeffs = |C_N - C_P| on processor-patch faces (QGDCoeffs.C:195-199) instead of
    2 min(|C_P - C_f|, |C_N - C_f|) (:303-308), and hQGD (:320-362) averages those.  This is the ONLY decomposition-dependent
    quantity of the path and it is host set-up code: the product's HostMesh::build on every rank's extended sub-mesh (coupled-face
    flags) against the oracle run on the decomposePar-layout processor mesh of the same rank (processor patches, coupled geometry)."""
    from qgdsolver_b200 import decompose
    mesh = mesh_fn()
    rank = decompose.geometric_split(mesh, parts)
    subs = decompose.extended_submeshes(mesh, rank)
    procs = decompose.processor_meshes(mesh, rank)
    decompose.couple_processor_geometry(procs)
    serial = oracle_mod.Oracle(mesh)
    hf_serial = serial.hQGDf()
    differs = 0.0
    for sd, pr in zip(subs, procs):
        host = Host(shim, sd.mesh, n_owned=sd.n_owned, coupled_face=sd.coupled_face)
        hf, hc = np.zeros(sd.mesh.n_faces), np.zeros(sd.mesh.n_cells)
        shim.hs_lengths(host.h, hf.ctypes.data_as(_dp), hc.ctypes.data_as(_dp))
        o = oracle_mod.Oracle(pr.mesh)
        ohf, ohc = o.hQGDf(), o.hQGD()
        # cells: owned cells of the sub-mesh are the processor mesh's cells, in cellProcAddressing order
        assert np.array_equal(sd.cell_global[:sd.n_owned], pr.cell_addr)
        assert np.abs(hc[:sd.n_owned] - ohc).max() < 1e-14 * np.abs(ohc).max()
        # faces: compare through the global face id (processor-mesh faces incl. its processor-patch faces)
        g_proc = np.abs(pr.face_addr.astype(np.int64)) - 1
        where = {int(g): i for i, g in enumerate(sd.face_global)}
        kinds = np.full(pr.mesh.n_faces, 0)
        for ptc in pr.mesh.patches:
            kinds[ptc.start:ptc.start + ptc.size] = ptc.kind
        idx = np.array([where[int(g)] for g in g_proc])
        keep = kinds != cases.pm.PATCH_EMPTY
        assert np.abs(hf[idx][keep] - ohf[keep]).max() < 1e-14 * np.abs(ohf[keep]).max()
        # the coupled faces really follow the processor-patch rule, and it differs from the serial value on this mesh
        proc_faces = kinds == cases.pm.PATCH_PROCESSOR
        if proc_faces.any():
            d = 1.0 / pr.mesh.deltaCoeffs[proc_faces]
            assert np.abs(hf[idx][proc_faces] - d).max() < 1e-14 * d.max()
            differs = max(differs, float(np.abs(hf[idx][proc_faces] - hf_serial[g_proc[proc_faces]]).max()))
    assert differs > 1e-4          # non-uniform mesh: the decomposed run is NOT the serial run
