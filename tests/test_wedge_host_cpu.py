"""Wedge patches, product side, without a GPU: the vertex list / normals the host set-up hands to k_wedge_points and the face
tensor of the wedge velocity condition (qgdsolver_b200/csrc/qgd_wedge.h, compiled with g++) against the oracle."""
import ctypes as C

import numpy as np

import cases
from qgdsolver_b200 import foamcase, polymesh as pm
from test_host_setup_cpu import Host, shim  # noqa: F401  (fixture)

_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def test_wedge_vertex_list_reproduces_the_oracle_point_constraint(shim, oracle_mod):
    L = shim
    L.hs_wedge_points.restype = C.c_int
    L.hs_wedge_points.argtypes = [C.c_void_p, _ip, _dp]
    m = pm.wedge_box(7, 5, angle_deg=7.0, perturb=0.1, seed=3)
    host = Host(L, m)
    n = L.hs_wedge_points(host.h, None, None)
    assert n == m.n_points                                   # every vertex of a wedge block lies on one of the two wedge patches
    pts, R = np.zeros(n, np.int32), np.zeros((n, 3, 3))
    L.hs_wedge_points(host.h, pts.ctypes.data_as(_ip), R.ctypes.data_as(_dp))
    assert sorted(pts) == list(range(m.n_points))
    h = np.deg2rad(7.0) / 2
    for sgn in (1.0, -1.0):                                  # one plane per vertex: R = I - n n with the wedge patch's normal
        nn = np.array([0.0, -np.sin(h), sgn * np.cos(h)])
        sel = (m.points[pts, 2] > 0) == (sgn > 0)
        assert np.abs(R[sel] - (np.eye(3) - np.outer(nn, nn))).max() < 1e-15
    # what k_patch_points + k_wedge_points compute for a vector field: boundary interpolation, then R . v
    rng = np.random.default_rng(8)
    cell, bnd = rng.random((m.n_cells, 3)), rng.random((m.n_bnd, 3))
    raw = np.stack([host.points(cell[:, j].copy(), bnd[:, j].copy()) for j in range(3)], 1)
    got = raw.copy()
    got[pts] = np.einsum("pab,pb->pa", R, raw[pts])
    want = oracle_mod.Oracle(m).vol_point_interpolate(cell, bnd)
    assert np.abs(got - want).max() < 1e-14
    # a mesh without wedge patches has no such vertices
    assert L.hs_wedge_points(Host(L, cases.case_2d((5, 4)).mesh).h, None, None) == 0


def test_wedge_face_tensor_matches_the_oracle_boundary_velocity(shim, oracle_mod):
    L = shim
    L.hs_wedge_face_t.argtypes = [_dp, _dp]
    c = cases.case_wedge(n=(6, 5), angle_deg=9.0)
    c.U0 = np.random.default_rng(5).random((c.mesh.n_cells, 3)) - 0.5
    o = c.make_oracle(oracle_mod)
    m = c.mesh
    U, Ub = o.get("U", with_bnd=True)
    nI = m.n_internal
    for p in m.patches:
        if p.kind != pm.PATCH_WEDGE:
            continue
        for f in range(p.start, p.start + p.size):
            n = np.ascontiguousarray(m.Sf[f] / m.magSf[f])
            T = np.zeros(9)
            L.hs_wedge_face_t(n.ctypes.data_as(_dp), T.ctypes.data_as(_dp))
            T = T.reshape(3, 3)
            assert np.abs(T @ T.T - np.eye(3)).max() < 1e-15 and abs(np.linalg.det(T) - 1) < 1e-15     # a rotation
            assert np.abs(Ub[f - nI] - T @ U[m.owner[f]]).max() < 1e-15
    # degenerate input: the patch normal is a coordinate axis -> identity
    T = np.zeros(9)
    L.hs_wedge_face_t(np.array([0.0, 0.0, 1.0]).ctypes.data_as(_dp), T.ctypes.data_as(_dp))
    assert np.array_equal(T.reshape(3, 3), np.eye(3))


def test_wedge_case_round_trip_through_the_case_reader(tmp_path):
    """polyMesh with wedge patches and 0/U with `type wedge;` -> patch kinds, geometricD and the device BC codes"""
    m = pm.wedge_box(4, 3)
    case = str(tmp_path)
    foamcase.write_polymesh(m, case)
    m2 = foamcase.read_polymesh(case)
    assert [p.kind for p in m2.patches] == [p.kind for p in m.patches]
    assert list(m2.geometric_d) == [1, 1, -1]
    U = foamcase.VolField("U", 3, np.zeros((m.n_cells, 3)), {p.name: ("wedge" if p.kind == pm.PATCH_WEDGE else "zeroGradient") for p in m.patches}, {}, {})
    T = foamcase.VolField("T", 1, np.ones(m.n_cells), dict(U.patch_types), {}, {})
    kU, _ = foamcase.bc_arrays(m2, U)
    kT, _ = foamcase.bc_arrays(m2, T)
    wedge = np.array([p.kind == pm.PATCH_WEDGE for p in m.patches])
    assert (kU[wedge] == 7).all() and (kU[~wedge] == 1).all() and (kT == 1).all()


def _symmetry_plane_box():
    """hex box whose yMin / yMax / xMax patches are of polyPatch type symmetryPlane (zMin / zMax ordinary walls, xMin an inlet)"""
    m = pm.hex_box(5, 4, 3, perturb=0.15, seed=6)
    m.patches = [pm.Patch(p.name, pm.PATCH_SYMMETRY_PLANE if p.name in ("yMin", "yMax", "xMax") else p.kind, p.start, p.size) for p in m.patches]
    return m


def test_symmetry_plane_vertices_carry_the_combined_point_constraint(shim, oracle_mod):
    """[OF-v2312 pointConstraints]: vertices inside a symmetryPlane patch keep I - n n, vertices on the edge where two such patches
    meet are confined to the line along n1 x n2, vertices of ordinary patches are free; product host set-up == oracle"""
    L = shim
    L.hs_wedge_points.restype = C.c_int
    L.hs_wedge_points.argtypes = [C.c_void_p, _ip, _dp]
    m = _symmetry_plane_box()
    host = Host(L, m)
    n = L.hs_wedge_points(host.h, None, None)
    pts, R = np.zeros(n, np.int32), np.zeros((n, 3, 3))
    L.hs_wedge_points(host.h, pts.ctypes.data_as(_ip), R.ctypes.data_as(_dp))
    x, y = m.points[pts, 0], m.points[pts, 1]
    on_y = (np.abs(y) < 1e-12) | (np.abs(y - 1) < 1e-12)
    on_x = np.abs(x - 1) < 1e-12
    assert (on_y | on_x).all() and n == int(((np.abs(m.points[:, 1]) < 1e-12) | (np.abs(m.points[:, 1] - 1) < 1e-12) | (np.abs(m.points[:, 0] - 1) < 1e-12)).sum())
    ey, ex, ez = np.diag([0.0, 1.0, 0.0]), np.diag([1.0, 0.0, 0.0]), np.diag([0.0, 0.0, 1.0])
    assert np.abs(R[on_y & ~on_x] - (np.eye(3) - ey)).max() < 1e-15
    assert np.abs(R[on_x & ~on_y] - (np.eye(3) - ex)).max() < 1e-15
    assert np.abs(R[on_x & on_y] - ez).max() < 1e-15 and (on_x & on_y).sum() == 2 * 4
    rng = np.random.default_rng(3)
    cell, bnd = rng.random((m.n_cells, 3)), rng.random((m.n_bnd, 3))
    raw = np.stack([host.points(cell[:, j].copy(), bnd[:, j].copy()) for j in range(3)], 1)
    got = raw.copy()
    got[pts] = np.einsum("pab,pb->pa", R, raw[pts])
    want = oracle_mod.Oracle(m).vol_point_interpolate(cell, bnd)
    assert np.abs(got - want).max() < 1e-14
    # scalars are untouched
    assert np.abs(host.points(cell[:, 0].copy(), bnd[:, 0].copy()) - oracle_mod.Oracle(m).vol_point_interpolate(cell[:, 0].copy(), bnd[:, 0].copy())).max() < 1e-14
