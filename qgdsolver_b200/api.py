"""ctypes binding of libqgd_b200.so (include/qgd_b200.h) for tests and bench.py.

This is harness glue, not the product: the product is the C-ABI library plus the C++ host mirror under
qgdsolver_b200/host/.  Names follow the reference: fvsc::fvscStencil (fvscStencil.H:46-137),
fvsc::grad / fvsc::div (fvsc.H:46-68), QGDFoam (QGDFoam.C:68-168).
There is no CPU fallback: every compute call raises QGDError when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .polymesh import PolyMesh

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqgd_b200.so")

BC_FIXED_VALUE, BC_ZERO_GRADIENT, BC_FIXED_GRADIENT, BC_QGD_FLUX, BC_CALCULATED, BC_QHD_FLUX, BC_SLIP, BC_WEDGE = 0, 1, 2, 3, 4, 5, 6, 7
QGD_OK, ERR_INVALID, ERR_UNKNOWN_MODEL, ERR_UNSUPPORTED, ERR_CUDA, ERR_COMM, ERR_STATE = 0, -1, -2, -3, -4, -5, -6

# every symbol include/qgd_b200.h declares (checked by tests/test_abi_cpu.py)
ABI_SYMBOLS = [
    "qgd_init", "qgd_last_error", "qgd_version", "qgd_device_synchronize",
    "qgd_mesh_create", "qgd_mesh_destroy", "qgd_mesh_get", "qgd_mesh_set_degenerate_stencil_faces",
    "qgd_mesh_set_pcg_blocks", "qgd_mesh_make_pcg_blocks", "qgd_mesh_get_pcg_blocks",
    "qgd_fvsc_create", "qgd_fvsc_destroy", "qgd_fvsc_grad", "qgd_fvsc_div",
    "qgd_qgdfoam_create", "qgd_qgdfoam_destroy", "qgd_qgdfoam_set_bcs", "qgd_qgdfoam_init_fields",
    "qgd_qgdfoam_set_const_sc_cells", "qgd_qgdfoam_set_sources", "qgd_qgdfoam_step", "qgd_qgdfoam_step_host", "qgd_qgdfoam_step_fields_host", "qgd_qgdfoam_face_kernel", "qgd_qgdfoam_get", "qgd_qgdfoam_get_flux",
    "qgd_qgdfoam_get_scalars", "qgd_qgdfoam_state_guard", "qgd_qgdfoam_launch_count", "qgd_qgdfoam_profile", "qgd_qgdfoam_kernel_times",
    "qgd_qgdfoam_set_pipeline", "qgd_qgdfoam_get_pipeline", "qgd_qgdfoam_diffusion_iterations", "qgd_qgdfoam_graph_steps",
    "qgd_timer_begin", "qgd_timer_end",
    "qgd_comm_unique_id", "qgd_comm_init", "qgd_comm_finalize", "qgd_qgdfoam_set_halo", "qgd_qgdfoam_set_halo_faces",
    "qgd_pcg_solve", "qgd_pcg_solve_stepwise", "qgd_pcg_solve_multi",
    "qgd_qhdfoam_create", "qgd_qhdfoam_destroy", "qgd_qhdfoam_set_bcs", "qgd_qhdfoam_init_fields", "qgd_qhdfoam_set_halo", "qgd_qhdfoam_step",
    "qgd_qhdfoam_get", "qgd_qhdfoam_get_flux", "qgd_qhdfoam_get_scalars", "qgd_qhdfoam_solver_info",
    "qgd_qhdfoam_launch_count",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class QGDError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[qgd status {code}] {msg}")
        self.code = code
        self.message = msg


class _MeshDesc(C.Structure):
    _fields_ = [("n_cells", C.c_int), ("n_faces", C.c_int), ("n_internal_faces", C.c_int), ("n_points", C.c_int),
                ("n_patches", C.c_int),
                ("points", _dp), ("face_offsets", _ip), ("face_verts", _ip), ("owner", _ip), ("neighbour", _ip),
                ("patch_start", _ip), ("patch_size", _ip), ("patch_kind", _ip),
                ("C", _dp), ("V", _dp), ("Cf", _dp), ("Sf", _dp), ("magSf", _dp), ("weights", _dp),
                ("deltaCoeffs", _dp), ("nonOrthDeltaCoeffs", _dp), ("neighb_cell_centres", _dp),
                ("geometric_d", C.c_int * 3), ("n_owned_cells", C.c_int), ("coupled_internal_face", _ip)]


class QGDFoamDesc(C.Structure):
    _fields_ = [("fvsc_scheme", C.c_char_p), ("qgd_coeffs_model", C.c_char_p),
                ("R", C.c_double), ("Cp", C.c_double), ("Hf", C.c_double), ("Tref", C.c_double), ("Hsref", C.c_double),
                ("mu", C.c_double), ("Pr", C.c_double), ("ScQGD", C.c_double), ("PrQGD", C.c_double),
                ("implicit_diffusion", C.c_int), ("alpha_eff_gamma_factor", C.c_int), ("energy_ddt_rhoE_quirk", C.c_int),
                ("adjust_time_step", C.c_int),
                ("max_co", C.c_double), ("max_delta_t", C.c_double), ("c_tau", C.c_double), ("delta_t", C.c_double),
                ("diff_tolerance", C.c_double), ("diff_rel_tol", C.c_double), ("diff_max_iter", C.c_int),
                ("diff_preconditioner", C.c_char_p),
                ("varsc_cSc1", C.c_double), ("varsc_minSc", C.c_double), ("varsc_maxSc", C.c_double),
                ("transport_model", C.c_char_p), ("As", C.c_double), ("Ts", C.c_double),
                ("mu0", C.c_double), ("T0", C.c_double), ("k_exp", C.c_double),
                ("thermo_model", C.c_char_p), ("Cv", C.c_double), ("Esref", C.c_double),
                ("varsc5_smoothCoeff", C.c_double), ("varsc5_rC", C.c_double), ("varsc5_badQualitySc", C.c_double),
                ("varsc5_maxAspectRatio", C.c_double)]


class QHDFoamDesc(C.Structure):
    _fields_ = [("fvsc_scheme", C.c_char_p), ("qgd_coeffs_model", C.c_char_p),
                ("rho0", C.c_double), ("mu", C.c_double), ("Pr", C.c_double), ("beta", C.c_double), ("g", C.c_double * 3),
                ("Tau", C.c_double), ("UQHD", C.c_double), ("Gr", C.c_double), ("T0", C.c_double),
                ("implicit_diffusion", C.c_int), ("p_tolerance", C.c_double), ("p_rel_tol", C.c_double),
                ("p_max_iter", C.c_int), ("p_preconditioner", C.c_char_p), ("p_ref_cell", C.c_int),
                ("p_ref_value", C.c_double), ("adjust_time_step", C.c_int),
                ("max_co", C.c_double), ("max_delta_t", C.c_double), ("c_tau", C.c_double), ("delta_t", C.c_double),
                ("diff_tolerance", C.c_double), ("diff_rel_tol", C.c_double), ("diff_max_iter", C.c_int),
                ("diff_preconditioner", C.c_char_p), ("scalar_transport", C.c_int)]


class _StateHost(C.Structure):
    _fields_ = [(n, _dp) for n in ("rho", "U", "e", "p", "T", "rhoU", "rhoE", "mu")]


class _FieldsHost(C.Structure):
    _fields_ = [(n, _dp) for n in ("U", "T", "p", "rho", "rhoU", "rhoE")]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; fails loudly when it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QGDError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `python -m qgdsolver_b200.build` "
                                 "(the product has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.qgd_last_error.restype = C.c_char_p
    L.qgd_init.argtypes = [C.c_int]
    L.qgd_mesh_create.argtypes = [C.POINTER(_MeshDesc), C.POINTER(C.c_void_p)]
    L.qgd_mesh_destroy.argtypes = [C.c_void_p]
    L.qgd_mesh_set_degenerate_stencil_faces.argtypes = [C.c_void_p, _ip, C.c_int]
    L.qgd_mesh_get.argtypes = [C.c_void_p, C.c_int, _dp]
    L.qgd_mesh_set_pcg_blocks.argtypes = [C.c_void_p, _ip]
    L.qgd_mesh_make_pcg_blocks.argtypes = [C.c_void_p, C.c_int, _ip]
    L.qgd_mesh_get_pcg_blocks.argtypes = [C.c_void_p, _ip]
    L.qgd_fvsc_create.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.qgd_fvsc_destroy.argtypes = [C.c_void_p]
    for fn in (L.qgd_fvsc_grad, L.qgd_fvsc_div):
        fn.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp]
    L.qgd_qgdfoam_create.argtypes = [C.c_void_p, C.POINTER(QGDFoamDesc), C.POINTER(C.c_void_p)]
    L.qgd_qgdfoam_destroy.argtypes = [C.c_void_p]
    L.qgd_qgdfoam_set_const_sc_cells.argtypes = [C.c_void_p, _ip, C.c_int]
    L.qgd_qgdfoam_set_sources.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.qgd_qgdfoam_set_bcs.argtypes = [C.c_void_p, _ip, _ip, _ip, _dp, _dp, _dp]
    L.qgd_qgdfoam_init_fields.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    L.qgd_qgdfoam_step.argtypes = [C.c_void_p, C.c_int]
    L.qgd_qgdfoam_step_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(_StateHost), C.POINTER(_StateHost)]
    L.qgd_qgdfoam_step_fields_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(_FieldsHost), C.POINTER(_FieldsHost)]
    L.qgd_qgdfoam_face_kernel.argtypes = [C.c_void_p, _ip]
    L.qgd_qgdfoam_face_kernel.restype = C.c_char_p
    L.qgd_qgdfoam_get.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.qgd_qgdfoam_get_flux.argtypes = [C.c_void_p, C.c_int, _dp]
    L.qgd_qgdfoam_get_scalars.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.qgd_qgdfoam_state_guard.argtypes = [C.c_void_p, _ip]
    L.qgd_qgdfoam_graph_steps.restype = C.c_longlong
    L.qgd_qgdfoam_graph_steps.argtypes = [C.c_void_p]
    L.qgd_qgdfoam_launch_count.restype = C.c_longlong
    L.qgd_qgdfoam_launch_count.argtypes = [C.c_void_p]
    L.qgd_qgdfoam_profile.argtypes = [C.c_void_p, C.c_int]
    L.qgd_qgdfoam_kernel_times.argtypes = [C.c_void_p, _dp, _dp, _dp, _ip]
    L.qgd_timer_end.argtypes = [C.POINTER(C.c_float)]
    L.qgd_qgdfoam_set_pipeline.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.qgd_qgdfoam_get_pipeline.argtypes = [C.c_void_p] + [_ip] * 6
    L.qgd_comm_unique_id.argtypes = [C.c_void_p]
    L.qgd_comm_init.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.qgd_qgdfoam_set_halo.argtypes = [C.c_void_p, C.c_int] + [_ip] * 9
    L.qgd_qgdfoam_set_halo_faces.argtypes = [C.c_void_p, C.c_int] + [_ip] * 5
    L.qgd_pcg_solve_multi.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int,
                                      C.c_int, _ip, _ip, _ip, _ip, _ip, _ip, _dp, _dp]
    L.qgd_pcg_solve_stepwise.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int,
                                         _ip, _dp, _dp]
    L.qgd_pcg_solve.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int,
                                _ip, _dp, _dp]
    L.qgd_qhdfoam_create.argtypes = [C.c_void_p, C.POINTER(QHDFoamDesc), C.POINTER(C.c_void_p)]
    L.qgd_qhdfoam_destroy.argtypes = [C.c_void_p]
    L.qgd_qhdfoam_set_bcs.argtypes = [C.c_void_p, _ip, _ip, _ip, _dp, _dp, _dp]
    L.qgd_qhdfoam_init_fields.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    L.qgd_qhdfoam_step.argtypes = [C.c_void_p, C.c_int]
    L.qgd_qhdfoam_set_halo.argtypes = [C.c_void_p, C.c_int] + [_ip] * 5 + [C.c_int] + [_ip] * 5
    L.qgd_qhdfoam_get.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.qgd_qhdfoam_get_flux.argtypes = [C.c_void_p, _dp]
    L.qgd_qhdfoam_get_scalars.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.qgd_qhdfoam_solver_info.argtypes = [C.c_void_p, _ip, _dp, _dp]
    L.qgd_qhdfoam_launch_count.restype = C.c_longlong
    L.qgd_qhdfoam_launch_count.argtypes = [C.c_void_p]
    _lib = L
    return L


def _check(code: int):
    if code != QGD_OK:
        raise QGDError(code, load_library().qgd_last_error().decode(errors="replace"))


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def init(device: int = 0):
    _check(load_library().qgd_init(device))


def synchronize():
    _check(load_library().qgd_device_synchronize())


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load_library().qgd_comm_unique_id(buf))
    return buf.raw


def comm_init(rank: int, n_ranks: int, uid: bytes):
    buf = C.create_string_buffer(uid, 128)
    _check(load_library().qgd_comm_init(rank, n_ranks, buf))


def comm_finalize():
    _check(load_library().qgd_comm_finalize())


class Mesh:
    """Device image of an fvMesh (qgd_mesh_create)."""

    def __init__(self, mesh: PolyMesh, n_owned: int = 0, coupled_face=None):
        self.mesh = mesh
        self.n_owned = n_owned or mesh.n_cells
        d = _MeshDesc()
        d.n_owned_cells = int(n_owned)
        self._coupled = None if coupled_face is None else np.ascontiguousarray(coupled_face, np.int32)
        d.coupled_internal_face = None if self._coupled is None else _i(self._coupled)
        d.n_cells, d.n_faces, d.n_internal_faces, d.n_points = mesh.n_cells, mesh.n_faces, mesh.n_internal, mesh.n_points
        d.n_patches = len(mesh.patches)
        k = dict(points=_f64(mesh.points), C=_f64(mesh.C), V=_f64(mesh.V), Cf=_f64(mesh.Cf), Sf=_f64(mesh.Sf),
                 magSf=_f64(mesh.magSf), weights=_f64(mesh.weights), deltaCoeffs=_f64(mesh.deltaCoeffs),
                 nonOrthDeltaCoeffs=_f64(mesh.nonOrthDeltaCoeffs),
                 neighb_cell_centres=_f64(np.nan_to_num(mesh.neighb_cell_centres)))
        ki = dict(face_offsets=np.ascontiguousarray(mesh.face_offsets, np.int32),
                  face_verts=np.ascontiguousarray(mesh.face_verts, np.int32),
                  owner=np.ascontiguousarray(mesh.owner, np.int32),
                  neighbour=np.ascontiguousarray(mesh.neighbour, np.int32),
                  patch_start=np.array([p.start for p in mesh.patches], np.int32),
                  patch_size=np.array([p.size for p in mesh.patches], np.int32),
                  patch_kind=np.array([p.kind for p in mesh.patches], np.int32))
        for n, a in k.items():
            setattr(d, n, _d(a))
        for n, a in ki.items():
            setattr(d, n, _i(a))
        for j in range(3):
            d.geometric_d[j] = int(mesh.geometric_d[j])
        self._h = C.c_void_p()
        _check(load_library().qgd_mesh_create(C.byref(d), C.byref(self._h)))

    def set_degenerate_stencil_faces(self, faces):
        """faceSet degenerateStencilFaces of the leastSquares scheme (leastSquaresStencil.C:63-132); before any stencil / solver"""
        f = np.ascontiguousarray(faces, np.int32)
        _check(load_library().qgd_mesh_set_degenerate_stencil_faces(self._h, _i(f), int(f.size)))

    def set_pcg_blocks(self, cell_block):
        """DIC blocks (block id per cell; None clears): the preconditioner of every PCG on this mesh becomes block-local"""
        if cell_block is None:
            _check(load_library().qgd_mesh_set_pcg_blocks(self._h, None))
            return
        b = np.ascontiguousarray(cell_block, np.int32)
        _check(load_library().qgd_mesh_set_pcg_blocks(self._h, _i(b)))

    def make_pcg_blocks(self, target_cells: int = 512) -> np.ndarray:
        """compact tiles of at most target_cells cells by recursive coordinate bisection; returns the block id per cell"""
        nb = C.c_int()
        _check(load_library().qgd_mesh_make_pcg_blocks(self._h, int(target_cells), C.byref(nb)))
        return self.pcg_blocks()

    def pcg_blocks(self) -> np.ndarray:
        out = np.empty(self.mesh.n_cells, np.int32)
        _check(load_library().qgd_mesh_get_pcg_blocks(self._h, _i(out)))
        return out

    def hQGDf(self):
        out = np.empty(self.mesh.n_faces)
        _check(load_library().qgd_mesh_get(self._h, 0, _d(out)))
        return out

    def hQGD(self):
        out = np.empty(self.mesh.n_cells)
        _check(load_library().qgd_mesh_get(self._h, 1, _d(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            load_library().qgd_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FvscStencil:
    """fvsc::fvscStencil::New(name, mesh) + Grad/Div (fvscStencil.H:86-130)."""

    def __init__(self, mesh: Mesh, name: str):
        self.mesh = mesh
        self._h = C.c_void_p()
        _check(load_library().qgd_fvsc_create(mesh._h, name.encode(), C.byref(self._h)))

    def _apply(self, fn, k, ok, cell, bnd, bsg, nbr):
        m = self.mesh.mesh
        cell, bnd, bsg, nbr = _f64(cell), _f64(bnd), _f64(bsg), _f64(nbr)
        out = np.zeros((m.n_faces, ok) if ok > 1 else m.n_faces)
        _check(fn(self._h, k, _d(cell), _d(bnd), _d(bsg), _d(nbr), _d(out)))
        return out

    def Grad(self, cell, bnd, bnd_sngrad, nbr=None):
        k = 1 if np.ndim(cell) == 1 else cell.shape[1]
        return self._apply(load_library().qgd_fvsc_grad, k, 3 * k, cell, bnd, bnd_sngrad, nbr)

    def Div(self, cell, bnd, bnd_sngrad, nbr=None):
        k = cell.shape[1]
        return self._apply(load_library().qgd_fvsc_div, k, k // 3, cell, bnd, bnd_sngrad, nbr)

    def close(self):
        if getattr(self, "_h", None):
            load_library().qgd_fvsc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


CELL_FIELDS = {"rho": (0, 1), "rhoU": (1, 3), "rhoE": (2, 1), "U": (3, 3), "e": (4, 1), "p": (5, 1), "T": (6, 1),
               "c": (7, 1), "mu": (8, 1), "alpha": (9, 1), "tauQGD": (10, 1), "H": (11, 1), "ScQGD": (12, 1)}


class QGDFoam:
    """The QGDFoam time loop on the device (QGDFoam.C:68-168)."""

    def __init__(self, mesh: Mesh, *, R, Cp, mu=0.0, Pr=1.0, Hf=0.0, Tref=0.0, Hsref=0.0, ScQGD=1.0, PrQGD=1.0,
                 fvsc_scheme="GaussVolPoint", qgd_coeffs="constScPrModel1", implicit_diffusion=False,
                 alpha_eff_gamma_factor=True, energy_ddt_rhoE_quirk=True, adjust_time_step=False, max_co=0.3,
                 max_delta_t=1e30, c_tau=0.75, delta_t=1e-4, diff_tol=1e-9, diff_rel_tol=0.0, diff_max_iter=1000,
                 diff_precond="DIC", varsc_cSc1=1.0, varsc_minSc=None, varsc_maxSc=None,
                 transport="const", As=0.0, Ts=0.0, mu0=0.0, T0=1.0, k_exp=0.0, thermo="hConst", Cv=0.0, Esref=0.0,
                 varsc5_smoothCoeff=0.1, varsc5_rC=0.5, varsc5_badQualitySc=0.05, varsc5_maxAspectRatio=1.5):
        self.mesh = mesh
        d = QGDFoamDesc()
        # dictionary defaults: varScModel7.C:96-119 (-1 = off), varScModel5.C:63-64 (0.05 / 1.0)
        if varsc_minSc is None:
            varsc_minSc = 0.05 if qgd_coeffs == "varScModel5" else -1.0
        if varsc_maxSc is None:
            varsc_maxSc = 1.0 if qgd_coeffs == "varScModel5" else -1.0
        d.varsc5_smoothCoeff, d.varsc5_rC = varsc5_smoothCoeff, varsc5_rC
        d.varsc5_badQualitySc, d.varsc5_maxAspectRatio = varsc5_badQualitySc, varsc5_maxAspectRatio
        self._thermo_names = (transport.encode(), thermo.encode())
        d.transport_model, d.thermo_model = self._thermo_names
        d.As, d.Ts, d.mu0, d.T0, d.k_exp, d.Cv, d.Esref = As, Ts, mu0, T0, k_exp, Cv, Esref
        d.varsc_cSc1, d.varsc_minSc, d.varsc_maxSc = varsc_cSc1, varsc_minSc, varsc_maxSc
        self._names = (fvsc_scheme.encode(), qgd_coeffs.encode(), diff_precond.encode())
        d.fvsc_scheme, d.qgd_coeffs_model, d.diff_preconditioner = self._names
        d.diff_tolerance, d.diff_rel_tol, d.diff_max_iter = diff_tol, diff_rel_tol, diff_max_iter
        d.R, d.Cp, d.Hf, d.Tref, d.Hsref, d.mu, d.Pr, d.ScQGD, d.PrQGD = R, Cp, Hf, Tref, Hsref, mu, Pr, ScQGD, PrQGD
        d.implicit_diffusion = int(implicit_diffusion)
        d.alpha_eff_gamma_factor = int(alpha_eff_gamma_factor)
        d.energy_ddt_rhoE_quirk = int(energy_ddt_rhoE_quirk)
        d.adjust_time_step = int(adjust_time_step)
        d.max_co, d.max_delta_t, d.c_tau, d.delta_t = max_co, max_delta_t, c_tau, delta_t
        self._h = C.c_void_p()
        _check(load_library().qgd_qgdfoam_create(mesh._h, C.byref(d), C.byref(self._h)))

    def set_sources(self, rhoSu=None, rhoUSu=None, rhoESu=None):
        """Explicit rhoSu / rhoUSu / rhoESu (createZeroSources.H:28-44), volume-integrated per cell; None = zero."""
        a = [_f64(x) for x in (rhoSu, rhoUSu, rhoESu)]
        _check(load_library().qgd_qgdfoam_set_sources(self._h, _d(a[0]), _d(a[1]), _d(a[2])))

    def set_const_sc_cells(self, cells):
        """varScModel7 / varScModel5 constScCellSet (varScModel7.C:143-158, varScModel5.C:134-149); call before init_fields."""
        c = np.ascontiguousarray(cells, np.int32)
        _check(load_library().qgd_qgdfoam_set_const_sc_cells(self._h, _i(c), int(c.size)))

    def set_bcs(self, bcU, bcT, bcP, valU=None, valT=None, valP=None):
        a = [np.ascontiguousarray(x, np.int32) for x in (bcU, bcT, bcP)]
        v = [_f64(x) for x in (valU, valT, valP)]
        _check(load_library().qgd_qgdfoam_set_bcs(self._h, _i(a[0]), _i(a[1]), _i(a[2]), _d(v[0]), _d(v[1]), _d(v[2])))

    def init_fields(self, U, T, p, alphaQGD=None):
        U, T, p, alphaQGD = _f64(U), _f64(T), _f64(p), _f64(alphaQGD)
        _check(load_library().qgd_qgdfoam_init_fields(self._h, _d(U), _d(T), _d(p), _d(alphaQGD)))

    def step(self, n_steps: int = 1):
        _check(load_library().qgd_qgdfoam_step(self._h, n_steps))

    def set_halo(self, sub):
        """Register the exchange lists of a decompose.SubDomain (call after comm_init)."""
        nbrs = sorted(set(sub.send_cells) | set(sub.recv_cells))

        def pack(d):
            off = np.zeros(len(nbrs) + 1, np.int32)
            parts = []
            for k, r in enumerate(nbrs):
                a = np.asarray(d.get(r, np.zeros(0, np.int32)), np.int32)
                parts.append(a)
                off[k + 1] = off[k] + a.size
            ids = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
            if ids.size == 0:
                ids = np.zeros(1, np.int32)
            return off, np.ascontiguousarray(ids)
        sco, sc = pack(sub.send_cells)
        rco, rc = pack(sub.recv_cells)
        sbo, sb = pack(sub.send_bfaces)
        rbo, rb = pack(sub.recv_bfaces)
        nb = np.asarray(nbrs, np.int32) if nbrs else np.zeros(1, np.int32)
        _check(load_library().qgd_qgdfoam_set_halo(self._h, len(nbrs), _i(nb), _i(sco), _i(sc), _i(rco), _i(rc),
                                                   _i(sbo), _i(sb), _i(rbo), _i(rb)))
        # face-neighbour subset (implicit branch: PCG search direction, fvc::grad(U))
        fn = sorted(set(sub.send_face_cells) | set(sub.recv_face_cells))
        if fn:
            def packf(d):
                off = np.zeros(len(fn) + 1, np.int32)
                parts = []
                for k, r in enumerate(fn):
                    a = np.asarray(d.get(r, np.zeros(0, np.int32)), np.int32)
                    parts.append(a)
                    off[k + 1] = off[k] + a.size
                ids = np.concatenate(parts).astype(np.int32)
                return off, np.ascontiguousarray(ids if ids.size else np.zeros(1, np.int32))
            fso, fsc = packf(sub.send_face_cells)
            fro, frc = packf(sub.recv_face_cells)
            fnb = np.asarray(fn, np.int32)
            _check(load_library().qgd_qgdfoam_set_halo_faces(self._h, len(fn), _i(fnb), _i(fso), _i(fsc), _i(fro), _i(frc)))

    @staticmethod
    def _state_struct(bufs) -> _StateHost:
        s = _StateHost()
        for n in ("rho", "U", "e", "p", "T", "rhoU", "rhoE", "mu"):
            setattr(s, n, _d(bufs[n]))
        return s

    def step_host(self, n_steps, state_in: Optional[dict], state_out: Optional[dict]):
        """state dicts hold contiguous float64 arrays rho,U,e,p,T,rhoU,rhoE,mu (ideally page-locked)."""
        si = self._state_struct(state_in) if state_in is not None else None
        so = self._state_struct(state_out) if state_out is not None else None
        _check(load_library().qgd_qgdfoam_step_host(self._h, n_steps, C.byref(si) if si else None,
                                                    C.byref(so) if so else None))

    def step_fields_host(self, n_steps, fields_in: Optional[dict], fields_out: Optional[dict]):
        """restart-style hand-off: in = U, T, p (as read by createFields.H); out = U, T, p [+ rho, rhoU, rhoE]"""
        def mk(d):
            if d is None:
                return None
            f = _FieldsHost()
            for n in ("U", "T", "p", "rho", "rhoU", "rhoE"):
                setattr(f, n, _d(d.get(n)))
            return f
        fi, fo = mk(fields_in), mk(fields_out)
        _check(load_library().qgd_qgdfoam_step_fields_host(self._h, n_steps, C.byref(fi) if fi else None,
                                                           C.byref(fo) if fo else None))

    def face_kernel(self):
        """(name of the internal-face kernel the step launches, its L2 cache-policy bits)"""
        h = C.c_int()
        name = load_library().qgd_qgdfoam_face_kernel(self._h, C.byref(h))
        return name.decode(), h.value

    def get(self, name: str, with_bnd: bool = False):
        fid, k = CELL_FIELDS[name]
        m = self.mesh.mesh
        cells = np.zeros((m.n_cells, k) if k > 1 else m.n_cells)
        bnd = np.zeros((m.n_bnd, k) if k > 1 else m.n_bnd) if with_bnd else None
        _check(load_library().qgd_qgdfoam_get(self._h, fid, _d(cells), _d(bnd)))
        return (cells, bnd) if with_bnd else cells

    def get_flux(self, which: int):
        m = self.mesh.mesh
        out = np.zeros((m.n_faces, 3) if which == 1 else m.n_faces)
        _check(load_library().qgd_qgdfoam_get_flux(self._h, which, _d(out)))
        return out

    def scalars(self):
        dt, co, t = C.c_double(), C.c_double(), C.c_double()
        _check(load_library().qgd_qgdfoam_get_scalars(self._h, C.byref(dt), C.byref(co), C.byref(t)))
        return dict(deltaT=dt.value, CoNum=co.value, time=t.value)

    def state_guard(self) -> int:
        """first step (1-based) whose update left min(e) <= 0 or min(rho) <= 0 (QGDFoam.C:142-147), 0 = never"""
        v = C.c_int()
        _check(load_library().qgd_qgdfoam_state_guard(self._h, C.byref(v)))
        return v.value

    def set_pipeline(self, mode: int, chunk_cells: int = 0, lag: int = -1, ring_slots: int = 0):
        """mode 1: pipelined face+cell kernel with the L2-resident flux ring; mode 0: two kernels, fluxes kept in HBM."""
        _check(load_library().qgd_qgdfoam_set_pipeline(self._h, mode, chunk_cells, lag, ring_slots))

    def get_pipeline(self):
        v = [C.c_int() for _ in range(6)]
        _check(load_library().qgd_qgdfoam_get_pipeline(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("mode", "chunk_cells", "lag", "ring_slots", "n_chunks", "grid"), [x.value for x in v]))

    def profile(self, enable: bool):
        _check(load_library().qgd_qgdfoam_profile(self._h, int(enable)))

    def kernel_times(self):
        a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        _check(load_library().qgd_qgdfoam_kernel_times(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        return dict(points_ms=a.value, face_ms=b.value, cell_ms=c.value, steps=n.value)

    def diffusion_iterations(self):
        it = (C.c_int * 4)()
        _check(load_library().qgd_qgdfoam_diffusion_iterations(self._h, it))
        return list(it)

    def graph_steps(self) -> int:
        """steps replayed from a captured CUDA graph so far (QGD_STEP_GRAPH=1)"""
        return int(load_library().qgd_qgdfoam_graph_steps(self._h))

    def launch_count(self) -> int:
        return int(load_library().qgd_qgdfoam_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            load_library().qgd_qgdfoam_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PRECONDS = {"none": 0, "diagonal": 1, "DIC": 2}


def pcg_solve(mesh: Mesh, diag, upper, b, x0, tol=1e-8, rel_tol=0.0, max_iter=1000, precond="DIC", stepwise=False):
    """lduMatrix PCG on the mesh addressing, fully on the device (qgd_pcg_solve; stepwise: the one-kernel-per-phase form
    of qgd_pcg_solve_stepwise that decomposed runs build on)."""
    diag, upper, b = _f64(diag), _f64(upper), _f64(b)
    x = np.array(x0, dtype=np.float64, copy=True)
    it, r0, r1 = C.c_int(), C.c_double(), C.c_double()
    fn = load_library().qgd_pcg_solve_stepwise if stepwise else load_library().qgd_pcg_solve
    _check(fn(mesh._h, _d(diag), _d(upper), _d(b), _d(x), tol, rel_tol, max_iter,
                                        PRECONDS[precond], C.byref(it), C.byref(r0), C.byref(r1)))
    return x, it.value, r0.value, r1.value


def pcg_solve_multi(mesh: Mesh, sub, diag, upper, b, x0, tol=1e-8, rel_tol=0.0, max_iter=1000, precond="diagonal"):
    """qgd_pcg_solve_multi on this rank's extended sub-mesh (`sub`: decompose.SubDomain; arrays in local numbering, n_cells long)"""
    nbrs = sorted(set(sub.send_face_cells) | set(sub.recv_face_cells))

    def pack(d):
        off = np.zeros(len(nbrs) + 1, np.int32)
        parts = []
        for k, r in enumerate(nbrs):
            a = np.asarray(d.get(r, np.zeros(0, np.int32)), np.int32)
            parts.append(a)
            off[k + 1] = off[k] + a.size
        ids = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
        return off, np.ascontiguousarray(ids if ids.size else np.zeros(1, np.int32))
    so, sc = pack(sub.send_face_cells)
    ro, rc = pack(sub.recv_face_cells)
    nb = np.ascontiguousarray(nbrs if nbrs else [0], np.int32)
    diag, upper, b = _f64(diag), _f64(upper), _f64(b)
    x = np.array(x0, dtype=np.float64, copy=True)
    it, r0, r1 = C.c_int(), C.c_double(), C.c_double()
    _check(load_library().qgd_pcg_solve_multi(mesh._h, _d(diag), _d(upper), _d(b), _d(x), tol, rel_tol, max_iter, PRECONDS[precond],
                                              len(nbrs), _i(nb), _i(so), _i(sc), _i(ro), _i(rc), C.byref(it), C.byref(r0), C.byref(r1)))
    return x, it.value, r0.value, r1.value


QHD_FIELDS = {"U": (0, 3), "T": (1, 1), "p": (2, 1), "tauQGD": (3, 1)}


class QHDFoam:
    """The QHDFoam time loop on the device (QHDFoam.C:83-139, explicit branch)."""

    def __init__(self, mesh: Mesh, *, rho0, mu, Pr, beta, g, fvsc_scheme="GaussVolPoint", qgd_coeffs="constTau",
                 Tau=0.0, UQHD=1.0, Gr=1.0, T0=1.0, implicit_diffusion=False, tol=1e-8, rel_tol=0.0, max_iter=1000,
                 precond="DIC", p_ref_cell=0, p_ref_value=0.0, adjust_time_step=False, max_co=0.3, max_delta_t=1e30,
                 c_tau=0.75, delta_t=1e-3, diff_tol=1e-9, diff_rel_tol=0.0, diff_max_iter=1000, diff_precond="DIC",
                 scalar_transport=False):
        self.mesh = mesh
        d = QHDFoamDesc()
        d.scalar_transport = int(scalar_transport)         # scalarTransportQHDFoam.C:70-135 loop body
        self._names = (fvsc_scheme.encode(), qgd_coeffs.encode(), precond.encode(), diff_precond.encode())
        d.fvsc_scheme, d.qgd_coeffs_model, d.p_preconditioner, d.diff_preconditioner = self._names
        d.diff_tolerance, d.diff_rel_tol, d.diff_max_iter = diff_tol, diff_rel_tol, diff_max_iter
        d.rho0, d.mu, d.Pr, d.beta = rho0, mu, Pr, beta
        for j in range(3):
            d.g[j] = g[j]
        d.Tau, d.UQHD, d.Gr, d.T0 = Tau, UQHD, Gr, T0
        d.implicit_diffusion = int(implicit_diffusion)
        d.p_tolerance, d.p_rel_tol, d.p_max_iter = tol, rel_tol, max_iter
        d.p_ref_cell, d.p_ref_value = p_ref_cell, p_ref_value
        d.adjust_time_step = int(adjust_time_step)
        d.max_co, d.max_delta_t, d.c_tau, d.delta_t = max_co, max_delta_t, c_tau, delta_t
        self._h = C.c_void_p()
        _check(load_library().qgd_qhdfoam_create(mesh._h, C.byref(d), C.byref(self._h)))

    def set_bcs(self, bcU, bcT, bcP, valU=None, valT=None, valP=None):
        a = [np.ascontiguousarray(x, np.int32) for x in (bcU, bcT, bcP)]
        v = [_f64(x) for x in (valU, valT, valP)]
        _check(load_library().qgd_qhdfoam_set_bcs(self._h, _i(a[0]), _i(a[1]), _i(a[2]), _d(v[0]), _d(v[1]), _d(v[2])))

    def init_fields(self, U, T, p, alphaQGD=None):
        U, T, p, alphaQGD = _f64(U), _f64(T), _f64(p), _f64(alphaQGD)
        _check(load_library().qgd_qhdfoam_init_fields(self._h, _d(U), _d(T), _d(p), _d(alphaQGD)))

    def step(self, n_steps: int = 1):
        _check(load_library().qgd_qhdfoam_step(self._h, n_steps))

    def set_halo(self, sub):
        """Register the exchange lists of a decompose.SubDomain (after comm_init, before init_fields): the vertex-ring halo and
        its face-neighbour subset."""
        def pack(send, recv):
            nbrs = sorted(set(send) | set(recv))
            out = []
            for d in (send, recv):
                off = np.zeros(len(nbrs) + 1, np.int32)
                parts = []
                for k, r in enumerate(nbrs):
                    a = np.asarray(d.get(r, np.zeros(0, np.int32)), np.int32)
                    parts.append(a)
                    off[k + 1] = off[k] + a.size
                ids = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
                out += [off, np.ascontiguousarray(ids if ids.size else np.zeros(1, np.int32))]
            return [len(nbrs), np.ascontiguousarray(nbrs if nbrs else [0], np.int32)] + out
        a = pack(sub.send_cells, sub.recv_cells)
        b = pack(sub.send_face_cells, sub.recv_face_cells)
        self._halo_keep = (a, b)
        _check(load_library().qgd_qhdfoam_set_halo(self._h, a[0], _i(a[1]), _i(a[2]), _i(a[3]), _i(a[4]), _i(a[5]),
                                                   b[0], _i(b[1]), _i(b[2]), _i(b[3]), _i(b[4]), _i(b[5])))

    def get(self, name: str, with_bnd: bool = False):
        fid, k = QHD_FIELDS[name]
        m = self.mesh.mesh
        cells = np.zeros((m.n_cells, k) if k > 1 else m.n_cells)
        bnd = np.zeros((m.n_bnd, k) if k > 1 else m.n_bnd) if with_bnd else None
        _check(load_library().qgd_qhdfoam_get(self._h, fid, _d(cells), _d(bnd)))
        return (cells, bnd) if with_bnd else cells

    def get_flux(self):
        out = np.zeros(self.mesh.mesh.n_faces)
        _check(load_library().qgd_qhdfoam_get_flux(self._h, _d(out)))
        return out

    def scalars(self):
        dt, co, t = C.c_double(), C.c_double(), C.c_double()
        _check(load_library().qgd_qhdfoam_get_scalars(self._h, C.byref(dt), C.byref(co), C.byref(t)))
        return dict(deltaT=dt.value, CoNum=co.value, time=t.value)

    def solver_info(self):
        it, r0, r1 = C.c_int(), C.c_double(), C.c_double()
        _check(load_library().qgd_qhdfoam_solver_info(self._h, C.byref(it), C.byref(r0), C.byref(r1)))
        return dict(iters=it.value, initial_residual=r0.value, final_residual=r1.value)

    def launch_count(self) -> int:
        return int(load_library().qgd_qhdfoam_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            load_library().qgd_qhdfoam_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def timer_begin():
    _check(load_library().qgd_timer_begin())


def timer_end() -> float:
    ms = C.c_float()
    _check(load_library().qgd_timer_end(C.byref(ms)))
    return float(ms.value)
