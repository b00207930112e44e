"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (parity unpinned, see qgd_oracle.hpp).

ctypes binding of oracle/libqgd_oracle.so.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package;
the product package (qgdsolver_b200) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqgd_oracle.so")

FVSC_GAUSSVOLPOINT, FVSC_REDUCED, FVSC_LEASTSQUARES = 0, 1, 2
FVSC_SCHEMES = {"GaussVolPoint": 0, "reduced": 1, "leastSquares": 2, "leastSquaresOpt": 3}
BC_FIXED_VALUE, BC_ZERO_GRADIENT, BC_FIXED_GRADIENT, BC_QGD_FLUX, BC_CALCULATED = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Mesh(C.Structure):
    _fields_ = [("nCells", C.c_int), ("nFaces", C.c_int), ("nInternal", C.c_int), ("nPoints", C.c_int),
                ("nPatches", C.c_int),
                ("points", _dp), ("faceOff", _ip), ("faceVerts", _ip), ("owner", _ip), ("neighbour", _ip),
                ("patchStart", _ip), ("patchSize", _ip), ("patchKind", _ip),
                ("C", _dp), ("V", _dp), ("Cf", _dp), ("Sf", _dp), ("magSf", _dp), ("weights", _dp),
                ("deltaCoeffs", _dp), ("nonOrthDeltaCoeffs", _dp), ("neighbCellCentres", _dp),
                ("geometricD", C.c_int * 3)]


class QGDParams(C.Structure):
    _fields_ = [("R", C.c_double), ("Cp", C.c_double), ("Hf", C.c_double), ("Tref", C.c_double),
                ("Hsref", C.c_double), ("mu", C.c_double), ("Pr", C.c_double), ("ScQGD", C.c_double),
                ("PrQGD", C.c_double), ("implicitDiffusion", C.c_int), ("alphaEffGammaFactor", C.c_int),
                ("energyDdtRhoEQuirk", C.c_int), ("qgdModel", C.c_int),
                ("diffTol", C.c_double), ("diffRelTol", C.c_double), ("diffMaxIter", C.c_int), ("diffPrecond", C.c_int),
                ("varScCSc1", C.c_double), ("varScMinSc", C.c_double), ("varScMaxSc", C.c_double),
                ("transportModel", C.c_int), ("mu0", C.c_double), ("T0", C.c_double), ("kExp", C.c_double),
                ("As", C.c_double), ("Ts", C.c_double), ("thermoModel", C.c_int), ("Cv", C.c_double), ("Esref", C.c_double),
                ("varSc5SmoothCoeff", C.c_double), ("varSc5RC", C.c_double), ("varSc5BadQualitySc", C.c_double),
                ("varSc5MaxAspectRatio", C.c_double)]


QGD_MODELS = {"constScPrModel1": 0, "constScPrModel1n": 1, "constScPrModel2": 2, "varScModel5": 5, "varScModel6": 6,
              "varScModel7": 7}


class QHDParams(C.Structure):
    _fields_ = [("rho0", C.c_double), ("mu", C.c_double), ("Pr", C.c_double), ("beta", C.c_double),
                ("g", C.c_double * 3), ("qgdModel", C.c_int), ("Tau", C.c_double), ("UQHD", C.c_double),
                ("Gr", C.c_double), ("T0", C.c_double), ("implicitDiffusion", C.c_int),
                ("pTol", C.c_double), ("pRelTol", C.c_double), ("pMaxIter", C.c_int), ("pPrecond", C.c_int),
                ("pRefCell", C.c_int), ("pRefValue", C.c_double),
                ("diffTol", C.c_double), ("diffRelTol", C.c_double), ("diffMaxIter", C.c_int), ("diffPrecond", C.c_int),
                ("scalarTransport", C.c_int)]


QHD_MODELS = {"constTau": 0, "H2bynuQHD": 1, "HbyUQHD": 2, "T0byGr": 3}
PRECONDS = {"none": 0, "diagonal": 1, "DIC": 2}


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("qgd_oracle.cpp", "qgd_oracle.hpp", "qhd_oracle.inc")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.or_create.restype = C.c_void_p
        L.or_create.argtypes = [C.POINTER(_Mesh), C.c_int]
        L.or_destroy.argtypes = [C.c_void_p]
        L.or_get_hQGDf.argtypes = [C.c_void_p, _dp]
        L.or_get_hQGD.argtypes = [C.c_void_p, _dp]
        for fn in (L.or_fvsc_grad, L.or_fvsc_div):
            fn.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.or_vol_point_interpolate.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.or_linear_interpolate.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.or_qgd_init.argtypes = [C.c_void_p, C.POINTER(QGDParams), C.c_int, _ip, _ip, _ip, _dp, _dp, _dp,
                                  _dp, _dp, _dp, _dp, C.c_double]
        L.or_qgd_set_const_sc_cells.argtypes = [C.c_void_p, _ip, C.c_int]
        L.or_qgd_set_sources.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.or_fvc_smooth.restype = C.c_int
        L.or_fvc_smooth.argtypes = [C.c_void_p, _dp, C.c_double]
        L.or_varsc5_cell_quality.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, _dp]
        L.or_qgd_step.restype = C.c_double
        L.or_qgd_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.or_qgd_deltaT.restype = C.c_double
        L.or_qgd_deltaT.argtypes = [C.c_void_p]
        L.or_qgd_get.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.or_qgd_get_face.argtypes = [C.c_void_p, C.c_int, _dp]
        L.or_pcg_solve.restype = C.c_int
        L.or_pcg_solve.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp]
        L.or_set_pcg_blocks.argtypes = [C.c_void_p, _ip]
        L.or_set_degenerate_faces.argtypes = [C.c_void_p, _ip, C.c_int]
        L.or_pcg_solve_blocks.restype = C.c_int
        L.or_pcg_solve_blocks.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _ip]
        L.or_qhd_init.argtypes = [C.c_void_p, C.POINTER(QHDParams), C.c_int, _ip, _ip, _ip, _dp, _dp, _dp,
                                  _dp, _dp, _dp, _dp, C.c_double]
        L.or_qhd_step.restype = C.c_double
        L.or_qhd_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.or_qhd_deltaT.restype = C.c_double
        L.or_qhd_deltaT.argtypes = [C.c_void_p]
        L.or_qhd_get.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.or_qhd_get_face.argtypes = [C.c_void_p, C.c_int, _dp]
        L.or_qhd_solver_info.argtypes = [C.c_void_p, _ip, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


_CELL_K = {0: 1, 1: 3, 2: 1, 3: 3, 4: 1, 5: 1, 6: 1, 7: 1, 8: 1, 9: 1, 10: 1, 12: 1}
_FACE_K = {0: 1, 1: 3, 2: 3, 3: 3, 4: 1, 5: 1, 6: 1, 7: 1, 8: 9, 9: 3, 10: 3, 11: 3, 12: 1, 13: 3, 14: 1}
CELL_FIELDS = {"rho": 0, "rhoU": 1, "rhoE": 2, "U": 3, "e": 4, "p": 5, "T": 6, "c": 7, "mu": 8, "alpha": 9,
               "tauQGD": 10, "ScQGD": 12}
FACE_FIELDS = {"phiJm": 0, "phiJmU": 1, "phiP": 2, "phiPi": 3, "phiJmH": 4, "phiQ": 5, "phiPiU": 6,
               "tauQGDf": 7, "gradUf": 8, "gradef": 9, "gradRhof": 10, "gradPf": 11, "phiwStar": 12, "phiTauMC": 13, "phiSigmaDotU": 14}


class Oracle:
    """One CPU-oracle context bound to a PolyMesh (qgdsolver_b200.polymesh.PolyMesh)."""

    def __init__(self, mesh, n_threads: int = 1):
        self.mesh = mesh
        m = _Mesh()
        m.nCells, m.nFaces, m.nInternal, m.nPoints = mesh.n_cells, mesh.n_faces, mesh.n_internal, mesh.n_points
        m.nPatches = len(mesh.patches)
        self._keep = dict(
            points=_f64(mesh.points), faceOff=np.ascontiguousarray(mesh.face_offsets, np.int32),
            faceVerts=np.ascontiguousarray(mesh.face_verts, np.int32),
            owner=np.ascontiguousarray(mesh.owner, np.int32), neighbour=np.ascontiguousarray(mesh.neighbour, np.int32),
            patchStart=np.array([p.start for p in mesh.patches], np.int32),
            patchSize=np.array([p.size for p in mesh.patches], np.int32),
            patchKind=np.array([p.kind for p in mesh.patches], np.int32),
            C=_f64(mesh.C), V=_f64(mesh.V), Cf=_f64(mesh.Cf), Sf=_f64(mesh.Sf), magSf=_f64(mesh.magSf),
            weights=_f64(mesh.weights), deltaCoeffs=_f64(mesh.deltaCoeffs),
            nonOrthDeltaCoeffs=_f64(mesh.nonOrthDeltaCoeffs),
            neighbCellCentres=_f64(np.nan_to_num(mesh.neighb_cell_centres)))
        k = self._keep
        for name in ("points", "C", "V", "Cf", "Sf", "magSf", "weights", "deltaCoeffs", "nonOrthDeltaCoeffs",
                     "neighbCellCentres"):
            setattr(m, name, _d(k[name]))
        for name in ("faceOff", "faceVerts", "owner", "neighbour", "patchStart", "patchSize", "patchKind"):
            setattr(m, name, _i(k[name]))
        for d in range(3):
            m.geometricD[d] = int(mesh.geometric_d[d])
        self._h = C.c_void_p(lib().or_create(C.byref(m), n_threads))

    def close(self):
        if self._h:
            lib().or_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- derived
    def hQGDf(self):
        out = np.empty(self.mesh.n_faces)
        lib().or_get_hQGDf(self._h, _d(out))
        return out

    def hQGD(self):
        out = np.empty(self.mesh.n_cells)
        lib().or_get_hQGD(self._h, _d(out))
        return out

    # ---- operators
    def _op(self, fn, scheme, k, ok, cell, bnd, bsg, nbr):
        cell, bnd, bsg, nbr = _f64(cell), _f64(bnd), _f64(bsg), _f64(nbr)
        out = np.zeros((self.mesh.n_faces, ok) if ok > 1 else self.mesh.n_faces)
        fn(self._h, scheme, k, _d(cell), _d(bnd), _d(bsg), _d(nbr), _d(out))
        return out

    def fvsc_grad(self, cell, bnd, bnd_sngrad, scheme=FVSC_GAUSSVOLPOINT, nbr=None):
        k = 1 if np.ndim(cell) == 1 else cell.shape[1]
        return self._op(lib().or_fvsc_grad, scheme, k, 3 * k, cell, bnd, bnd_sngrad, nbr)

    def fvsc_div(self, cell, bnd, bnd_sngrad, scheme=FVSC_GAUSSVOLPOINT, nbr=None):
        k = cell.shape[1]
        return self._op(lib().or_fvsc_div, scheme, k, k // 3, cell, bnd, bnd_sngrad, nbr)

    def vol_point_interpolate(self, cell, bnd):
        k = 1 if np.ndim(cell) == 1 else cell.shape[1]
        cell, bnd = _f64(cell), _f64(bnd)
        out = np.zeros((self.mesh.n_points, k) if k > 1 else self.mesh.n_points)
        lib().or_vol_point_interpolate(self._h, k, _d(cell), _d(bnd), _d(out))
        return out

    def linear_interpolate(self, cell, bnd):
        k = 1 if np.ndim(cell) == 1 else cell.shape[1]
        cell, bnd = _f64(cell), _f64(bnd)
        out = np.zeros((self.mesh.n_faces, k) if k > 1 else self.mesh.n_faces)
        lib().or_linear_interpolate(self._h, k, _d(cell), _d(bnd), _d(out))
        return out

    # ---- QGDFoam
    def qgd_init(self, params: QGDParams, bcU, bcT, bcP, bvU, bvT, bvP, U0, T0, p0, alphaQGD=None,
                 deltaT=1e-4, scheme=FVSC_GAUSSVOLPOINT, const_sc_cells=None):
        if const_sc_cells is not None:
            cs = np.ascontiguousarray(const_sc_cells, np.int32)
            lib().or_qgd_set_const_sc_cells(self._h, _i(cs), len(cs))
        a = [np.ascontiguousarray(x, np.int32) for x in (bcU, bcT, bcP)]
        f = [_f64(x) for x in (bvU, bvT, bvP, U0, T0, p0, alphaQGD)]
        lib().or_qgd_init(self._h, C.byref(params), scheme, _i(a[0]), _i(a[1]), _i(a[2]),
                          *[_d(x) for x in f], deltaT)

    def fvc_smooth(self, field, coeff):
        """[OF-v2312] fvc::smooth(field, coeff) as varScModel5.C:232 uses it; returns (smoothed field, FaceCellWave iterations)"""
        f = np.array(field, dtype=np.float64, copy=True)
        it = lib().or_fvc_smooth(self._h, _d(f), float(coeff))
        return f, it

    def varsc5_cell_quality(self, bad_quality_sc=0.05, max_aspect_ratio=1.5):
        """varScModel5.C:112-132: (cqSc, aspectRatio) per cell"""
        q, ar = np.zeros(self.mesh.n_cells), np.zeros(self.mesh.n_cells)
        lib().or_varsc5_cell_quality(self._h, bad_quality_sc, max_aspect_ratio, _d(q), _d(ar))
        return q, ar

    def qgd_set_sources(self, suRho=None, suU=None, suE=None):
        a = [_f64(x) for x in (suRho, suU, suE)]
        lib().or_qgd_set_sources(self._h, _d(a[0]), _d(a[1]), _d(a[2]))

    def qgd_step(self, n_steps=1, adjust=False, maxCo=0.3, maxDeltaT=1e30, cTau=0.75):
        return lib().or_qgd_step(self._h, n_steps, int(adjust), maxCo, maxDeltaT, cTau)

    def deltaT(self):
        return lib().or_qgd_deltaT(self._h)

    def get(self, name, with_bnd=False):
        fid = CELL_FIELDS[name]
        k = _CELL_K[fid]
        cells = np.zeros((self.mesh.n_cells, k) if k > 1 else self.mesh.n_cells)
        bnd = np.zeros((self.mesh.n_bnd, k) if k > 1 else self.mesh.n_bnd)
        lib().or_qgd_get(self._h, fid, _d(cells), _d(bnd))
        return (cells, bnd) if with_bnd else cells

    def get_face(self, name):
        fid = FACE_FIELDS[name]
        k = _FACE_K[fid]
        out = np.zeros((self.mesh.n_faces, k) if k > 1 else self.mesh.n_faces)
        lib().or_qgd_get_face(self._h, fid, _d(out))
        return out

    def set_degenerate_faces(self, faces):
        """faceSet degenerateStencilFaces of the leastSquares scheme (leastSquaresStencil.C:63-132)"""
        f = np.ascontiguousarray(faces, np.int32)
        lib().or_set_degenerate_faces(self._h, _i(f), int(f.size))

    def set_pcg_blocks(self, cell_block=None):
        """linear solvers of the following steps in the decomposed-run form (block-local preconditioner)"""
        if cell_block is None:
            lib().or_set_pcg_blocks(self._h, None)
        else:
            cb = np.ascontiguousarray(cell_block, np.int32)
            lib().or_set_pcg_blocks(self._h, _i(cb))

    def pcg_solve(self, diag, upper, b, x0, tol=1e-8, relTol=0.0, maxIter=1000, precond=2, cell_block=None):
        """cell_block: processor of each cell -> the decomposed-run solver (block-local preconditioner, global reductions)"""
        diag, upper, b = _f64(diag), _f64(upper), _f64(b)
        x = np.array(x0, dtype=np.float64, copy=True)
        r0, r1 = C.c_double(), C.c_double()
        if cell_block is None:
            it = lib().or_pcg_solve(self._h, _d(diag), _d(upper), _d(b), _d(x), tol, relTol, maxIter, precond,
                                    C.byref(r0), C.byref(r1))
        else:
            cb = np.ascontiguousarray(cell_block, np.int32)
            it = lib().or_pcg_solve_blocks(self._h, _d(diag), _d(upper), _d(b), _d(x), tol, relTol, maxIter, precond,
                                           C.byref(r0), C.byref(r1), _i(cb))
        return x, it, r0.value, r1.value

    # ---- QHDFoam
    QHD_CELL = {"U": (0, 3), "T": (1, 1), "p": (2, 1), "tauQGD": (3, 1)}
    QHD_FACE = {"phi": (0, 1), "phiu": (1, 1), "phiwo": (2, 1), "tauQGDf": (3, 1), "gradPf": (4, 3), "gradUf": (5, 9),
                "gradTf": (6, 3)}

    def qhd_init(self, params: QHDParams, bcU, bcT, bcP, bvU, bvT, bvP, U0, T0, p0, alphaQGD=None,
                 deltaT=1e-4, scheme=FVSC_GAUSSVOLPOINT):
        a = [np.ascontiguousarray(x, np.int32) for x in (bcU, bcT, bcP)]
        f = [_f64(x) for x in (bvU, bvT, bvP, U0, T0, p0, alphaQGD)]
        lib().or_qhd_init(self._h, C.byref(params), scheme, _i(a[0]), _i(a[1]), _i(a[2]),
                          *[_d(x) for x in f], deltaT)

    def qhd_step(self, n_steps=1, adjust=False, maxCo=0.3, maxDeltaT=1e30, cTau=0.75):
        return lib().or_qhd_step(self._h, n_steps, int(adjust), maxCo, maxDeltaT, cTau)

    def qhd_deltaT(self):
        return lib().or_qhd_deltaT(self._h)

    def qhd_get(self, name, with_bnd=False):
        fid, k = self.QHD_CELL[name]
        cells = np.zeros((self.mesh.n_cells, k) if k > 1 else self.mesh.n_cells)
        bnd = np.zeros((self.mesh.n_bnd, k) if k > 1 else self.mesh.n_bnd)
        lib().or_qhd_get(self._h, fid, _d(cells), _d(bnd))
        return (cells, bnd) if with_bnd else cells

    def qhd_get_face(self, name):
        fid, k = self.QHD_FACE[name]
        out = np.zeros((self.mesh.n_faces, k) if k > 1 else self.mesh.n_faces)
        lib().or_qhd_get_face(self._h, fid, _d(out))
        return out

    def qhd_solver_info(self):
        it, r0, r1 = C.c_int(), C.c_double(), C.c_double()
        lib().or_qhd_solver_info(self._h, C.byref(it), C.byref(r0), C.byref(r1))
        return dict(iters=it.value, initial_residual=r0.value, final_residual=r1.value)
