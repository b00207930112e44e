"""OpenFOAM-polyMesh-compatible mesh substrate (host side, numpy).

The reference (QGDsolver) never builds meshes itself: it consumes OpenFOAM's
``fvMesh`` (points / faces / owner / neighbour / boundary + the geometry
OpenFOAM derives from them).  OpenFOAM is not available here, so this module
produces exactly those arrays for synthetic cases:

* topology in polyMesh conventions (upper-triangular internal face order,
  boundary faces grouped per patch, face normals owner -> neighbour);
* the geometry ``fvMesh`` exposes and the reference reads:
  ``C, V, Cf, Sf, magSf, weights, deltaCoeffs, nonOrthDeltaCoeffs``
  [OF-v2312 semantics; unverified here].

Everything the reference itself derives (GaussVolPoint coefficients, hQGD,
volPointInterpolation weights ...) is NOT computed here: the CUDA library and
the CPU oracle each derive those independently from these arrays.

Used by: tests, bench.py (synthetic inputs), and as the input format of the
C-ABI ``qgd_mesh_create`` (include/qgd_b200.h).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# patch kinds (shared with include/qgd_b200.h : qgd_patch_kind)
PATCH_GENERIC = 0      # patch / wall : ordinary boundary
PATCH_EMPTY = 1        # empty        : 1D/2D reduction
PATCH_PROCESSOR = 2    # processor    : inter-subdomain
PATCH_WEDGE = 3        # wedge (no face derivatives, skipped in hQGD; vertices constrained)
PATCH_SYMMETRY_PLANE = 4   # symmetryPlane: an ordinary patch for the face derivatives, its vertices are constrained [OF pointConstraints]

_KIND_NAMES = {"patch": PATCH_GENERIC, "wall": PATCH_GENERIC,
               "empty": PATCH_EMPTY, "processor": PATCH_PROCESSOR,
               "wedge": PATCH_WEDGE, "symmetryPlane": PATCH_SYMMETRY_PLANE}


@dataclass
class Patch:
    name: str
    kind: int
    start: int
    size: int
    neighb_rank: int = -1


@dataclass
class PolyMesh:
    """polyMesh arrays + fvMesh geometry (all C-contiguous, float64/int32)."""
    points: np.ndarray            # (nPoints,3)
    face_offsets: np.ndarray      # (nFaces+1,) int32  CSR into face_verts
    face_verts: np.ndarray        # (sum nv,)  int32
    owner: np.ndarray             # (nFaces,)  int32
    neighbour: np.ndarray         # (nInternal,) int32
    patches: List[Patch]
    n_cells: int
    geometric_d: np.ndarray = field(default_factory=lambda: np.ones(3, np.int32))
    # fvMesh geometry, filled by compute_geometry()
    C: Optional[np.ndarray] = None        # (nCells,3)
    V: Optional[np.ndarray] = None        # (nCells,)
    Cf: Optional[np.ndarray] = None       # (nFaces,3)
    Sf: Optional[np.ndarray] = None       # (nFaces,3)
    magSf: Optional[np.ndarray] = None    # (nFaces,)
    weights: Optional[np.ndarray] = None  # (nFaces,) boundary = 1 (coupled: see decompose)
    deltaCoeffs: Optional[np.ndarray] = None         # (nFaces,)
    nonOrthDeltaCoeffs: Optional[np.ndarray] = None  # (nFaces,)
    # processor patches only: neighbour-side cell centres per boundary face (nBnd,3), NaN elsewhere
    neighb_cell_centres: Optional[np.ndarray] = None

    @property
    def n_points(self) -> int:
        return self.points.shape[0]

    @property
    def n_faces(self) -> int:
        return self.owner.shape[0]

    @property
    def n_internal(self) -> int:
        return self.neighbour.shape[0]

    @property
    def n_bnd(self) -> int:
        return self.n_faces - self.n_internal

    @property
    def n_geometric_d(self) -> int:
        return int((self.geometric_d > 0).sum())

    def patch_kind_per_bface(self) -> np.ndarray:
        k = np.zeros(self.n_bnd, np.int32)
        for p in self.patches:
            s = p.start - self.n_internal
            k[s:s + p.size] = p.kind
        return k

    def patch_id_per_bface(self) -> np.ndarray:
        k = np.full(self.n_bnd, -1, np.int32)
        for i, p in enumerate(self.patches):
            s = p.start - self.n_internal
            k[s:s + p.size] = i
        return k

    def face_nverts(self) -> np.ndarray:
        return np.diff(self.face_offsets)

    # ------------------------------------------------------------------ geometry
    def compute_geometry(self) -> "PolyMesh":
        """fvMesh geometry [OF-v2312: primitiveMeshTools::faceCentresAndAreas,
        cellCentresAndVols; surfaceInterpolation::makeWeights/makeDeltaCoeffs/
        makeNonOrthDeltaCoeffs; fvPatch::delta = Cf - Cn]."""
        pts = self.points
        nF = self.n_faces
        nI = self.n_internal
        nv = self.face_nverts()
        Cf = np.zeros((nF, 3))
        Sf = np.zeros((nF, 3))
        for n in np.unique(nv):
            if (nv == n).all():
                idx = slice(None)
                v = self.face_verts.reshape(-1, n)
                cnt_n = nF
            else:
                idx = np.nonzero(nv == n)[0]
                v = self.face_verts[(self.face_offsets[idx][:, None] + np.arange(n)[None, :])]
                cnt_n = idx.size
            P = pts[v]                                   # (m,n,3)
            if n == 3:
                Cf[idx] = (P[:, 0] + P[:, 1] + P[:, 2]) / 3.0
                Sf[idx] = 0.5 * _cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
            else:
                fC = P[:, 0].copy()
                for k in range(1, n):
                    fC += P[:, k]
                fC /= n
                sumN = np.zeros((cnt_n, 3))
                sumA = np.zeros(cnt_n)
                sumAc = np.zeros((cnt_n, 3))
                for k in range(n):
                    p0 = P[:, k]
                    p1 = P[:, (k + 1) % n]
                    c = p0 + p1 + fC
                    nn = _cross(p1 - p0, fC - p0)
                    a = np.sqrt((nn * nn).sum(1))
                    sumN += nn
                    sumA += a
                    sumAc += a[:, None] * c
                Cf[idx] = (1.0 / 3.0) * sumAc / sumA[:, None]
                Sf[idx] = 0.5 * sumN
        magSf = np.sqrt((Sf * Sf).sum(1))
        own = self.owner
        nei = self.neighbour
        nC = self.n_cells
        # cell centres / volumes
        cEst = np.zeros((nC, 3))
        cnt = np.bincount(own, minlength=nC) + np.bincount(nei, minlength=nC)
        for d in range(3):
            cEst[:, d] = (np.bincount(own, weights=Cf[:, d], minlength=nC)
                          + np.bincount(nei, weights=Cf[:nI, d], minlength=nC))
        cEst /= cnt[:, None]
        pyrO = (Sf * (Cf - cEst[own])).sum(1)
        pcO = 0.75 * Cf + 0.25 * cEst[own]
        pyrN = (Sf[:nI] * (cEst[nei] - Cf[:nI])).sum(1)
        pcN = 0.75 * Cf[:nI] + 0.25 * cEst[nei]
        V3 = np.bincount(own, weights=pyrO, minlength=nC) + np.bincount(nei, weights=pyrN, minlength=nC)
        C = np.zeros((nC, 3))
        for d in range(3):
            C[:, d] = (np.bincount(own, weights=pyrO * pcO[:, d], minlength=nC)
                       + np.bincount(nei, weights=pyrN * pcN[:, d], minlength=nC))
        C /= V3[:, None]
        V = V3 * (1.0 / 3.0)
        # interpolation weights and delta coefficients
        w = np.ones(nF)
        dC = np.zeros(nF)
        ndC = np.zeros(nF)
        SfOwn = np.abs((Sf[:nI] * (Cf[:nI] - C[own[:nI]])).sum(1))
        SfNei = np.abs((Sf[:nI] * (C[nei] - Cf[:nI])).sum(1))
        w[:nI] = SfNei / (SfOwn + SfNei)
        delta = C[nei] - C[own[:nI]]
        md = np.sqrt((delta * delta).sum(1))
        dC[:nI] = 1.0 / md
        nf = Sf[:nI] / magSf[:nI, None]
        ndC[:nI] = 1.0 / np.maximum((nf * delta).sum(1), 0.05 * md)
        # boundary [OF-v2312 fvPatch::delta(), "use patch-normal delta for all non-coupled BCs"]: delta = nf (nf . (Cf - Cn)), so
        # deltaCoeffs = nonOrthDeltaCoeffs = 1 / |nf . (Cf - Cn)| (the Foundation line uses the full vector Cf - Cn; the two differ
        # on non-orthogonal boundary cells only).  Processor patches (full vector, neighbour cell centre) are fixed up by decompose()
        with np.errstate(divide="ignore", invalid="ignore"):
            nfb = Sf[nI:] / magSf[nI:, None]
            dn = ((Cf[nI:] - C[own[nI:]]) * nfb).sum(1)
            db = nfb * dn[:, None]
            mdb = np.sqrt((db * db).sum(1))
            dC[nI:] = 1.0 / mdb
            ndC[nI:] = 1.0 / np.maximum((nfb * db).sum(1), 0.05 * mdb)
        self.C, self.V, self.Cf, self.Sf, self.magSf = C, V, Cf, Sf, magSf
        self.weights, self.deltaCoeffs, self.nonOrthDeltaCoeffs = w, dC, ndC
        if self.neighb_cell_centres is None:
            self.neighb_cell_centres = np.full((self.n_bnd, 3), np.nan)
        return self

    # ------------------------------------------------------------------ adjacency helpers
    def cell_faces_csr(self) -> Tuple[np.ndarray, np.ndarray]:
        """cell -> faces (ascending face id per cell, like a sequential face loop visits them)."""
        nC = self.n_cells
        allc = np.concatenate([self.owner, self.neighbour])
        allf = np.concatenate([np.arange(self.n_faces, dtype=np.int64),
                               np.arange(self.n_internal, dtype=np.int64)])
        order = np.lexsort((allf, allc))
        off = np.zeros(nC + 1, np.int64)
        np.cumsum(np.bincount(allc, minlength=nC), out=off[1:])
        return off.astype(np.int32), allf[order].astype(np.int32)


def _cross(a, b):
    out = np.empty_like(a)
    out[:, 0] = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    out[:, 1] = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    out[:, 2] = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    return out


# ---------------------------------------------------------------------- generators
def _axis(n: int, length: float, origin: float, grading: float) -> np.ndarray:
    """n+1 node coordinates; grading = last/first cell-size ratio (blockMesh simpleGrading)."""
    if n <= 0:
        raise ValueError("n must be positive")
    if grading == 1.0 or n == 1:
        return origin + length * np.arange(n + 1) / n
    r = grading ** (1.0 / (n - 1))
    s = np.concatenate([[0.0], np.cumsum(r ** np.arange(n))])
    return origin + length * s / s[-1]


def hex_box(nx: int, ny: int, nz: int,
            lengths: Sequence[float] = (1.0, 1.0, 1.0),
            origin: Sequence[float] = (0.0, 0.0, 0.0),
            grading: Sequence[float] = (1.0, 1.0, 1.0),
            patch_kinds: Optional[Dict[str, str]] = None,
            perturb: float = 0.0, seed: int = 0,
            compute_geometry: bool = True) -> PolyMesh:
    """Structured hex block in blockMesh ordering (cells and points i-fastest,
    internal faces upper-triangular, six patches xMin,xMax,yMin,yMax,zMin,zMax).

    patch_kinds maps patch name -> 'patch'|'wall'|'empty'.  ``perturb`` moves
    interior vertices by perturb*min(spacing) (seeded) to give a genuinely
    non-orthogonal unstructured-like geometry for tests.
    """
    kinds = {"xMin": "patch", "xMax": "patch", "yMin": "patch", "yMax": "patch",
             "zMin": "patch", "zMax": "patch"}
    if patch_kinds:
        kinds.update(patch_kinds)
    x = _axis(nx, lengths[0], origin[0], grading[0])
    y = _axis(ny, lengths[1], origin[1], grading[1])
    z = _axis(nz, lengths[2], origin[2], grading[2])
    npx, npy, npz = nx + 1, ny + 1, nz + 1
    K, J, I = np.meshgrid(np.arange(npz), np.arange(npy), np.arange(npx), indexing="ij")
    pts = np.stack([x[I], y[J], z[K]], axis=-1).reshape(-1, 3).astype(np.float64)
    if perturb > 0.0:
        rng = np.random.default_rng(seed)
        h = min(np.diff(x).min(), np.diff(y).min(), np.diff(z).min())
        d = (rng.random(pts.shape) - 0.5) * 2.0 * perturb * h
        # keep boundary points on their planes (and 2D/1D extrusion straight)
        ii, jj, kk = I.reshape(-1), J.reshape(-1), K.reshape(-1)
        d[(ii == 0) | (ii == nx), 0] = 0.0
        d[(jj == 0) | (jj == ny), 1] = 0.0
        d[(kk == 0) | (kk == nz), 2] = 0.0
        if kinds["zMin"] == "empty":   # extruded: identical displacement through z, none in z
            d[:, 2] = 0.0
            d2 = d.reshape(npz, npy, npx, 3)
            d2[:] = d2[0:1]
        if kinds["yMin"] == "empty":
            d[:, 1] = 0.0
            d2 = d.reshape(npz, npy, npx, 3)
            d2[:] = d2[:, 0:1]
        if kinds["xMin"] == "empty":
            d[:, 0] = 0.0
            d2 = d.reshape(npz, npy, npx, 3)
            d2[:] = d2[:, :, 0:1]
        pts = pts + d

    def pid(i, j, k):
        return (k * npy + j) * npx + i

    def cid(i, j, k):
        return (k * ny + j) * nx + i

    kc, jc, ic = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ic, jc, kc = ic.reshape(-1), jc.reshape(-1), kc.reshape(-1)
    cell = cid(ic, jc, kc)

    def xface(i, j, k):   # normal +x, at node plane i
        return np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], 1)

    def yface(i, j, k):   # normal +y
        return np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)], 1)

    def zface(i, j, k):   # normal +z
        return np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], 1)

    # internal faces: per owner cell in order, to +x, +y, +z neighbours (ascending neighbour => upper-triangular
    # order falls out of a cell-major (cell, direction) layout without sorting)
    nC = nx * ny * nz
    valid = np.stack([ic < nx - 1, jc < ny - 1, kc < nz - 1], axis=1).reshape(-1)
    fall = np.empty((nC, 3, 4), np.int32)
    fall[:, 0] = xface(np.minimum(ic + 1, nx), jc, kc)
    fall[:, 1] = yface(ic, np.minimum(jc + 1, ny), kc)
    fall[:, 2] = zface(ic, jc, np.minimum(kc + 1, nz))
    fi = fall.reshape(-1, 4)[valid]
    del fall
    oi = np.repeat(cell, 3)[valid]
    ni = (cell[:, None] + np.array([1, nx, nx * ny])[None, :]).reshape(-1)[valid]

    # boundary faces (outward normals), ordered by owner cell inside each patch
    bfaces, bown, patches = [], [], []
    start = fi.shape[0]

    def add(name, verts, own):
        nonlocal start
        o = np.argsort(own, kind="stable")
        bfaces.append(verts[o]); bown.append(own[o])
        patches.append(Patch(name, _KIND_NAMES[kinds[name]], start, own.size))
        start += own.size

    m = ic == 0
    add("xMin", xface(ic[m], jc[m], kc[m])[:, ::-1], cell[m])
    m = ic == nx - 1
    add("xMax", xface(ic[m] + 1, jc[m], kc[m]), cell[m])
    m = jc == 0
    add("yMin", yface(ic[m], jc[m], kc[m])[:, ::-1], cell[m])
    m = jc == ny - 1
    add("yMax", yface(ic[m], jc[m] + 1, kc[m]), cell[m])
    m = kc == 0
    add("zMin", zface(ic[m], jc[m], kc[m])[:, ::-1], cell[m])
    m = kc == nz - 1
    add("zMax", zface(ic[m], jc[m], kc[m] + 1), cell[m])

    fv = np.concatenate([fi] + bfaces).astype(np.int32)
    owner = np.concatenate([oi] + bown).astype(np.int32)
    nF = fv.shape[0]
    gd = np.ones(3, np.int32)
    for d, nm in enumerate(("xMin", "yMin", "zMin")):
        if kinds[nm] == "empty":
            gd[d] = -1
    mesh = PolyMesh(points=np.ascontiguousarray(pts),
                    face_offsets=(4 * np.arange(nF + 1)).astype(np.int32),
                    face_verts=np.ascontiguousarray(fv.reshape(-1)),
                    owner=owner, neighbour=ni.astype(np.int32),
                    patches=patches, n_cells=nx * ny * nz, geometric_d=gd)
    if compute_geometry:
        mesh.compute_geometry()
    return mesh


def wedge_box(nx: int, nr: int, lengths=(1.0, 1.0), r0: float = 0.5, angle_deg: float = 5.0, perturb: float = 0.0,
              seed: int = 0) -> PolyMesh:
    """Axisymmetric wedge block (axis = x): hex cells over x in [0, Lx], r in [r0, r0 + Lr], one cell thick in the circumferential
    direction, front / back planes at -+angle/2 about the centre plane z = 0 - what blockMesh produces for a `wedge` case away
    from the axis (r0 > 0: no prism cells, which GaussVolPoint rejects together with wedge patches, fvsc.C:65-82).  Patches
    xMin, xMax, yMin (inner radius), yMax (outer radius) and the two `wedge` patches zMin, zMax; geometricD = (1, 1, -1)."""
    if not r0 > 0.0:
        raise ValueError("wedge_box: r0 must be positive (cells on the axis would be prisms)")
    m = hex_box(nx, nr, 1, lengths=(lengths[0], lengths[1], 1.0), origin=(0.0, r0, 0.0), patch_kinds={"zMin": "empty", "zMax": "empty"},
                perturb=perturb, seed=seed, compute_geometry=False)
    half = np.deg2rad(angle_deg) / 2.0
    r, side = m.points[:, 1].copy(), np.where(m.points[:, 2] > 0.5, 1.0, -1.0)
    m.points[:, 1] = r * np.cos(half)
    m.points[:, 2] = side * r * np.sin(half)
    m.patches = [Patch(p.name, PATCH_WEDGE if p.name in ("zMin", "zMax") else p.kind, p.start, p.size) for p in m.patches]
    m.geometric_d = np.array([1, 1, -1], np.int32)
    return m.compute_geometry()


def prism_box(nx: int, ny: int, nz: int, lengths=(1.0, 1.0, 1.0), perturb: float = 0.0,
              seed: int = 0) -> PolyMesh:
    """Each hex of a box split into two triangular prisms (diagonal in the xy plane):
    z-faces become triangles, side faces stay quads -> exercises the tri-face
    GaussVolPoint path (GaussVolPointBase3D.C:161-318, 844-854)."""
    base = hex_box(nx, ny, nz, lengths, perturb=perturb, seed=seed, compute_geometry=False)
    pts = base.points
    npx, npy = nx + 1, ny + 1

    def pid(i, j, k):
        return (k * npy + j) * npx + i

    faces: List[Tuple[Tuple[int, ...], int, int, str]] = []   # (verts, owner, neighbour|-1, patch)
    # cell numbering: hex h -> prisms 2h (lower-right: (i,j),(i+1,j),(i+1,j+1)) and 2h+1 (upper-left)
    def hid(i, j, k):
        return (k * ny + j) * nx + i

    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                h = hid(i, j, k)
                a, b = 2 * h, 2 * h + 1
                p00, p10, p11, p01 = pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)
                q00, q10, q11, q01 = pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j + 1, k + 1), pid(i, j + 1, k + 1)
                # diagonal quad between a and b (normal from a to b): plane through p00-p11
                faces.append(((p00, q00, q11, p11), a, b, ""))
                # +x face of prism a
                if i < nx - 1:
                    faces.append(((p10, p11, q11, q10), a, 2 * hid(i + 1, j, k) + 1, ""))
                else:
                    faces.append(((p10, p11, q11, q10), a, -1, "xMax"))
                # -y face of prism a
                if j == 0:
                    faces.append(((p00, p10, q10, q00), a, -1, "yMin"))
                # +y face of prism b
                if j < ny - 1:
                    faces.append(((p01, q01, q11, p11), b, 2 * hid(i, j + 1, k), ""))
                else:
                    faces.append(((p01, q01, q11, p11), b, -1, "yMax"))
                # -x face of prism b
                if i == 0:
                    faces.append(((p00, q00, q01, p01), b, -1, "xMin"))
                # z faces (triangles)
                if k < nz - 1:
                    faces.append(((q00, q10, q11), a, 2 * hid(i, j, k + 1), ""))
                    faces.append(((q00, q11, q01), b, 2 * hid(i, j, k + 1) + 1, ""))
                else:
                    faces.append(((q00, q10, q11), a, -1, "zMax"))
                    faces.append(((q00, q11, q01), b, -1, "zMax"))
                if k == 0:
                    faces.append(((p00, p11, p10), a, -1, "zMin"))
                    faces.append(((p00, p01, p11), b, -1, "zMin"))
    return _assemble(pts, faces, 2 * nx * ny * nz,
                     ["xMin", "xMax", "yMin", "yMax", "zMin", "zMax"])


def hexprism_poly(nx: int, ny: int, nz: int, a: float = 1.0, lz: float = 1.0) -> PolyMesh:
    """Honeycomb of hexagonal prisms extruded through nz layers: caps are hexagons
    (>4 vertices -> the reference's "other faces" nf*snGrad path,
    GaussVolPointBase3D.C:103-106,760-768), sides are quads.  Boundary: 'sides','zMin','zMax'."""
    # pointy-top hexagons, axial rows offset by half a cell
    s3 = np.sqrt(3.0)
    pt_index: Dict[Tuple[int, int], int] = {}
    xy: List[Tuple[float, float]] = []

    def corner(cx, cy, m):
        ang = np.pi / 6.0 + m * np.pi / 3.0
        px, py = cx + a * np.cos(ang), cy + a * np.sin(ang)
        key = (int(round(px / a * 1e6)), int(round(py / a * 1e6)))
        if key not in pt_index:
            pt_index[key] = len(xy)
            xy.append((px, py))
        return pt_index[key]

    hexes = []
    for j in range(ny):
        for i in range(nx):
            cx = s3 * a * (i + 0.5 * (j % 2))
            cy = 1.5 * a * j
            hexes.append([corner(cx, cy, m) for m in range(6)])   # counter-clockwise
    n2 = len(xy)
    xy_arr = np.array(xy)
    pts = np.zeros(((nz + 1) * n2, 3))
    for k in range(nz + 1):
        pts[k * n2:(k + 1) * n2, :2] = xy_arr
        pts[k * n2:(k + 1) * n2, 2] = lz * k / nz
    edge_owner: Dict[Tuple[int, int], int] = {}
    for c, hx in enumerate(hexes):
        for m in range(6):
            e = (hx[m], hx[(m + 1) % 6])
            edge_owner[e] = c
    faces = []
    ncl = nx * ny
    for k in range(nz):
        for c, hx in enumerate(hexes):
            cell = k * ncl + c
            lo = [v + k * n2 for v in hx]
            hi = [v + (k + 1) * n2 for v in hx]
            for m in range(6):
                e = (hx[m], hx[(m + 1) % 6])
                other = edge_owner.get((e[1], e[0]), -1)
                quad = (lo[m], lo[(m + 1) % 6], hi[(m + 1) % 6], hi[m])   # outward for ccw hexagon
                if other < 0:
                    faces.append((quad, cell, -1, "sides"))
                elif other > c:
                    faces.append((quad, cell, k * ncl + other, ""))
            if k < nz - 1:
                faces.append((tuple(hi), cell, cell + ncl, ""))
            else:
                faces.append((tuple(hi), cell, -1, "zMax"))
            if k == 0:
                faces.append((tuple(reversed(lo)), cell, -1, "zMin"))
    return _assemble(pts, faces, ncl * nz, ["sides", "zMin", "zMax"])


def _assemble(pts, faces, n_cells, patch_names, kinds: Optional[Dict[str, str]] = None) -> PolyMesh:
    kinds = kinds or {}
    internal = [(o, n, v) for (v, o, n, p) in faces if n >= 0]
    internal.sort(key=lambda t: (t[0], t[1]))
    fv, owner, neigh = [], [], []
    for o, n, v in internal:
        assert o < n, "owner must be the lower-numbered cell"
        fv.append(v); owner.append(o); neigh.append(n)
    patches = []
    start = len(fv)
    for nm in patch_names:
        pf = [(o, v) for (v, o, n, p) in faces if n < 0 and p == nm]
        pf.sort(key=lambda t: t[0])
        for o, v in pf:
            fv.append(v); owner.append(o)
        patches.append(Patch(nm, _KIND_NAMES[kinds.get(nm, "patch")], start, len(pf)))
        start += len(pf)
    off = np.zeros(len(fv) + 1, np.int32)
    off[1:] = np.cumsum([len(v) for v in fv])
    mesh = PolyMesh(points=np.ascontiguousarray(pts, dtype=np.float64), face_offsets=off,
                    face_verts=np.array([x for v in fv for x in v], np.int32),
                    owner=np.array(owner, np.int32), neighbour=np.array(neigh, np.int32),
                    patches=patches, n_cells=n_cells)
    return mesh.compute_geometry()


def subset_mesh(mesh: PolyMesh, keep: np.ndarray, exposed_patch: str = "oldInternalFaces", exposed_kind: int = PATCH_GENERIC) -> PolyMesh:
    """subsetMesh [OF]: the cells with keep[c] true, renumbered in ascending order; faces between a kept and a removed cell
    become boundary faces of the new patch `exposed_patch` (reversed when the kept cell was the neighbour), unused points are
    dropped.  Internal faces keep their upper-triangular order, patches their order (the exposed patch comes last)."""
    keep = np.asarray(keep, bool)
    nI = mesh.n_internal
    c2n = np.full(mesh.n_cells, -1, np.int64)
    c2n[keep] = np.arange(int(keep.sum()))
    ko, kn = keep[mesh.owner[:nI]], keep[mesh.neighbour]
    f_int = np.nonzero(ko & kn)[0]
    faces = [f_int]
    flip = [np.zeros(f_int.size, bool)]
    owner = [c2n[mesh.owner[f_int]]]
    neigh = c2n[mesh.neighbour[f_int]]
    patches, start = [], f_int.size
    for p in mesh.patches:
        ids = np.arange(p.start, p.start + p.size)
        sel = ids[keep[mesh.owner[ids]]]
        faces.append(sel); flip.append(np.zeros(sel.size, bool)); owner.append(c2n[mesh.owner[sel]])
        patches.append(Patch(p.name, p.kind, start, sel.size, p.neighb_rank))
        start += sel.size
    exp_o = np.nonzero(ko & ~kn)[0]                       # kept owner: orientation stays
    exp_n = np.nonzero(~ko & kn)[0]                       # kept neighbour: face reversed
    ids = np.concatenate([exp_o, exp_n])
    fl = np.concatenate([np.zeros(exp_o.size, bool), np.ones(exp_n.size, bool)])
    own_new = np.concatenate([c2n[mesh.owner[exp_o]], c2n[mesh.neighbour[exp_n]]])
    o = np.lexsort((ids, own_new))
    faces.append(ids[o]); flip.append(fl[o]); owner.append(own_new[o])
    patches.append(Patch(exposed_patch, exposed_kind, start, ids.size))
    gf, gflip = np.concatenate(faces), np.concatenate(flip)
    used = np.zeros(mesh.n_points, bool)
    nv = mesh.face_nverts()
    idx = np.repeat(mesh.face_offsets[gf], nv[gf]) + _ragged(nv[gf])
    used[mesh.face_verts[idx]] = True
    p2n = np.full(mesh.n_points, -1, np.int64)
    p2n[used] = np.arange(int(used.sum()))
    verts, offs = [], [0]
    for f, r in zip(gf, gflip):
        v = mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]
        if r:
            v = np.concatenate([v[:1], v[:0:-1]])
        verts.append(p2n[v]); offs.append(offs[-1] + v.size)
    sub = PolyMesh(points=np.ascontiguousarray(mesh.points[used]), face_offsets=np.asarray(offs, np.int32),
                   face_verts=np.concatenate(verts).astype(np.int32), owner=np.concatenate(owner).astype(np.int32),
                   neighbour=neigh.astype(np.int32), patches=patches, n_cells=int(keep.sum()), geometric_d=mesh.geometric_d.copy())
    return sub.compute_geometry()


def _ragged(counts: np.ndarray) -> np.ndarray:
    counts = np.asarray(counts, np.int64)
    starts = np.zeros(counts.size, np.int64)
    np.cumsum(counts[:-1], out=starts[1:])
    return np.arange(int(counts.sum()), dtype=np.int64) - np.repeat(starts, counts)


def forward_step(n: int = 40, thick: float = 0.05) -> PolyMesh:
    """The Mach-3 forward-facing step (Woodward & Colella; BASELINE configs[1]): channel 3 x 1 with a step of height 0.2 starting
    at x = 0.6, one cell thick (empty front / back), square cells of size 1/n.  Patches: xMin inlet, xMax outlet, yMin / yMax
    walls, zMin / zMax empty, `step` = the two faces of the step."""
    nx, ny = 3 * n, n
    m = hex_box(nx, ny, 1, lengths=(3.0, 1.0, thick), patch_kinds={"zMin": "empty", "zMax": "empty"})
    inside = (m.C[:, 0] > 0.6) & (m.C[:, 1] < 0.2)
    s = subset_mesh(m, ~inside, "step")
    s.geometric_d = np.array([1, 1, -1], np.int32)
    return s


def truncated_octahedron_box(nx: int, ny: int, nz: int, h: float = 1.0) -> PolyMesh:
    """Space-filling polyhedral mesh (bitruncated cubic honeycomb): truncated octahedra centred on a BCC lattice with cubic
    constant h - nx*ny*nz cells on the corner sub-lattice plus (nx-1)(ny-1)(nz-1) on the body-centre sub-lattice.  Every cell
    has 14 faces (8 hexagons shared with the other sub-lattice, 6 squares with its own): about 7 internal faces per cell, most
    of them polygons with more than 4 vertices - the "other faces" path of GaussVolPoint (GaussVolPointBase3D.C:760-768), i.e. the
    shape of BASELINE configs[4].  Fully vectorised (no Python loop over cells); cells are numbered plane by plane so that
    neighbours are close in index; the ragged outer surface is split into the six patches xMin..zMax by position."""
    ia, ja, ka = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ib, jb, kb = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing="ij")
    cen = np.concatenate([np.stack([4 * ia, 4 * ja, 4 * ka], -1).reshape(-1, 3),
                          np.stack([4 * ib + 2, 4 * jb + 2, 4 * kb + 2], -1).reshape(-1, 3)]).astype(np.int64)
    order = np.lexsort((cen[:, 0], cen[:, 1], cen[:, 2]))           # z-major planes, A and B layers interleaved
    cen = cen[order]
    nC = cen.shape[0]
    big = 8 * (max(nx, ny, nz) + 2)
    key = lambda p: ((p[..., 2] + 4) * big + (p[..., 1] + 4)) * big + (p[..., 0] + 4)
    ckeys = key(cen)
    csort = np.argsort(ckeys)

    def cell_at(p):                                                 # cell index of centre p, -1 if absent
        k = key(p)
        pos = np.searchsorted(ckeys[csort], k)
        pos = np.minimum(pos, nC - 1)
        hit = ckeys[csort][pos] == k
        return np.where(hit, csort[pos], -1)

    # the 14 face templates: direction to the neighbour centre and the vertex offsets, counter-clockwise seen from outside
    tmpl = []
    for ax in range(3):
        for sg in (1, -1):
            a1, a2 = (ax + 1) % 3, (ax + 2) % 3
            ring = [(1, 0), (0, 1), (-1, 0), (0, -1)]
            if sg < 0:
                ring = ring[::-1]
            v = np.zeros((4, 3), np.int64)
            for q, (u, w) in enumerate(ring):
                v[q, ax], v[q, a1], v[q, a2] = 2 * sg, u, w
            d = np.zeros(3, np.int64); d[ax] = 4 * sg
            tmpl.append((d, v))
    base = np.array([(2, 1, 0), (1, 2, 0), (0, 2, 1), (0, 1, 2), (1, 0, 2), (2, 0, 1)], np.int64)
    for sx in (1, -1):
        for sy in (1, -1):
            for sz in (1, -1):
                s = np.array([sx, sy, sz], np.int64)
                v = base * s
                if sx * sy * sz < 0:
                    v = v[::-1]
                tmpl.append((2 * s, v))
    f_verts, f_nv, f_own, f_nei, f_cf = [], [], [], [], []
    me = np.arange(nC)
    for d, v in tmpl:
        nb = cell_at(cen + d)
        mk = (nb > me) | (nb < 0)                                   # one face per pair (owner = lower index) + boundary faces
        own = me[mk]
        f_own.append(own); f_nei.append(nb[mk])
        f_verts.append((cen[own][:, None, :] + v[None]).reshape(-1, 3))
        f_nv.append(np.full(own.size, v.shape[0], np.int64))
        f_cf.append(cen[own] * 2 + d)                               # twice the face centre (integer)
    own = np.concatenate(f_own); nei = np.concatenate(f_nei); nv = np.concatenate(f_nv)
    vcoord = np.concatenate(f_verts); cf2 = np.concatenate(f_cf)
    starts = np.zeros(own.size + 1, np.int64); np.cumsum(nv, out=starts[1:])
    vk, vinv = np.unique(key(vcoord), return_inverse=True)
    first = np.zeros(vk.size, np.int64); first[vinv[::-1]] = np.arange(vcoord.shape[0])[::-1]
    pts = vcoord[first].astype(np.float64) * (h / 4.0)
    # face order: internal faces upper-triangular, then the six patches (by position of the face centre), owner-sorted
    internal = nei >= 0
    lo = np.array([0, 0, 0]) * 2 - 0; hi = np.array([4 * (nx - 1), 4 * (ny - 1), 4 * (nz - 1)]) * 2
    dist = np.stack([lo[0] - cf2[:, 0], cf2[:, 0] - hi[0], lo[1] - cf2[:, 1], cf2[:, 1] - hi[1], lo[2] - cf2[:, 2], cf2[:, 2] - hi[2]], 1)
    patch_of = np.argmax(dist, 1)
    grp = np.where(internal, -1, patch_of)
    sort_key = np.lexsort((np.where(internal, nei, 0), own, grp))
    own, nei, nv, grp = own[sort_key], nei[sort_key], nv[sort_key], grp[sort_key]
    src_start = starts[:-1][sort_key]
    idx = np.repeat(src_start, nv) + _ragged(nv)
    face_verts = vinv[idx].astype(np.int32)
    offs = np.zeros(own.size + 1, np.int32); np.cumsum(nv, out=offs[1:])
    nI = int(internal.sum())
    names = ["xMin", "xMax", "yMin", "yMax", "zMin", "zMax"]
    patches, start = [], nI
    for pi, nm in enumerate(names):
        cnt = int((grp == pi).sum())
        patches.append(Patch(nm, PATCH_GENERIC, start, cnt))
        start += cnt
    mesh = PolyMesh(points=np.ascontiguousarray(pts), face_offsets=offs, face_verts=face_verts, owner=own.astype(np.int32),
                    neighbour=nei[:nI].astype(np.int32), patches=patches, n_cells=nC)
    return mesh.compute_geometry()
