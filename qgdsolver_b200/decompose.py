"""Spatial domain decomposition for the multi-GPU path (one subdomain per GPU).

The reference runs on OpenFOAM `decomposePar` output: every rank owns `processorN/` and talks to its neighbours
through processor patches (cell halo, C1/C2 in SURVEY 5.8) and `globalMeshData` point synchronisation (C3).  The
B200 design replaces both exchanges by ONE packed exchange per step over a vertex-ring halo:

* a rank's device mesh is an *extended sub-mesh*: its owned cells first, then every cell of another rank that shares
  at least one mesh point with an owned cell (the halo), with all faces whose two cells are in that set;
* faces between an owned and a halo cell play the role of the reference's processor-patch faces: the owned cell is
  the face owner, the face is flipped if needed (OpenFOAM stores the neighbour side reversed), and the coupled-patch
  rules for weights and hQGDf apply (QGDCoeffs.C:195-199,310-317);
* because every cell around every point of an owned cell is present, the inverse-distance point interpolation is
  complete locally and no point synchronisation is needed;
* halo cell state (and the boundary state of physical boundary faces of halo cells) is received from the owning rank
  after every cell update; nothing computed locally for halo cells is ever used.

This module is host-side set-up logic (numpy).  `cell_rank` may come from OpenFOAM's `cellDecomposition` file
(scotch) when run as a plug-in; `geometric_split` is the deterministic stand-in used for synthetic benches (scotch is
not available in this image).  All maps are kept so results go back to polyMesh numbering bit-exactly.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from .polymesh import PATCH_EMPTY, PATCH_PROCESSOR, Patch, PolyMesh

PATCH_CUT = PATCH_EMPTY   # faces towards cells outside the extended set carry no data: treated like `empty`


def split_factors(n_parts: int) -> Tuple[int, int, int]:
    f = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2), 3: (3, 1, 1), 6: (3, 2, 1), 16: (4, 2, 2)}
    if n_parts not in f:
        raise ValueError(f"no geometric split defined for {n_parts} parts")
    return f[n_parts]


def geometric_split(mesh: PolyMesh, n_parts: int) -> np.ndarray:
    """decomposePar `simple`-like split: equal-count slabs along x, then y, then z (by cell-centre rank)."""
    fx, fy, fz = split_factors(n_parts)
    rank = np.zeros(mesh.n_cells, np.int32)

    def slabs(vals, k, groups):
        out = np.zeros(vals.size, np.int32)
        for g in np.unique(groups):
            idx = np.nonzero(groups == g)[0]
            order = idx[np.argsort(vals[idx], kind="stable")]
            out[order] = (np.arange(order.size) * k) // order.size
        return out

    ix = slabs(mesh.C[:, 0], fx, np.zeros(mesh.n_cells, np.int32))
    iy = slabs(mesh.C[:, 1], fy, ix)
    iz = slabs(mesh.C[:, 2], fz, ix + fx * iy)
    rank = ix + fx * (iy + fy * iz)
    return rank.astype(np.int32)


@dataclass
class SubDomain:
    rank: int
    mesh: PolyMesh                      # extended sub-mesh (owned cells first)
    n_owned: int
    cell_global: np.ndarray             # local cell  -> global cell   (cellProcAddressing analogue)
    face_global: np.ndarray             # local face  -> global face   (faceProcAddressing analogue)
    face_flipped: np.ndarray            # local face stored reversed w.r.t. the global face
    point_global: np.ndarray            # local point -> global point  (pointProcAddressing analogue)
    coupled_face: np.ndarray            # (n_internal,) 1 where the face joins an owned and a halo cell
    # exchange lists, per neighbour rank: local ids, ordered by global id on both sides
    recv_cells: Dict[int, np.ndarray] = field(default_factory=dict)
    send_cells: Dict[int, np.ndarray] = field(default_factory=dict)
    recv_bfaces: Dict[int, np.ndarray] = field(default_factory=dict)   # boundary-face index (face - n_internal)
    send_bfaces: Dict[int, np.ndarray] = field(default_factory=dict)
    # the face-neighbour subset of the halo (cells across a coupled face): what a linear-solver iteration exchanges
    # (the search direction of PCG), as opposed to the full vertex-ring state exchange once per step
    recv_face_cells: Dict[int, np.ndarray] = field(default_factory=dict)
    send_face_cells: Dict[int, np.ndarray] = field(default_factory=dict)


def _point_cells(mesh: PolyMesh):
    """(point, cell) incidences, one per (face vertex, face cell) - a pair may repeat (a point is reached through several
    faces of the same cell); every consumer is a boolean scatter, so duplicates are harmless and the 50 M-entry
    `unique` they would cost at 256^3 is not paid."""
    nv = mesh.face_nverts()
    nI = mesh.n_internal
    p = np.concatenate([mesh.face_verts, mesh.face_verts[:mesh.face_offsets[nI]]])            # int32: 1.6 GB each at 256^3
    c = np.concatenate([np.repeat(mesh.owner, nv), np.repeat(mesh.neighbour, nv[:nI])])
    return p, c


def extended_submeshes(mesh: PolyMesh, cell_rank: np.ndarray, ranks=None) -> List[SubDomain]:
    """Build the extended sub-mesh (+ exchange lists) of every rank in `ranks` (default: all)."""
    n_parts = int(cell_rank.max()) + 1
    ranks = list(range(n_parts)) if ranks is None else list(ranks)
    pp, pc = _point_cells(mesh)                     # (point, cell) incidences (with repeats)
    nI = mesh.n_internal
    nv = mesh.face_nverts()
    out = []
    ext_sets = {}
    for r in range(n_parts):
        owned = cell_rank == r
        pts_owned = np.zeros(mesh.n_points, bool)
        pts_owned[pp[owned[pc]]] = True
        ext = np.zeros(mesh.n_cells, bool)
        ext[pc[pts_owned[pp]]] = True
        ext_sets[r] = (owned, ext, pts_owned)
    for r in ranks:
        owned, ext, pts_owned = ext_sets[r]
        own_ids = np.nonzero(owned)[0]
        halo_ids = np.nonzero(ext & ~owned)[0]
        cell_global = np.concatenate([own_ids, halo_ids]).astype(np.int64)
        n_owned = own_ids.size
        g2l = np.full(mesh.n_cells, -1, np.int64)
        g2l[cell_global] = np.arange(cell_global.size)
        # ---- internal faces: both cells in the extended set
        go, gn = mesh.owner[:nI].astype(np.int64), mesh.neighbour.astype(np.int64)
        keep = ext[go] & ext[gn]
        fi = np.nonzero(keep)[0]
        lo, ln = g2l[go[fi]], g2l[gn[fi]]
        flip = lo > ln                                # the lower local id is the owner (owned cells come first)
        own_l = np.where(flip, ln, lo)
        nei_l = np.where(flip, lo, ln)
        order = np.lexsort((nei_l, own_l))
        fi, flip, own_l, nei_l = fi[order], flip[order], own_l[order], nei_l[order]
        coupled = ((own_l < n_owned) & (nei_l >= n_owned)).astype(np.int32)
        # ---- boundary faces: physical boundary faces of extended cells, per patch; then the cut faces
        bown_g = mesh.owner[nI:].astype(np.int64)
        faces_g, flips, owners_l, patches = [fi], [flip], [own_l], []
        start = fi.size
        for p in mesh.patches:
            ids = np.arange(p.start, p.start + p.size)
            sel = ids[ext[mesh.owner[ids]]]
            ol = g2l[mesh.owner[sel]]
            o2 = np.argsort(ol, kind="stable")
            sel, ol = sel[o2], ol[o2]
            faces_g.append(sel); flips.append(np.zeros(sel.size, bool)); owners_l.append(ol)
            patches.append(Patch(p.name, p.kind, start, sel.size))
            start += sel.size
        cut = np.nonzero(ext[go] ^ ext[gn])[0]        # exactly one side present
        cut_flip = ~ext[go[cut]]
        cut_owner = g2l[np.where(cut_flip, gn[cut], go[cut])]
        o2 = np.argsort(cut_owner, kind="stable")
        faces_g.append(cut[o2]); flips.append(cut_flip[o2]); owners_l.append(cut_owner[o2])
        patches.append(Patch("cutFaces", PATCH_CUT, start, cut.size))
        face_global = np.concatenate(faces_g)
        face_flipped = np.concatenate(flips)
        owner_l = np.concatenate(owners_l).astype(np.int32)
        # ---- points
        n_loc_faces = face_global.size
        nvl = nv[face_global]
        offs = np.zeros(n_loc_faces + 1, np.int64)
        np.cumsum(nvl, out=offs[1:])
        # gather vertex lists (reverse flipped faces keeping the first vertex, like OpenFOAM's reverseFace)
        src = np.repeat(mesh.face_offsets[face_global].astype(np.int64), nvl) + _ragged_arange(nvl)
        verts_g = mesh.face_verts[src].astype(np.int64)
        if face_flipped.any():
            pos = _ragged_arange(nvl)
            cnt = np.repeat(nvl, nvl)
            fl = np.repeat(face_flipped, nvl)
            rev_pos = np.where(pos == 0, 0, cnt - pos)
            src2 = np.repeat(mesh.face_offsets[face_global].astype(np.int64), nvl) + np.where(fl, rev_pos, pos)
            verts_g = mesh.face_verts[src2].astype(np.int64)
        point_global = np.unique(verts_g)
        pg2l = np.full(mesh.n_points, -1, np.int64)
        pg2l[point_global] = np.arange(point_global.size)
        sgn = np.where(face_flipped, -1.0, 1.0)
        w = mesh.weights[face_global].copy()
        w[:fi.size] = np.where(flip, 1.0 - w[:fi.size], w[:fi.size])
        sub = PolyMesh(points=np.ascontiguousarray(mesh.points[point_global]),
                       face_offsets=offs.astype(np.int32), face_verts=pg2l[verts_g].astype(np.int32),
                       owner=owner_l, neighbour=nei_l.astype(np.int32), patches=patches,
                       n_cells=cell_global.size, geometric_d=mesh.geometric_d.copy())
        sub.C = np.ascontiguousarray(mesh.C[cell_global]); sub.V = np.ascontiguousarray(mesh.V[cell_global])
        sub.Cf = np.ascontiguousarray(mesh.Cf[face_global]); sub.Sf = np.ascontiguousarray(mesh.Sf[face_global] * sgn[:, None])
        sub.magSf = np.ascontiguousarray(mesh.magSf[face_global]); sub.weights = w
        sub.deltaCoeffs = np.ascontiguousarray(mesh.deltaCoeffs[face_global])
        sub.nonOrthDeltaCoeffs = np.ascontiguousarray(mesh.nonOrthDeltaCoeffs[face_global])
        # cut faces: deltaCoeffs of an internal face are fine; geometry there is never used
        sub.neighb_cell_centres = np.full((sub.n_bnd, 3), np.nan)
        sd = SubDomain(rank=r, mesh=sub, n_owned=n_owned, cell_global=cell_global, face_global=face_global,
                       face_flipped=face_flipped, point_global=point_global, coupled_face=coupled)
        out.append(sd)
    # ---- exchange lists (need every rank's numbering: cheap maps only)
    l_of = {}
    for r in range(n_parts):
        owned, ext, _ = ext_sets[r]
        own_ids = np.nonzero(owned)[0]
        halo_ids = np.nonzero(ext & ~owned)[0]
        g2l = np.full(mesh.n_cells, -1, np.int64)
        g2l[np.concatenate([own_ids, halo_ids])] = np.arange(own_ids.size + halo_ids.size)
        l_of[r] = (g2l, ext)
    bface_maps = {}

    def bface_map(r):
        """global boundary face -> local boundary-face index on rank r (or -1)."""
        if r not in bface_maps:
            _, ext = l_of[r]
            m = np.full(mesh.n_faces, -1, np.int64)
            k = 0
            for p in mesh.patches:
                ids = np.arange(p.start, p.start + p.size)
                sel = ids[ext[mesh.owner[ids]]]
                ol = l_of[r][0][mesh.owner[sel]]
                sel = sel[np.argsort(ol, kind="stable")]
                m[sel] = k + np.arange(sel.size)
                k += sel.size
            bface_maps[r] = m
        return bface_maps[r]

    for sd in out:
        r = sd.rank
        g2l_r, ext_r = l_of[r]
        halo_g = sd.cell_global[sd.n_owned:]
        for s in np.unique(cell_rank[halo_g]):
            s = int(s)
            cells_g = halo_g[cell_rank[halo_g] == s]                 # ascending global id
            sd.recv_cells[s] = g2l_r[cells_g].astype(np.int32)
            # physical boundary faces owned by those cells (kept on r): receive their boundary state from s
            bm_r = bface_map(r)
            isb = np.zeros(mesh.n_cells, bool); isb[cells_g] = True
            bf_g = np.nonzero(isb[mesh.owner[nI:]])[0] + nI
            bf_g = bf_g[bm_r[bf_g] >= 0]
            sd.recv_bfaces[s] = bm_r[bf_g].astype(np.int32)
        # what r must send: cells of r that are halo on s
        for s in range(n_parts):
            if s == r:
                continue
            _, ext_s = l_of[s]
            mine = np.nonzero((cell_rank == r) & ext_s)[0]
            if mine.size == 0:
                continue
            sd.send_cells[s] = g2l_r[mine].astype(np.int32)
            bm_r = bface_map(r)
            isb = np.zeros(mesh.n_cells, bool); isb[mine] = True
            bf_g = np.nonzero(isb[mesh.owner[nI:]])[0] + nI
            sd.send_bfaces[s] = bm_r[bf_g].astype(np.int32)
        # face-neighbour halo: cells on either side of the faces cut between r and s, ascending global id on both sides
        go_, gn_ = mesh.owner[:nI], mesh.neighbour
        ro_, rn_ = cell_rank[go_], cell_rank[gn_]
        for s in sorted(set(sd.recv_cells) | set(sd.send_cells)):
            a = (ro_ == r) & (rn_ == s)
            b = (ro_ == s) & (rn_ == r)
            theirs = np.unique(np.concatenate([gn_[a], go_[b]]))
            mine = np.unique(np.concatenate([go_[a], gn_[b]]))
            if theirs.size:
                sd.recv_face_cells[s] = g2l_r[theirs].astype(np.int32)
                sd.send_face_cells[s] = g2l_r[mine].astype(np.int32)
    return out


def _ragged_arange(counts: np.ndarray) -> np.ndarray:
    """concatenate([arange(c) for c in counts]) without a Python loop."""
    counts = np.asarray(counts, np.int64)
    total = int(counts.sum())
    starts = np.zeros(counts.size, np.int64)
    np.cumsum(counts[:-1], out=starts[1:])
    return np.arange(total, dtype=np.int64) - np.repeat(starts, counts)


def gather_owned(subs: List[SubDomain], fields: List[np.ndarray], n_cells: int) -> np.ndarray:
    """Reassemble a global cell field from per-rank local fields (owned parts only)."""
    shape = (n_cells,) + fields[0].shape[1:]
    out = np.zeros(shape, fields[0].dtype)
    for sd, f in zip(subs, fields):
        out[sd.cell_global[:sd.n_owned]] = f[:sd.n_owned]
    return out


# ------------------------------------------------------------------------------------------------ decomposePar layout
@dataclass
class ProcMesh:
    """One `processorN/constant/polyMesh` of a decomposePar output with its four addressing lists [OF-v2312
    domainDecomposition::decomposeMesh semantics, restated; the reference consumes these through `-parallel` runs]."""
    rank: int
    mesh: PolyMesh                      # processor mesh: physical patches (same order as the global mesh), then processor patches
    cell_addr: np.ndarray               # cellProcAddressing     local cell  -> global cell
    face_addr: np.ndarray               # faceProcAddressing     local face  -> +-(global face + 1), negative = stored reversed
    point_addr: np.ndarray              # pointProcAddressing    local point -> global point
    boundary_addr: np.ndarray           # boundaryProcAddressing local patch -> global patch, -1 for processor patches


def _reverse_face(v: np.ndarray) -> np.ndarray:
    """face::reverseFace [OF]: first vertex kept, the rest reversed."""
    return np.concatenate([v[:1], v[:0:-1]])


def processor_meshes(mesh: PolyMesh, cell_rank: np.ndarray, ranks=None) -> List[ProcMesh]:
    """The processor meshes decomposePar writes for the cell->processor map `cell_rank`:
    cells and points keep ascending global order; faces = internal faces of the processor (ascending global face id),
    the physical patches in global order (zero-sized ones kept), then one processor patch per neighbour rank
    (ascending) whose faces are in ascending global face id on both sides, reversed on the side of the global neighbour
    cell."""
    cell_rank = np.asarray(cell_rank, np.int32)
    n_parts = int(cell_rank.max()) + 1
    ranks = list(range(n_parts)) if ranks is None else list(ranks)
    nI = mesh.n_internal
    go, gn = mesh.owner[:nI], mesh.neighbour
    ro, rn = cell_rank[go], cell_rank[gn]
    out = []
    for r in ranks:
        cells = np.nonzero(cell_rank == r)[0]
        g2l = np.full(mesh.n_cells, -1, np.int64)
        g2l[cells] = np.arange(cells.size)
        f_int = np.nonzero((ro == r) & (rn == r))[0]
        faces = [f_int + 1]
        owners = [g2l[go[f_int]]]
        neigh = g2l[gn[f_int]]
        patches, baddr = [], []
        start = f_int.size
        for pi, p in enumerate(mesh.patches):
            ids = np.arange(p.start, p.start + p.size)
            sel = ids[cell_rank[mesh.owner[ids]] == r]
            faces.append(sel + 1)
            owners.append(g2l[mesh.owner[sel]])
            patches.append(Patch(p.name, p.kind, start, sel.size, p.neighb_rank))
            baddr.append(pi)
            start += sel.size
        as_owner = (ro == r) & (rn != r)
        as_neigh = (rn == r) & (ro != r)
        nbrs = np.unique(np.concatenate([rn[as_owner], ro[as_neigh]]))
        for q in nbrs:
            fo = np.nonzero(as_owner & (rn == q))[0]
            fn = np.nonzero(as_neigh & (ro == q))[0]
            ids = np.concatenate([fo, fn])
            sign = np.concatenate([np.ones(fo.size, np.int64), -np.ones(fn.size, np.int64)])
            loc = np.concatenate([g2l[go[fo]], g2l[gn[fn]]])
            o = np.argsort(ids, kind="stable")
            faces.append(sign[o] * (ids[o] + 1))
            owners.append(loc[o])
            patches.append(Patch(f"procBoundary{r}to{int(q)}", PATCH_PROCESSOR, start, ids.size, int(q)))
            baddr.append(-1)
            start += ids.size
        face_addr = np.concatenate(faces).astype(np.int64)
        owner = np.concatenate(owners).astype(np.int32)
        gf = np.abs(face_addr) - 1
        used = np.zeros(mesh.n_points, bool)
        for f in gf:
            used[mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]] = True
        pts = np.nonzero(used)[0]
        p2l = np.full(mesh.n_points, -1, np.int64)
        p2l[pts] = np.arange(pts.size)
        offs, verts = [0], []
        for f, a in zip(gf, face_addr):
            v = mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]
            if a < 0:
                v = _reverse_face(v)
            verts.append(p2l[v])
            offs.append(offs[-1] + v.size)
        pm = PolyMesh(points=np.ascontiguousarray(mesh.points[pts]), face_offsets=np.asarray(offs, np.int32),
                      face_verts=np.concatenate(verts).astype(np.int32) if verts else np.zeros(0, np.int32), owner=owner,
                      neighbour=neigh.astype(np.int32), patches=patches, n_cells=int(cells.size),
                      geometric_d=mesh.geometric_d.copy())
        out.append(ProcMesh(r, pm, cells.astype(np.int32), face_addr.astype(np.int32), pts.astype(np.int32),
                            np.asarray(baddr, np.int32)))
    return out


def cell_rank_from_procs(procs: List[ProcMesh], n_cells: int) -> np.ndarray:
    """cell -> processor map recovered from the cellProcAddressing lists (what `decomposePar -cellDist` writes as
    constant/cellDecomposition); checks that the lists partition the mesh."""
    rank = np.full(n_cells, -1, np.int32)
    for p in procs:
        if (rank[p.cell_addr] != -1).any():
            raise ValueError("cellProcAddressing lists overlap")
        rank[p.cell_addr] = p.rank
    if (rank < 0).any():
        raise ValueError("cellProcAddressing lists do not cover the mesh")
    return rank


def check_processor_mesh(mesh: PolyMesh, p: ProcMesh) -> None:
    """Bit-exact consistency of a processor mesh with the undecomposed mesh through its addressing lists."""
    pm = p.mesh
    if not np.array_equal(pm.points, mesh.points[p.point_addr]):
        raise ValueError("pointProcAddressing: coordinates differ")
    gf = np.abs(p.face_addr.astype(np.int64)) - 1
    for lf in range(pm.n_faces):
        v = p.point_addr[pm.face_verts[pm.face_offsets[lf]:pm.face_offsets[lf + 1]]]
        g = mesh.face_verts[mesh.face_offsets[gf[lf]]:mesh.face_offsets[gf[lf] + 1]]
        if p.face_addr[lf] < 0:
            g = _reverse_face(g)
        if not np.array_equal(v, g):
            raise ValueError(f"faceProcAddressing: local face {lf} does not match global face {gf[lf]}")
    nIl = pm.n_internal
    flipped = p.face_addr < 0
    own_g = np.where(flipped, mesh.neighbour[np.minimum(gf, mesh.n_internal - 1)] if mesh.n_internal else 0, mesh.owner[gf])
    if not np.array_equal(p.cell_addr[pm.owner], own_g):
        raise ValueError("owner does not map through cellProcAddressing")
    if flipped[:nIl].any() or not np.array_equal(p.cell_addr[pm.neighbour], mesh.neighbour[gf[:nIl]]):
        raise ValueError("neighbour does not map through cellProcAddressing")
    if nIl and not (pm.owner[:nIl] < pm.neighbour).all():
        raise ValueError("processor mesh is not in owner < neighbour order")
    for lp, patch in enumerate(pm.patches):
        ga = p.boundary_addr[lp]
        fs = gf[patch.start:patch.start + patch.size]
        if ga >= 0:
            gp = mesh.patches[ga]
            if patch.name != gp.name or ((fs < gp.start) | (fs >= gp.start + gp.size)).any():
                raise ValueError(f"patch {patch.name}: faces outside the global patch")
        elif (fs >= mesh.n_internal).any():
            raise ValueError(f"processor patch {patch.name} contains a global boundary face")


def couple_processor_geometry(procs: List[ProcMesh]) -> None:
    """Fill the coupled-patch geometry of every processor patch from the neighbour processor's mesh
    [OF-v2312 processorFvPatch / coupledFvPatch]: neighbFaceCellCentres, weights
    |Sf.(Cn-Cf)| / (|Sf.(Cf-Cp)| + |Sf.(Cn-Cf)|), deltaCoeffs 1/|Cn-Cp|, nonOrthDeltaCoeffs 1/max(nf.(Cn-Cp), 0.05|Cn-Cp|)."""
    for p in procs:
        if p.mesh.C is None:
            p.mesh.compute_geometry()
    by_rank = {p.rank: p for p in procs}
    for p in procs:
        m = p.mesh
        nI = m.n_internal
        for patch in m.patches:
            if patch.kind != PATCH_PROCESSOR or patch.size == 0:
                continue
            q = by_rank[patch.neighb_rank]
            other = [x for x in q.mesh.patches if x.kind == PATCH_PROCESSOR and x.neighb_rank == p.rank][0]
            sl = slice(patch.start, patch.start + patch.size)
            Cn = q.mesh.C[q.mesh.owner[other.start:other.start + other.size]]
            Cp = m.C[m.owner[sl]]
            m.neighb_cell_centres[patch.start - nI:patch.start - nI + patch.size] = Cn
            so = np.abs((m.Sf[sl] * (m.Cf[sl] - Cp)).sum(1))
            sn = np.abs((m.Sf[sl] * (Cn - m.Cf[sl])).sum(1))
            m.weights[sl] = sn / (so + sn)
            d = Cn - Cp
            md = np.sqrt((d * d).sum(1))
            m.deltaCoeffs[sl] = 1.0 / md
            nf = m.Sf[sl] / m.magSf[sl, None]
            m.nonOrthDeltaCoeffs[sl] = 1.0 / np.maximum((nf * d).sum(1), 0.05 * md)
