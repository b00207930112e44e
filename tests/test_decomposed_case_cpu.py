"""decomposePar-layout processor meshes (SURVEY 8e "Partitioning", 8f rank 2): addressing is bit-exact, the on-disk
round trip is lossless, the cell->processor map that shards the multi-GPU run is recovered from the case, and the
oracle's processor-patch code path reproduces the serial operators face by face."""
import numpy as np
import pytest

import cases
from qgdsolver_b200 import decompose, foamcase


def _irregular_partition(mesh, n_parts, seed=4):
    """scotch-like irregular but connected-ish partition: nearest of n_parts random seeds (plus a few swapped cells)."""
    rng = np.random.default_rng(seed)
    seeds = mesh.C[rng.choice(mesh.n_cells, n_parts, replace=False)]
    rank = np.argmin(((mesh.C[:, None, :] - seeds[None]) ** 2).sum(2), axis=1).astype(np.int32)
    return rank


MESHES = {
    "hex": lambda: cases.pm.hex_box(6, 5, 4, perturb=0.2, seed=9),
    "prism": lambda: cases.pm.prism_box(4, 4, 3, perturb=0.1, seed=2),
    "poly": lambda: cases.pm.hexprism_poly(5, 4, 3, a=0.1, lz=0.4),
    "2d": lambda: cases.case_2d((10, 8), perturb=0.1).mesh,
}


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("n_parts", [2, 4, 5])
def test_processor_meshes_are_bit_exact_views_of_the_global_mesh(name, n_parts):
    mesh = MESHES[name]()
    rank = decompose.geometric_split(mesh, n_parts) if n_parts in (2, 4) else _irregular_partition(mesh, n_parts)
    procs = decompose.processor_meshes(mesh, rank)
    assert len(procs) == int(rank.max()) + 1
    assert np.array_equal(decompose.cell_rank_from_procs(procs, mesh.n_cells), rank)
    seen = np.zeros(mesh.n_faces, int)
    for p in procs:
        decompose.check_processor_mesh(mesh, p)
        assert np.array_equal(p.cell_addr, np.nonzero(rank == p.rank)[0])          # ascending global order
        assert (np.diff(p.point_addr) > 0).all()
        np.add.at(seen, np.abs(p.face_addr) - 1, 1)
        # same owned-cell order as the device sub-domain: per-rank results map through cellProcAddressing directly
        sub = decompose.extended_submeshes(mesh, rank, ranks=[p.rank])[0]
        assert np.array_equal(sub.cell_global[:sub.n_owned], p.cell_addr)
    nI = mesh.n_internal
    cut = rank[mesh.owner[:nI]] != rank[mesh.neighbour]
    assert np.array_equal(seen[:nI], 1 + cut.astype(int))       # cut faces appear once on each side
    assert (seen[nI:] == 1).all()
    # processor patches pair up: same global faces in the same order, opposite orientation
    for p in procs:
        for patch in p.mesh.patches:
            if patch.kind != cases.pm.PATCH_PROCESSOR:
                continue
            q = procs[patch.neighb_rank]
            other = [x for x in q.mesh.patches if x.kind == cases.pm.PATCH_PROCESSOR and x.neighb_rank == p.rank][0]
            a = p.face_addr[patch.start:patch.start + patch.size]
            b = q.face_addr[other.start:other.start + other.size]
            assert np.array_equal(a, -b)


def test_decomposed_case_round_trips_through_disk(tmp_path):
    mesh = cases.pm.hex_box(5, 4, 3, perturb=0.15, seed=3)
    rank = _irregular_partition(mesh, 3)
    written = foamcase.write_decomposed_case(mesh, rank, str(tmp_path))
    read = foamcase.read_decomposed_case(str(tmp_path))
    assert len(read) == len(written) == 3
    for w, r in zip(written, read):
        for attr in ("cell_addr", "face_addr", "point_addr", "boundary_addr"):
            assert np.array_equal(getattr(w, attr), getattr(r, attr)), attr
        assert np.array_equal(w.mesh.points, r.mesh.points)                        # repr() floats: bit-exact
        for attr in ("face_offsets", "face_verts", "owner", "neighbour"):
            assert np.array_equal(getattr(w.mesh, attr), getattr(r.mesh, attr)), attr
        assert [(p.name, p.kind, p.start, p.size, p.neighb_rank) for p in w.mesh.patches] == \
               [(p.name, p.kind, p.start, p.size, p.neighb_rank) for p in r.mesh.patches]
        decompose.check_processor_mesh(mesh, r)
    assert np.array_equal(foamcase.read_cell_decomposition(str(tmp_path), mesh.n_cells), rank)
    import os
    os.remove(os.path.join(str(tmp_path), "constant", "cellDecomposition"))        # fall back to cellProcAddressing
    assert np.array_equal(foamcase.read_cell_decomposition(str(tmp_path), mesh.n_cells), rank)


def test_corrupt_addressing_is_detected():
    mesh = cases.pm.hex_box(4, 3, 3)
    procs = decompose.processor_meshes(mesh, decompose.geometric_split(mesh, 2))
    p = procs[0]
    p.face_addr = p.face_addr.copy()
    p.face_addr[0], p.face_addr[1] = p.face_addr[1], p.face_addr[0]
    with pytest.raises(ValueError):
        decompose.check_processor_mesh(mesh, p)
    q = procs[1]
    q.cell_addr = q.cell_addr.copy()
    q.cell_addr[0] = procs[0].cell_addr[0]
    with pytest.raises(ValueError):
        decompose.cell_rank_from_procs(procs, mesh.n_cells)


def test_processor_fields_scatter_and_gather(tmp_path):
    c = cases.case_hex3d(n=(5, 4, 3), bcs="fixed")
    mesh = c.mesh
    rank = decompose.geometric_split(mesh, 2)
    procs = foamcase.write_decomposed_case(mesh, rank, str(tmp_path))
    T = np.sin(mesh.C[:, 0] * 3) + mesh.C[:, 1]
    U = np.stack([T, 2 * T, -T], 1)
    types = {p.name: "fixedValue" for p in mesh.patches}
    foamcase.write_processor_fields(str(tmp_path), "0", "T", procs, T, types, c.bvT, mesh.n_internal)
    foamcase.write_processor_fields(str(tmp_path), "0", "U", procs, U, types, c.bvU, mesh.n_internal)
    back_T, back_U = np.zeros_like(T), np.zeros_like(U)
    import os
    for p in foamcase.read_decomposed_case(str(tmp_path)):
        fT = foamcase.read_field(os.path.join(str(tmp_path), f"processor{p.rank}", "0", "T"), p.mesh)
        fU = foamcase.read_field(os.path.join(str(tmp_path), f"processor{p.rank}", "0", "U"), p.mesh)
        back_T[p.cell_addr] = fT.internal
        back_U[p.cell_addr] = fU.internal
        nIl = p.mesh.n_internal
        for patch in p.mesh.patches:
            if patch.kind == cases.pm.PATCH_PROCESSOR:
                assert fT.patch_types[patch.name] == "processor"
                continue
            gf = np.abs(p.face_addr[patch.start:patch.start + patch.size]) - 1 - mesh.n_internal
            assert np.array_equal(fT.patch_values[patch.name], c.bvT[gf])
            assert np.array_equal(fU.patch_values[patch.name], c.bvU[gf])
        del nIl
    assert np.array_equal(back_T, T) and np.array_equal(back_U, U)


@pytest.mark.parametrize("name", ["hex", "poly"])
def test_oracle_processor_patch_path_reproduces_the_serial_operators(oracle_mod, name):
    """fvsc `reduced` (nf*snGrad) on the decomposePar processor meshes, fed with patchNeighbourField values, equals the
    serial operator on every face: pins the oracle's coupled-patch branches (linearInterpolate, snGrad, hQGDf rule)."""
    mesh = MESHES[name]()
    rank = _irregular_partition(mesh, 3, seed=7)
    procs = decompose.processor_meshes(mesh, rank)
    decompose.couple_processor_geometry(procs)
    nI = mesh.n_internal
    phi = np.sin(3 * mesh.C[:, 0]) + mesh.C[:, 1] ** 2 - 0.5 * mesh.C[:, 2]
    vec = np.stack([phi, np.cos(2 * mesh.C[:, 1]), phi * mesh.C[:, 0]], 1)
    bnd = np.cos(2 * mesh.Cf[nI:, 1]) + mesh.Cf[nI:, 0]
    bvec = np.stack([bnd, 2 * bnd, -bnd], 1)
    sch = oracle_mod.FVSC_SCHEMES["reduced"]
    og = oracle_mod.Oracle(mesh)
    bsg = mesh.deltaCoeffs[nI:] * (bnd - phi[mesh.owner[nI:]])
    bsgv = mesh.deltaCoeffs[nI:, None] * (bvec - vec[mesh.owner[nI:]])
    ref_g = og.fvsc_grad(phi, bnd, bsg, scheme=sch)
    ref_d = og.fvsc_div(vec, bvec, bsgv, scheme=sch)
    for p in procs:
        m = p.mesh
        nIl = m.n_internal
        gf = np.abs(p.face_addr.astype(np.int64)) - 1
        gb = gf[nIl:]
        phys = gb >= nI
        lphi, lvec = phi[p.cell_addr], vec[p.cell_addr]
        # boundary values: physical faces from the global boundary field; processor faces carry the neighbour cell value
        across = np.where(p.face_addr[nIl:] < 0, mesh.owner[np.minimum(gb, nI - 1)], mesh.neighbour[np.minimum(gb, nI - 1)])
        lb = np.where(phys, bnd[np.where(phys, gb - nI, 0)], 0.0)
        lbv = np.where(phys[:, None], bvec[np.where(phys, gb - nI, 0)], 0.0)
        nbr = np.where(phys, 0.0, phi[across])
        nbrv = np.where(phys[:, None], 0.0, vec[across])
        # patch snGrad: deltaCoeffs on ordinary patches; coupled patches get the snGrad scheme's nonOrthDeltaCoeffs [OF fvc::snGrad]
        dc = np.where(phys, m.deltaCoeffs[nIl:], m.nonOrthDeltaCoeffs[nIl:])
        lsg = dc * (np.where(phys, lb, nbr) - lphi[m.owner[nIl:]])
        lsgv = dc[:, None] * (np.where(phys[:, None], lbv, nbrv) - lvec[m.owner[nIl:]])
        o = oracle_mod.Oracle(m)
        g = o.fvsc_grad(lphi, lb, lsg, scheme=sch, nbr=nbr)
        d = o.fvsc_div(lvec, lbv, lsgv, scheme=sch, nbr=nbrv)
        assert np.abs(g - ref_g[gf]).max() < 1e-11 * np.abs(ref_g).max()
        assert np.abs(d - ref_d[gf]).max() < 1e-11 * np.abs(ref_d).max()


def test_decomposed_case_on_disk_drives_the_device_sharding(tmp_path):
    """A decomposePar'd case on disk -> cell->processor map -> the per-GPU extended sub-meshes: every rank owns exactly
    its processorN cells in cellProcAddressing order, halo cells are owned by the rank that sends them, and the
    send / receive lists of each pair of ranks name the same global cells in the same order."""
    mesh = cases.pm.hex_box(7, 6, 5, perturb=0.1, seed=8)
    rank = _irregular_partition(mesh, 4, seed=11)
    foamcase.write_polymesh(mesh, str(tmp_path))
    foamcase.write_decomposed_case(mesh, rank, str(tmp_path))
    g = foamcase.read_polymesh(str(tmp_path))
    cell_rank = foamcase.read_cell_decomposition(str(tmp_path), g.n_cells)
    assert np.array_equal(cell_rank, rank)
    procs = foamcase.read_decomposed_case(str(tmp_path))
    subs = decompose.extended_submeshes(g, cell_rank)
    for p, sub in zip(procs, subs):
        assert np.array_equal(sub.cell_global[:sub.n_owned], p.cell_addr)
        halo = sub.cell_global[sub.n_owned:]
        assert (cell_rank[halo] != p.rank).all()
        for q, ids in sub.recv_cells.items():
            assert (cell_rank[sub.cell_global[ids]] == q).all()
            other = subs[q]
            assert np.array_equal(other.cell_global[other.send_cells[p.rank]], sub.cell_global[ids])
    # owned results gathered through the addressing reproduce a global field bit for bit
    fld = np.sin(g.C[:, 0] * 5) + g.C[:, 2]
    back = decompose.gather_owned(subs, [fld[s.cell_global[:s.n_owned]] for s in subs], g.n_cells)
    assert np.array_equal(back, fld)


def test_decompose_and_reconstruct_command_line(tmp_path):
    """python -m qgdsolver_b200.decompose_case: case -> processor directories with the start-time fields; after copying the
    processor fields to a new time, -reconstruct gives back the original fields bit for bit (gradient-type patches keep no value)."""
    import os
    import shutil
    from qgdsolver_b200 import decompose_case as dc
    c = cases.case_hex3d(n=(6, 5, 4), perturb=0.1, bcs="fixed")
    m = c.mesh
    foamcase.write_polymesh(m, str(tmp_path))
    fv = {p.name: "fixedValue" for p in m.patches}
    zg = {p.name: "zeroGradient" for p in m.patches}
    foamcase.write_field(str(tmp_path / "0" / "U"), m, "U", c.U0, fv, c.bvU)
    foamcase.write_field(str(tmp_path / "0" / "T"), m, "T", c.T0, zg)
    assert dc.main([str(tmp_path), "-n", "4"]) == 0
    procs = foamcase.read_decomposed_case(str(tmp_path))
    assert len(procs) == 4 and os.path.exists(tmp_path / "constant" / "cellDecomposition")
    for p in procs:
        shutil.copytree(tmp_path / f"processor{p.rank}" / "0", tmp_path / f"processor{p.rank}" / "0.5")
        f = foamcase.read_field(str(tmp_path / f"processor{p.rank}" / "0" / "U"), p.mesh)
        assert np.array_equal(f.internal, c.U0[p.cell_addr])
    assert dc.main([str(tmp_path), "-reconstruct", "0.5"]) == 0
    U = foamcase.read_field(str(tmp_path / "0.5" / "U"), m)
    T = foamcase.read_field(str(tmp_path / "0.5" / "T"), m)
    assert np.array_equal(U.internal, c.U0) and np.array_equal(T.internal, c.T0)
    nI = m.n_internal
    for p in m.patches:
        assert U.patch_types[p.name] == "fixedValue" and np.array_equal(U.patch_values[p.name], c.bvU[p.start - nI:p.start - nI + p.size])
        assert T.patch_types[p.name] == "zeroGradient"
