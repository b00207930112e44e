"""OpenFOAM ASCII case ingestion / write-back (qgdsolver_b200/foamcase.py): round trips on synthetic meshes and a
hand-written tutorial-style case (the reference repository ships no case files)."""
import os

import numpy as np
import pytest

import cases
from qgdsolver_b200 import foamcase as fc


@pytest.mark.parametrize("mk", [lambda: cases.pm.hex_box(5, 4, 3, perturb=0.2, seed=1),
                                lambda: cases.pm.prism_box(3, 3, 2, perturb=0.1, seed=2),
                                lambda: cases.pm.hexprism_poly(4, 3, 2, a=0.1, lz=0.3),
                                lambda: cases.case_2d((6, 5)).mesh, lambda: cases.case_sod(12).mesh])
def test_polymesh_round_trip_is_bit_exact(tmp_path, mk):
    m = mk()
    fc.write_polymesh(m, str(tmp_path))
    r = fc.read_polymesh(str(tmp_path))
    assert r.n_cells == m.n_cells and len(r.patches) == len(m.patches)
    for a in ("points", "face_offsets", "face_verts", "owner", "neighbour"):
        assert np.array_equal(getattr(r, a), getattr(m, a)), a          # repr() floats round-trip exactly
    for p, q in zip(r.patches, m.patches):
        assert (p.name, p.kind, p.start, p.size) == (q.name, q.kind, q.start, q.size)
    for a in ("C", "V", "Cf", "Sf", "weights", "deltaCoeffs"):
        assert np.array_equal(getattr(r, a), getattr(m, a)), a
    assert np.array_equal(r.geometric_d, m.geometric_d)


def test_field_round_trip_and_bc_arrays(tmp_path):
    c = cases.case_hex3d(n=(4, 3, 3), bcs="mixed")
    m = c.mesh
    names = [p.name for p in m.patches]
    code = {0: "fixedValue", 1: "zeroGradient", 3: "qgdFlux"}
    ub = np.zeros((m.n_bnd, 3))
    fc.write_field(str(tmp_path / "0" / "U"), m, "U", c.U0, {n: code[int(k)] for n, k in zip(names, c.bcU)}, ub, "[0 1 -1 0 0 0 0]")
    fc.write_field(str(tmp_path / "0" / "p"), m, "p", c.p0, {n: code[int(k)] for n, k in zip(names, c.bcP)}, np.full(m.n_bnd, 0.7))
    U = fc.read_field(str(tmp_path / "0" / "U"), m)
    p = fc.read_field(str(tmp_path / "0" / "p"), m)
    assert U.ncmpt == 3 and np.array_equal(U.internal, c.U0) and np.array_equal(p.internal, c.p0)
    kU, vU = fc.bc_arrays(m, U)
    kP, vP = fc.bc_arrays(m, p)
    assert np.array_equal(kU, c.bcU) and np.array_equal(kP, c.bcP)
    assert vU.shape == (m.n_bnd, 3) and vP.shape == (m.n_bnd,)


SOD_P = """FoamFile { version 2.0; format ascii; class volScalarField; object p; }
dimensions [1 -1 -2 0 0 0 0];
internalField nonuniform List<scalar> 4 ( 1 1 0.1 0.1 );   // left / right state
boundaryField
{
    xMin { type zeroGradient; }
    xMax { type fixedValue; value uniform 0.1; }
    yMin { type empty; } yMax { type empty; } zMin { type empty; } zMax { type empty; }
}
"""
SOD_U = """FoamFile { version 2.0; format ascii; class volVectorField; object U; }
dimensions [0 1 -1 0 0 0 0];
internalField uniform (0 0 0);
boundaryField
{
    xMin { type fixedValue; value uniform (0.5 0 0); }
    xMax { type fixedGradient; gradient uniform (0 0 0); }
    yMin { type empty; } yMax { type empty; } zMin { type empty; } zMax { type empty; }
}
"""


def test_hand_written_case_files(tmp_path):
    m = cases.pm.hex_box(4, 1, 1, lengths=(1.0, 0.1, 0.1),
                         patch_kinds={"zMin": "empty", "zMax": "empty", "yMin": "empty", "yMax": "empty"})
    fc.write_polymesh(m, str(tmp_path))
    os.makedirs(tmp_path / "0")
    (tmp_path / "0" / "p").write_text(SOD_P)
    (tmp_path / "0" / "U").write_text(SOD_U)
    r = fc.read_polymesh(str(tmp_path))
    assert list(r.geometric_d) == [1, -1, -1]
    p = fc.read_field(str(tmp_path / "0" / "p"), r)
    U = fc.read_field(str(tmp_path / "0" / "U"), r)
    assert np.array_equal(p.internal, [1, 1, 0.1, 0.1]) and U.internal.shape == (4, 3) and not U.internal.any()
    kP, vP = fc.bc_arrays(r, p)
    kU, vU = fc.bc_arrays(r, U)
    pid = r.patch_id_per_bface()
    names = [q.name for q in r.patches]
    assert kP[names.index("xMin")] == 1 and kP[names.index("xMax")] == 0 and vP[pid == names.index("xMax")][0] == 0.1
    assert kU[names.index("xMin")] == 0 and np.array_equal(vU[pid == names.index("xMin")][0], [0.5, 0, 0])
    assert kU[names.index("xMax")] == 2


def test_rejects_what_it_cannot_read(tmp_path):
    (tmp_path / "x").write_text("FoamFile { format binary; class labelList; }\n3(1 2 3)")
    with pytest.raises(fc.FoamFormatError):
        fc.read_labels(str(tmp_path / "x"))
    m = cases.pm.hex_box(2, 2, 2)
    bad = SOD_P.replace("zeroGradient", "totalPressure")
    (tmp_path / "p").write_text(bad.replace("4 ( 1 1 0.1 0.1 )", "8 ( 1 1 1 1 1 1 1 1 )"))
    f = fc.read_field(str(tmp_path / "p"), m)
    with pytest.raises(fc.FoamFormatError):
        fc.bc_arrays(m, f)


def test_binary_format_round_trip_equals_ascii(tmp_path):
    """`format binary` polyMesh (raw points, faceCompactList faces, raw labels) and fields (raw nonuniform lists) read
    back to the same arrays as the ASCII files [OF-v2312 binary list layout: N newline ( raw bytes )]."""
    c = cases.case_hex3d(n=(5, 4, 3), perturb=0.15, bcs="fixed")
    m = c.mesh
    a_dir, b_dir = tmp_path / "ascii", tmp_path / "binary"
    fc.write_polymesh(m, str(a_dir))
    fc.write_polymesh(m, str(b_dir), binary=True)
    assert b"format      binary" in open(b_dir / "constant" / "polyMesh" / "points", "rb").read(400)
    ma, mb = fc.read_polymesh(str(a_dir)), fc.read_polymesh(str(b_dir))
    for attr in ("points", "face_offsets", "face_verts", "owner", "neighbour", "C", "V", "Sf", "weights"):
        assert np.array_equal(getattr(ma, attr), getattr(mb, attr)), attr
        assert np.array_equal(getattr(mb, attr), getattr(m, attr)) or attr in ("C", "V", "Sf", "weights"), attr
    types = {p.name: "fixedValue" for p in m.patches}
    for d, binary in ((a_dir, False), (b_dir, True)):
        fc.write_field(str(d / "0" / "U"), m, "U", c.U0, types, c.bvU, binary=binary)
        fc.write_field(str(d / "0" / "T"), m, "T", c.T0, types, c.bvT, binary=binary)
    for name, ref, bref in (("U", c.U0, c.bvU), ("T", c.T0, c.bvT)):
        fa, fb = fc.read_field(str(a_dir / "0" / name), ma), fc.read_field(str(b_dir / "0" / name), mb)
        assert np.array_equal(fa.internal, ref) and np.array_equal(fb.internal, ref)
        nI = m.n_internal
        for p in m.patches:
            assert np.array_equal(fb.patch_values[p.name], bref[p.start - nI:p.start - nI + p.size])
            assert fa.patch_types == fb.patch_types
    # the raw bytes on disk are the array itself
    raw = open(b_dir / "constant" / "polyMesh" / "owner", "rb").read()
    assert m.owner.astype("<i4").tobytes() in raw


def test_binary_errors(tmp_path):
    m = cases.pm.hex_box(2, 2, 2)
    fc.write_polymesh(m, str(tmp_path), binary=True)
    p = tmp_path / "constant" / "polyMesh" / "points"
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-40])                               # truncated raw block
    with pytest.raises(fc.FoamFormatError):
        fc.read_points(str(p))
    open(p, "wb").write(data.replace(b"LSB", b"MSB"))
    with pytest.raises(fc.FoamFormatError):
        fc.read_points(str(p))


def test_tutorial_idioms_one_line_lists_and_internalfield_macro(tmp_path):
    """what OpenFOAM itself writes / tutorials contain: short lists on one line `4(0 1 2 3)`, `inGroups 1(wall);` in the
    boundary file, `value $internalField;`, extra header keys (arch, note)."""
    m = cases.pm.hex_box(2, 1, 1, lengths=(1.0, 0.1, 0.1), patch_kinds={"zMin": "empty", "zMax": "empty", "yMin": "empty", "yMax": "empty"})
    fc.write_polymesh(m, str(tmp_path))
    d = tmp_path / "constant" / "polyMesh"
    hdr = 'FoamFile\n{\n    version 2.0;\n    format ascii;\n    arch "LSB;label=32;scalar=64";\n    note "nPoints:12";\n    class labelList;\n    object owner;\n}\n'
    (d / "owner").write_text(hdr + f"{m.owner.size}(" + " ".join(str(int(v)) for v in m.owner) + ")\n")
    b = (d / "boundary").read_text().replace("type            patch;", "type            wall;\n        inGroups        1(wall);")
    (d / "boundary").write_text(b)
    r = fc.read_polymesh(str(tmp_path))
    assert np.array_equal(r.owner, m.owner) and [p.kind for p in r.patches] == [p.kind for p in m.patches]
    os.makedirs(tmp_path / "0", exist_ok=True)
    (tmp_path / "0" / "T").write_text(
        "FoamFile { version 2.0; format ascii; class volScalarField; object T; }\ndimensions [0 0 0 1 0 0 0];\n"
        "internalField uniform 300;\nboundaryField\n{\n    xMin { type fixedValue; value $internalField; }\n"
        "    xMax { type zeroGradient; }\n    yMin { type empty; } yMax { type empty; } zMin { type empty; } zMax { type empty; }\n}\n")
    T = fc.read_field(str(tmp_path / "0" / "T"), r)
    assert np.array_equal(T.internal, [300.0, 300.0]) and np.array_equal(T.patch_values["xMin"], [300.0])
    (tmp_path / "0" / "T").write_text((tmp_path / "0" / "T").read_text().replace("uniform 300", "nonuniform List<scalar> 2(300 301)"))
    with pytest.raises(fc.FoamFormatError):
        fc.read_field(str(tmp_path / "0" / "T"), r)


def test_subset_mesh_and_forward_step_geometry():
    """polymesh.subset_mesh (subsetMesh): kept cells ascending, exposed faces in a new last patch with outward normals, closed
    cells, upper-triangular internal faces; forward_step = 3 x 1 channel minus the 2.4 x 0.2 step."""
    m = cases.pm.forward_step(20)
    nI = m.n_internal
    assert m.n_cells == 60 * 20 - 48 * 4 and [p.name for p in m.patches] == ["xMin", "xMax", "yMin", "yMax", "zMin", "zMax", "step"]
    assert m.patches[-1].size == 48 + 4 and m.patches[1].size == 16 and m.patches[2].size == 12
    assert abs(m.V.sum() - (3.0 - 2.4 * 0.2) * 0.05) < 1e-14
    s = np.zeros((m.n_cells, 3))
    for k in range(3):
        s[:, k] = np.bincount(m.owner, weights=m.Sf[:, k], minlength=m.n_cells) - np.bincount(m.neighbour, weights=m.Sf[:nI, k], minlength=m.n_cells)
    assert np.abs(s).max() < 1e-15
    assert (m.owner[:nI] < m.neighbour).all() and (np.diff(m.owner[:nI].astype(np.int64) * m.n_cells + m.neighbour) > 0).all()
    assert ((m.Sf[nI:] * (m.Cf[nI:] - m.C[m.owner[nI:]])).sum(1) > 0).all()
    step = m.patches[-1]
    cf = m.Cf[step.start:step.start + step.size]
    assert (np.isclose(cf[:, 1], 0.2) | np.isclose(cf[:, 0], 0.6)).all()
    # generic use on a 3D perturbed mesh: remove a random third of the cells, mesh stays closed
    h = cases.pm.hex_box(6, 5, 4, perturb=0.2, seed=1)
    keep = np.random.default_rng(0).random(h.n_cells) > 0.33
    sub = cases.pm.subset_mesh(h, keep)
    assert sub.n_cells == int(keep.sum()) and np.allclose(np.sort(sub.V), np.sort(h.V[keep]), rtol=1e-13)
    nIs = sub.n_internal
    s = np.zeros((sub.n_cells, 3))
    for k in range(3):
        s[:, k] = np.bincount(sub.owner, weights=sub.Sf[:, k], minlength=sub.n_cells) - np.bincount(sub.neighbour, weights=sub.Sf[:nIs, k], minlength=sub.n_cells)
    assert np.abs(s).max() < 1e-14


def test_truncated_octahedron_polyhedral_mesh():
    """polymesh.truncated_octahedron_box: BCC lattice of truncated octahedra, 14 faces per cell (8 hexagons + 6 squares),
    closed cells of volume h^3/2, upper-triangular internal faces, outward boundary normals, every point shared by 4 cells inside."""
    h = 0.25
    m = cases.pm.truncated_octahedron_box(5, 4, 3, h=h)
    nI = m.n_internal
    assert m.n_cells == 5 * 4 * 3 + 4 * 3 * 2
    per_cell = np.bincount(m.owner, minlength=m.n_cells) + np.bincount(m.neighbour, minlength=m.n_cells)
    assert (per_cell == 14).all() and set(np.unique(m.face_nverts())) == {4, 6}
    assert np.abs(m.V - h ** 3 / 2).max() < 1e-15
    s = np.zeros((m.n_cells, 3))
    for k in range(3):
        s[:, k] = np.bincount(m.owner, weights=m.Sf[:, k], minlength=m.n_cells) - np.bincount(m.neighbour, weights=m.Sf[:nI, k], minlength=m.n_cells)
    assert np.abs(s).max() < 1e-15
    assert (m.owner[:nI] < m.neighbour).all() and (np.diff(m.owner[:nI].astype(np.int64) * m.n_cells + m.neighbour) > 0).all()
    assert ((m.Sf[nI:] * (m.Cf[nI:] - m.C[m.owner[nI:]])).sum(1) > 0).all()
    assert ((m.Sf[:nI] * (m.C[m.neighbour] - m.C[m.owner[:nI]])).sum(1) > 0).all()
    assert sum(p.size for p in m.patches) == m.n_bnd and all(p.size > 0 for p in m.patches)
    big = cases.pm.truncated_octahedron_box(12, 12, 12)
    assert 6.0 < big.n_internal / big.n_cells < 7.0                                  # ~7 internal faces per cell at scale
    # round trip through the OpenFOAM files
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        fc.write_polymesh(m, tmp)
        r = fc.read_polymesh(tmp)
        assert np.array_equal(r.face_verts, m.face_verts) and np.array_equal(r.points, m.points) and np.array_equal(r.V, m.V)


def test_fixed_gradient_round_trip_and_uniform_lists(tmp_path):
    """write -> read -> bc_arrays keeps the gradient of fixedGradient / qhdFlux patches (a restart must not continue with
    gradient 0); a fixedGradient patch without `gradient` is an error on both sides; `N{v}` uniform lists are read"""
    import numpy as np
    import pytest
    import cases
    from qgdsolver_b200 import foamcase as fc
    m = cases.pm.hex_box(3, 2, 2)
    names = [p.name for p in m.patches]
    types = {n: "zeroGradient" for n in names}
    types[names[0]], types[names[1]] = "fixedGradient", "qhdFlux"
    rng = np.random.default_rng(1)
    g0, g1 = rng.random(m.patches[0].size), rng.random(m.patches[1].size)
    bnd = rng.random(m.n_bnd)
    path = str(tmp_path / "0" / "p")
    fc.write_field(path, m, "p", rng.random(m.n_cells), types, bnd, gradients={names[0]: g0, names[1]: g1})
    f = fc.read_field(path, m)
    assert np.array_equal(f.patch_gradients[names[0]], g0) and np.array_equal(f.patch_gradients[names[1]], g1)
    kinds, vals = fc.bc_arrays(m, f)
    nI = m.n_internal
    assert kinds[0] == 2 and kinds[1] == 5
    assert np.array_equal(vals[m.patches[0].start - nI:m.patches[0].start - nI + m.patches[0].size], g0)
    assert np.array_equal(vals[m.patches[1].start - nI:m.patches[1].start - nI + m.patches[1].size], g1)
    with pytest.raises(fc.FoamFormatError, match="needs its gradient"):
        fc.write_field(path, m, "p", np.zeros(m.n_cells), types, bnd)
    f.patch_gradients.pop(names[0])
    with pytest.raises(fc.FoamFormatError, match="without a `gradient` entry"):
        fc.bc_arrays(m, f)
    # uniform lists N{v}
    (tmp_path / "lab").write_text("FoamFile { version 2.0; format ascii; class labelList; object lab; }\n8{3}\n")
    assert np.array_equal(fc.read_labels(str(tmp_path / "lab")), np.full(8, 3))
    assert np.array_equal(fc._parse_value("nonuniform List<scalar> 4{0.5}", 4, 1), np.full(4, 0.5))
    assert np.array_equal(fc._parse_value("nonuniform List<vector> 2{(1 2 3)}", 2, 3), np.tile([1.0, 2.0, 3.0], (2, 1)))
