"""OpenFOAM case ingestion and write-back (ASCII): the on-disk half of the drop-in.

Reads what the reference solvers read through OpenFOAM's IO layer
  constant/polyMesh/{points,faces,owner,neighbour,boundary}      (createMesh.H)
  <time>/U, T, p, alphaQGD                                        (QGDFoam/createFields.H:24-35, QGDCoeffs.C:119-160)
and writes volScalarField / volVectorField files back, so a case prepared for QGDFoam / QHDFoam can be stepped by
libqgd_b200 and inspected with the usual tools.  Host-side harness code (numpy); geometry comes from
PolyMesh.compute_geometry() [OF-v2312 semantics].

Supported subset: ASCII and little-endian binary format (label=32|64, scalar=32|64), label/scalar lists in the `N ( ... )` and `N{v}` forms, faces as `n(v0 v1 ...)`,
boundary patches of type patch | wall | empty | processor | wedge | symmetry*, patch fields fixedValue | zeroGradient |
fixedGradient | qgdFlux | qhdFlux | empty | calculated | slip (-> unsupported by the device BC set), `uniform` and
`nonuniform List<...>` values; decomposePar output (processorN meshes + *ProcAddressing, cellDecomposition).  `#include`
directives and big-endian files are rejected with a clear error.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .polymesh import PATCH_EMPTY, PATCH_GENERIC, PATCH_PROCESSOR, PATCH_SYMMETRY_PLANE, PATCH_WEDGE, Patch, PolyMesh

# `symmetry` (non-planar) is read as an ordinary patch: its per-vertex normals are not restated; `symmetryPlane` keeps its vertex constraint
_KINDS = {"patch": PATCH_GENERIC, "wall": PATCH_GENERIC, "symmetry": PATCH_GENERIC, "symmetryPlane": PATCH_SYMMETRY_PLANE,
          "empty": PATCH_EMPTY, "processor": PATCH_PROCESSOR, "wedge": PATCH_WEDGE}
_KIND_WORD = {PATCH_GENERIC: "patch", PATCH_EMPTY: "empty", PATCH_PROCESSOR: "processor", PATCH_WEDGE: "wedge",
              PATCH_SYMMETRY_PLANE: "symmetryPlane"}


class FoamFormatError(ValueError):
    pass


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


_NCMPT = {"label": 1, "scalar": 1, "vector": 3, "symmTensor": 6, "tensor": 9}


def _load(path: str) -> str:
    """File content as ASCII text.  `format binary` files (OpenFOAM writes a list as `N` newline `(` raw bytes `)`
    [OF-v2312 UList::writeList / OSstream::beginRawWrite]; sizes from the header's arch "LSB;label=32|64;scalar=32|64")
    are transcoded to the equivalent ASCII file, doubles through repr() so nothing is lost."""
    data = open(path, "rb").read()
    m = re.search(rb"FoamFile\s*\{(.*?)\}", data, flags=re.S)
    if not m or not re.search(rb"format\s+binary\s*;", m.group(1)):
        return data.decode("ascii", errors="replace")
    hdr = m.group(1).decode("ascii")
    arch = re.search(r'arch\s+"([^"]*)"', hdr)
    lab = re.search(r"label=(\d+)", arch.group(1)) if arch else None
    sca = re.search(r"scalar=(\d+)", arch.group(1)) if arch else None
    if arch and "MSB" in arch.group(1):
        raise FoamFormatError(f"{path}: big-endian binary files are not supported")
    ldt = np.dtype("<i8") if lab and lab.group(1) == "64" else np.dtype("<i4")
    sdt = np.dtype("<f4") if sca and sca.group(1) == "32" else np.dtype("<f8")
    cls = re.search(r"class\s+(\w+)\s*;", hdr).group(1)
    out = [data[:m.end()].decode("ascii").replace("binary", "ascii")]
    pos = m.end()

    def raw_list(at: int, dt: np.dtype, ncmpt: int):
        """`N (raw)` starting at `at` (leading whitespace / comment lines allowed) -> (array (N,ncmpt), end position)"""
        mm = re.compile(rb"(?:\s|//[^\n]*\n)*(\d+)\s*\(").match(data, at)
        if not mm:
            raise FoamFormatError(f"{path}: expected a binary list at byte {at}")
        n = int(mm.group(1))
        nb = n * ncmpt * dt.itemsize
        if mm.end() + nb + 1 > len(data):
            raise FoamFormatError(f"{path}: binary list of {n} entries is truncated")
        a = np.frombuffer(data, dt, n * ncmpt, mm.end()).reshape(n, ncmpt)
        if data[mm.end() + nb:mm.end() + nb + 1] != b")":
            raise FoamFormatError(f"{path}: binary list of {n} entries is not closed by `)`")
        return a, mm.end() + nb + 1

    def fmt_rows(a: np.ndarray, integer: bool) -> str:
        if a.shape[1] == 1:
            return "\n".join(str(int(v)) if integer else repr(float(v)) for v in a[:, 0])
        return "\n".join("(" + " ".join(repr(float(x)) for x in v) + ")" for v in a)

    if cls in ("vectorField", "labelList", "scalarField"):
        a, pos = raw_list(pos, ldt if cls == "labelList" else sdt, 3 if cls == "vectorField" else 1)
        out.append(f"\n{a.shape[0]}\n(\n{fmt_rows(a, cls == 'labelList')}\n)\n")
    elif cls == "faceCompactList":                      # CompactListList: offsets (nFaces+1) then the flat vertex labels
        offs, pos = raw_list(pos, ldt, 1)
        flat, pos = raw_list(pos, ldt, 1)
        offs, flat = offs[:, 0], flat[:, 0]
        out[0] = out[0].replace("faceCompactList", "faceList")
        out.append(f"\n{offs.size - 1}\n(\n" + "\n".join(
            f"{int(offs[i + 1] - offs[i])}(" + " ".join(str(int(v)) for v in flat[offs[i]:offs[i + 1]]) + ")" for i in range(offs.size - 1)) + "\n)\n")
    else:                                               # dictionaries / fields: binary only inside `List<type> N(raw)`
        pat = re.compile(rb"List<(\w+)>\s*(?=\d)")
        while True:
            mm = pat.search(data, pos)
            if not mm:
                break
            typ = mm.group(1).decode()
            if typ not in _NCMPT:
                raise FoamFormatError(f"{path}: binary List<{typ}> is not supported")
            a, end = raw_list(mm.end(), ldt if typ == "label" else sdt, _NCMPT[typ])
            out.append(data[pos:mm.end()].decode("ascii", errors="replace"))
            out.append(f"{a.shape[0]}\n(\n{fmt_rows(a, typ == 'label')}\n)")
            pos = end
    out.append(data[pos:].decode("ascii", errors="replace"))
    return "".join(out)


def _split_header(text: str) -> Tuple[Dict[str, str], str]:
    """FoamFile header dictionary and the remaining body."""
    text = _strip_comments(text)
    m = re.search(r"FoamFile\s*\{(.*?)\}", text, flags=re.S)
    hdr: Dict[str, str] = {}
    body = text
    if m:
        for k, v in re.findall(r"(\w+)\s+([^;]+);", m.group(1)):
            hdr[k] = v.strip().strip('"')
        body = text[m.end():]
    if hdr.get("format", "ascii") != "ascii":
        raise FoamFormatError("binary content must go through _load() (internal error)")
    if "#include" in body:
        raise FoamFormatError("#include directives are not supported")
    return hdr, body


def _read_list_body(body: str) -> Tuple[int, str]:
    """`N ( ... )` -> (N, inner text); the uniform form `N{v}` is expanded to N copies of v"""
    u = re.match(r"\s*(\d+)\s*\{([^}]*)\}", body)
    if u:
        n = int(u.group(1))
        return n, " ".join([u.group(2).strip()] * n)
    m = re.search(r"(\d+)\s*\(", body)
    if not m:
        raise FoamFormatError("expected a list `N ( ... )`")
    n = int(m.group(1))
    start = m.end()
    end = body.rindex(")")
    return n, body[start:end]


def read_points(path: str) -> np.ndarray:
    _, body = _split_header(_load(path))
    n, inner = _read_list_body(body)
    vals = np.array(re.sub(r"[()]", " ", inner).split(), dtype=np.float64)
    if vals.size != 3 * n:
        raise FoamFormatError(f"{path}: expected {n} points")
    return np.ascontiguousarray(vals.reshape(n, 3))


def read_labels(path: str) -> np.ndarray:
    _, body = _split_header(_load(path))
    n, inner = _read_list_body(body)
    vals = np.array(inner.split(), dtype=np.int64)
    if vals.size != n:
        raise FoamFormatError(f"{path}: expected {n} labels")
    return vals.astype(np.int32)


def read_faces(path: str) -> Tuple[np.ndarray, np.ndarray]:
    _, body = _split_header(_load(path))
    n, inner = _read_list_body(body)
    offs = [0]
    verts: List[int] = []
    for cnt, vs in re.findall(r"(\d+)\s*\(([^()]*)\)", inner):
        v = vs.split()
        if len(v) != int(cnt):
            raise FoamFormatError(f"{path}: face with {cnt} vertices lists {len(v)}")
        verts.extend(int(x) for x in v)
        offs.append(len(verts))
    if len(offs) - 1 != n:
        raise FoamFormatError(f"{path}: expected {n} faces, found {len(offs) - 1}")
    return np.asarray(offs, np.int32), np.asarray(verts, np.int32)


def read_boundary(path: str) -> List[Patch]:
    _, body = _split_header(_load(path))
    n, inner = _read_list_body(body)
    patches = []
    for name, blk in re.findall(r"(\w+)\s*\{([^{}]*)\}", inner):
        d = {k: v.strip() for k, v in re.findall(r"(\w+)\s+([^;]+);", blk)}
        typ = d.get("type", "patch")
        if typ not in _KINDS:
            raise FoamFormatError(f"{path}: patch {name} has unsupported type {typ}")
        patches.append(Patch(name, _KINDS[typ], int(d["startFace"]), int(d["nFaces"]), int(d.get("neighbProcNo", -1))))
    if len(patches) != n:
        raise FoamFormatError(f"{path}: expected {n} patches, found {len(patches)}")
    return patches


def read_polymesh(case_dir: str, region: str = "") -> PolyMesh:
    """constant/polyMesh -> PolyMesh with fvMesh geometry."""
    d = os.path.join(case_dir, "constant", region, "polyMesh")
    pts = read_points(os.path.join(d, "points"))
    offs, verts = read_faces(os.path.join(d, "faces"))
    owner = read_labels(os.path.join(d, "owner"))
    neighbour = read_labels(os.path.join(d, "neighbour"))
    patches = read_boundary(os.path.join(d, "boundary"))
    n_cells = int(max(owner.max(), neighbour.max() if neighbour.size else -1)) + 1
    if owner.size != offs.size - 1:
        raise FoamFormatError("owner and faces disagree on the number of faces")
    if neighbour.size and not (owner[:neighbour.size] < neighbour).all():
        raise FoamFormatError("polyMesh is not in upper-triangular order (owner < neighbour)")
    mesh = PolyMesh(points=pts, face_offsets=offs, face_verts=verts, owner=owner, neighbour=neighbour, patches=patches,
                    n_cells=n_cells)
    mesh.compute_geometry()
    # geometricD: a direction is solved unless an `empty` or `wedge` patch removes it  [OF polyMesh::calcDirections]
    gd = np.ones(3, np.int32)
    nI = mesh.n_internal
    for p in patches:
        if p.kind in (PATCH_EMPTY, PATCH_WEDGE) and p.size:
            nf = mesh.Sf[p.start:p.start + p.size] / mesh.magSf[p.start:p.start + p.size, None]
            gd[int(np.argmax(np.abs(nf).mean(0)))] = -1
    mesh.geometric_d = gd
    del nI
    return mesh


# ------------------------------------------------------------------------------------------------ writers
_HDR = """/*--------------------------------*- C++ -*----------------------------------*\\
| qgd-b200 case writer (OpenFOAM ASCII format)                                |
\\*---------------------------------------------------------------------------*/
FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    location    "{loc}";
    object      {obj};
}}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


def _fmt(v: float) -> str:
    return repr(float(v))


_ARCH = '    arch        "LSB;label=32;scalar=64";\n'


def _hdr(cls: str, loc: str, obj: str, binary: bool) -> bytes:
    h = _HDR.format(cls=cls, loc=loc, obj=obj)
    if binary:
        h = h.replace("format      ascii;\n", "format      binary;\n" + _ARCH)
    return h.encode("ascii")


def _raw(a: np.ndarray, dt: str) -> bytes:
    """OpenFOAM binary list: N newline ( raw bytes )"""
    a = np.ascontiguousarray(a, dtype=dt)
    return f"\n{a.shape[0]}\n(".encode("ascii") + a.tobytes() + b")\n"


def write_polymesh(mesh: PolyMesh, case_dir: str, binary: bool = False) -> None:
    """constant/polyMesh under `case_dir` (a case or a processorN directory).  binary: `format binary` files
    (points as raw doubles, faces as a faceCompactList, owner / neighbour as raw 32-bit labels); the boundary file is
    always a plain dictionary."""
    d = os.path.join(case_dir, "constant", "polyMesh")
    os.makedirs(d, exist_ok=True)
    if binary:
        with open(os.path.join(d, "points"), "wb") as f:
            f.write(_hdr("vectorField", "constant/polyMesh", "points", True) + _raw(mesh.points, "<f8"))
        with open(os.path.join(d, "faces"), "wb") as f:
            f.write(_hdr("faceCompactList", "constant/polyMesh", "faces", True) + _raw(mesh.face_offsets, "<i4") + _raw(mesh.face_verts, "<i4"))
        for name, arr in (("owner", mesh.owner), ("neighbour", mesh.neighbour)):
            with open(os.path.join(d, name), "wb") as f:
                f.write(_hdr("labelList", "constant/polyMesh", name, True) + _raw(arr, "<i4"))
    else:
        with open(os.path.join(d, "points"), "w") as f:
            f.write(_HDR.format(cls="vectorField", loc="constant/polyMesh", obj="points"))
            f.write(f"{mesh.n_points}\n(\n" + "\n".join(f"({_fmt(p[0])} {_fmt(p[1])} {_fmt(p[2])})" for p in mesh.points) + "\n)\n")
        with open(os.path.join(d, "faces"), "w") as f:
            f.write(_HDR.format(cls="faceList", loc="constant/polyMesh", obj="faces"))
            o = mesh.face_offsets
            f.write(f"{mesh.n_faces}\n(\n" + "\n".join(
                f"{o[i + 1] - o[i]}(" + " ".join(str(int(v)) for v in mesh.face_verts[o[i]:o[i + 1]]) + ")" for i in range(mesh.n_faces)) + "\n)\n")
        for name, arr in (("owner", mesh.owner), ("neighbour", mesh.neighbour)):
            with open(os.path.join(d, name), "w") as f:
                f.write(_HDR.format(cls="labelList", loc="constant/polyMesh", obj=name))
                f.write(f"{arr.size}\n(\n" + "\n".join(str(int(v)) for v in arr) + "\n)\n")
    with open(os.path.join(d, "boundary"), "w") as f:
        f.write(_HDR.format(cls="polyBoundaryMesh", loc="constant/polyMesh", obj="boundary"))
        f.write(f"{len(mesh.patches)}\n(\n")
        for p in mesh.patches:
            f.write(f"    {p.name}\n    {{\n        type            {_KIND_WORD[p.kind]};\n")
            if p.kind == PATCH_PROCESSOR:       # processorPolyPatch::write [OF-v2312]
                me = re.match(r"procBoundary(\d+)to\d+", p.name)
                f.write("        inGroups        1(processor);\n")
            f.write(f"        nFaces          {p.size};\n        startFace       {p.start};\n")
            if p.kind == PATCH_PROCESSOR:
                f.write("        matchTolerance  0.0001;\n        transform       unknown;\n"
                        f"        myProcNo        {int(me.group(1)) if me else 0};\n        neighbProcNo    {p.neighb_rank};\n")
            f.write("    }\n")
        f.write(")\n")


def write_labels(path: str, arr: np.ndarray, loc: str = "constant/polyMesh") -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(_HDR.format(cls="labelList", loc=loc, obj=os.path.basename(path)))
        f.write(f"{arr.size}\n(\n" + "\n".join(str(int(v)) for v in arr) + "\n)\n")


# ------------------------------------------------------------------------------------------------ decomposePar output
_ADDR = (("cellProcAddressing", "cell_addr"), ("faceProcAddressing", "face_addr"), ("pointProcAddressing", "point_addr"),
         ("boundaryProcAddressing", "boundary_addr"))


def write_decomposed_case(mesh: PolyMesh, cell_rank: np.ndarray, case_dir: str, write_cell_dist: bool = True):
    """processor0..N-1/constant/polyMesh + the four *ProcAddressing lists, laid out as `decomposePar` writes them
    (decompose.processor_meshes), and constant/cellDecomposition as `decomposePar -cellDist` does.  Returns the ProcMesh list."""
    from . import decompose
    procs = decompose.processor_meshes(mesh, cell_rank)
    for p in procs:
        d = os.path.join(case_dir, f"processor{p.rank}")
        write_polymesh(p.mesh, d)
        for fname, attr in _ADDR:
            write_labels(os.path.join(d, "constant", "polyMesh", fname), getattr(p, attr))
    if write_cell_dist:
        write_labels(os.path.join(case_dir, "constant", "cellDecomposition"), np.asarray(cell_rank), loc="constant")
    return procs


def read_decomposed_case(case_dir: str):
    """All processorN directories of a decomposed case -> list of decompose.ProcMesh (geometry computed per processor)."""
    from . import decompose
    procs = []
    r = 0
    while os.path.isdir(os.path.join(case_dir, f"processor{r}")):
        d = os.path.join(case_dir, f"processor{r}")
        pm = read_polymesh(d)
        a = {attr: read_labels(os.path.join(d, "constant", "polyMesh", fname)) for fname, attr in _ADDR}
        procs.append(decompose.ProcMesh(r, pm, a["cell_addr"], a["face_addr"], a["point_addr"], a["boundary_addr"]))
        r += 1
    if not procs:
        raise FoamFormatError(f"{case_dir}: no processor0 directory (run decomposePar first)")
    return procs


def read_cell_decomposition(case_dir: str, n_cells: int) -> np.ndarray:
    """cell -> processor map of a decomposed case: constant/cellDecomposition when present (decomposePar -cellDist),
    else rebuilt from the cellProcAddressing lists.  This is the map the multi-GPU path shards by, so the device
    decomposition is the decomposePar (scotch) decomposition bit for bit."""
    from . import decompose
    path = os.path.join(case_dir, "constant", "cellDecomposition")
    if os.path.exists(path):
        rank = read_labels(path)
        if rank.size != n_cells:
            raise FoamFormatError(f"{path}: {rank.size} entries for a mesh of {n_cells} cells")
        return rank.astype(np.int32)
    procs = []
    r = 0
    while os.path.isdir(os.path.join(case_dir, f"processor{r}")):
        ca = read_labels(os.path.join(case_dir, f"processor{r}", "constant", "polyMesh", "cellProcAddressing"))
        procs.append(decompose.ProcMesh(r, None, ca, None, None, None))
        r += 1
    if not procs:
        raise FoamFormatError(f"{case_dir}: neither constant/cellDecomposition nor processor directories")
    return decompose.cell_rank_from_procs(procs, n_cells)


def proc_patch_gradients(mesh: PolyMesh, gradients: Optional[Dict[str, np.ndarray]], proc) -> Optional[Dict[str, np.ndarray]]:
    """`gradient` entries of a global field (patch name -> per-face array in global patch order) cut down to the faces a
    processor mesh keeps of each patch (through faceProcAddressing)"""
    if not gradients:
        return None
    start = {p.name: p.start for p in mesh.patches}
    out = {}
    for patch in proc.mesh.patches:
        if patch.name in gradients and patch.name in start:
            gf = np.abs(proc.face_addr[patch.start:patch.start + patch.size].astype(np.int64)) - 1
            out[patch.name] = np.asarray(gradients[patch.name])[gf - start[patch.name]]
    return out


def write_processor_fields(case_dir: str, time: str, name: str, procs, internal: np.ndarray, patch_types: Dict[str, str],
                           boundary: Optional[np.ndarray] = None, n_internal_global: int = 0,
                           dimensions: str = "[0 0 0 0 0 0 0]", gradients: Optional[Dict[str, np.ndarray]] = None,
                           mesh: Optional[PolyMesh] = None) -> None:
    """Scatter a global cell field (and optionally its boundary values, nBnd[,k] in global boundary-face order) into
    processorN/<time>/<name>, as `decomposePar -fields` / a parallel run would leave it; processor patches get type
    `processor` with the value of the cell across the patch left to the reader (calculated from the owner here)."""
    internal = np.asarray(internal)
    for p in procs:
        loc = internal[p.cell_addr]
        pt = dict(patch_types)
        bl = None
        if boundary is not None:
            nIl = p.mesh.n_internal
            gf = np.abs(p.face_addr[nIl:].astype(np.int64)) - 1
            phys = gf >= n_internal_global
            b = np.asarray(boundary)
            bl = np.where(phys.reshape((-1,) + (1,) * (b.ndim - 1)), b[np.where(phys, gf - n_internal_global, 0)],
                          loc[p.mesh.owner[nIl:]])
        for patch in p.mesh.patches:
            if patch.kind == PATCH_PROCESSOR:
                pt[patch.name] = "processor"
        write_field(os.path.join(case_dir, f"processor{p.rank}", time, name), p.mesh, name, loc, pt, bl, dimensions,
                    gradients=proc_patch_gradients(mesh, gradients, p) if mesh is not None else None)


def reconstruct_fields(case_dir: str, time: str, names, mesh: PolyMesh, binary: bool = False) -> Dict[str, np.ndarray]:
    """reconstructPar for vol fields: processorN/<time>/<name> of every processor -> <time>/<name> of the undecomposed
    case through cellProcAddressing / faceProcAddressing; patch types are taken from processor0 (processor patches dropped).
    Returns the reconstructed internal fields."""
    procs = read_decomposed_case(case_dir)
    nI = mesh.n_internal
    out = {}
    for name in names:
        internal, boundary, types, have_values, grads = None, None, {}, False, {}
        for p in procs:
            f = read_field(os.path.join(case_dir, f"processor{p.rank}", time, name), p.mesh)
            if internal is None:
                shape = (mesh.n_cells,) + f.internal.shape[1:]
                internal = np.zeros(shape)
                boundary = np.zeros((mesh.n_bnd,) + f.internal.shape[1:])
            internal[p.cell_addr] = f.internal
            for patch in p.mesh.patches:
                if patch.kind == PATCH_PROCESSOR:
                    continue
                types.setdefault(patch.name, f.patch_types[patch.name])
                gf = np.abs(p.face_addr[patch.start:patch.start + patch.size].astype(np.int64)) - 1
                if patch.size and patch.name in f.patch_values:
                    have_values = True
                    boundary[gf - nI] = f.patch_values[patch.name]
                if patch.size and patch.name in f.patch_gradients:
                    gp = next(q for q in mesh.patches if q.name == patch.name)
                    g = grads.setdefault(patch.name, np.zeros((gp.size,) + f.internal.shape[1:]))
                    g[gf - gp.start] = f.patch_gradients[patch.name]
        write_field(os.path.join(case_dir, time, name), mesh, name, internal, types, boundary if have_values else None, binary=binary,
                    gradients=grads or None)
        out[name] = internal
    return out


@dataclass
class VolField:
    """GeometricField<Type, fvPatchField, volMesh> as read from a time directory."""
    name: str
    ncmpt: int
    internal: np.ndarray                                   # (nCells,) or (nCells,3)
    patch_types: Dict[str, str] = field(default_factory=dict)
    patch_values: Dict[str, np.ndarray] = field(default_factory=dict)      # value entry, expanded to the patch size
    patch_gradients: Dict[str, np.ndarray] = field(default_factory=dict)   # gradient entry
    dimensions: str = "[0 0 0 0 0 0 0]"


def _parse_value(txt: str, n: int, ncmpt: int) -> np.ndarray:
    txt = txt.strip()
    shape = (n, ncmpt) if ncmpt > 1 else (n,)
    if txt.startswith("uniform"):
        vals = np.array(re.sub(r"[()]", " ", txt[len("uniform"):]).split(), dtype=np.float64)
        if vals.size != ncmpt:
            raise FoamFormatError(f"uniform value with {vals.size} components, expected {ncmpt}")
        return np.ascontiguousarray(np.broadcast_to(vals if ncmpt > 1 else vals[0], shape)).copy()
    u = re.match(r"nonuniform\s+List<\w+>\s*(\d+)\s*\{([^}]*)\}\s*$", txt, flags=re.S)
    if u:                                                       # `N{v}`: N identical entries
        vals = np.array(re.sub(r"[()]", " ", u.group(2)).split(), dtype=np.float64)
        if int(u.group(1)) != n or vals.size != ncmpt:
            raise FoamFormatError(f"uniform list {u.group(1)}{{...}} with {vals.size} components, expected {n} entries of {ncmpt}")
        return np.ascontiguousarray(np.broadcast_to(vals if ncmpt > 1 else vals[0], shape)).copy()
    m = re.match(r"nonuniform\s+List<\w+>\s*(\d+)\s*\((.*)\)\s*$", txt, flags=re.S)
    if not m:
        raise FoamFormatError(f"cannot parse field value: {txt[:60]}...")
    cnt = int(m.group(1))
    vals = np.array(re.sub(r"[()]", " ", m.group(2)).split(), dtype=np.float64)
    if cnt != n or vals.size != n * ncmpt:
        raise FoamFormatError(f"nonuniform list of {cnt} entries, expected {n}")
    return np.ascontiguousarray(vals.reshape(shape))


def _entries(block: str) -> Dict[str, str]:
    """`key value;` entries of a dictionary block (values may span lines and contain parentheses)."""
    out, i = {}, 0
    for m in re.finditer(r"(\w+)\s+((?:[^;{}]|\n)*?);", block):
        out[m.group(1)] = m.group(2).strip()
    del i
    return out


def read_field(path: str, mesh: PolyMesh) -> VolField:
    hdr, body = _split_header(_load(path))
    cls = hdr.get("class", "volScalarField")
    ncmpt = {"volScalarField": 1, "volVectorField": 3}.get(cls)
    if ncmpt is None:
        raise FoamFormatError(f"{path}: unsupported field class {cls}")
    mb = re.search(r"boundaryField\s*\{", body)
    if not mb:
        raise FoamFormatError(f"{path}: no boundaryField")
    top = _entries(body[:mb.start()])
    fld = VolField(os.path.basename(path), ncmpt, _parse_value(top["internalField"], mesh.n_cells, ncmpt),
                   dimensions=top.get("dimensions", "[0 0 0 0 0 0 0]"))
    depth, j = 1, mb.end()
    while depth and j < len(body):
        depth += {"{": 1, "}": -1}.get(body[j], 0)
        j += 1
    inner = body[mb.end():j - 1]
    sizes = {p.name: p.size for p in mesh.patches}
    for name, blk in re.findall(r"(\w+)\s*\{([^{}]*)\}", inner):
        if name not in sizes:
            raise FoamFormatError(f"{path}: boundaryField entry {name} is not a patch of the mesh")
        e = _entries(blk)
        fld.patch_types[name] = e.get("type", "calculated")
        for key in ("value", "gradient"):       # `value $internalField;` (the usual tutorial idiom): only a uniform field can be expanded
            if e.get(key, "").startswith("$internalField"):
                if not top["internalField"].strip().startswith("uniform"):
                    raise FoamFormatError(f"{path}: `{key} $internalField` on patch {name} with a nonuniform internalField")
                e[key] = top["internalField"]
        if "value" in e:
            fld.patch_values[name] = _parse_value(e["value"], sizes[name], ncmpt)
        if "gradient" in e:
            fld.patch_gradients[name] = _parse_value(e["gradient"], sizes[name], ncmpt)
    missing = [p.name for p in mesh.patches if p.name not in fld.patch_types]
    if missing:
        raise FoamFormatError(f"{path}: no boundaryField entry for patches {missing}")
    return fld


def write_field(path: str, mesh: PolyMesh, name: str, internal: np.ndarray, patch_types: Dict[str, str],
                boundary: Optional[np.ndarray] = None, dimensions: str = "[0 0 0 0 0 0 0]", binary: bool = False,
                gradients: Optional[Dict[str, np.ndarray]] = None) -> None:
    """Write a volScalarField / volVectorField; `boundary` (nBnd[,3]) supplies `value` entries, `gradients` (patch name ->
    per-face gradient) the mandatory `gradient` entry of fixedGradient / qgdFlux / qhdFlux patches (OpenFOAM cannot read such a
    patch without it, and a restart would silently continue with gradient 0).  binary: nonuniform lists are written as raw
    little-endian doubles (`format binary`)."""
    internal = np.asarray(internal, np.float64)
    ncmpt = 1 if internal.ndim == 1 else internal.shape[1]
    cls, typ = ("volScalarField", "scalar") if ncmpt == 1 else ("volVectorField", "vector")

    def lst(a):
        if binary:
            return f"nonuniform List<{typ}> ".encode("ascii") + _raw(np.asarray(a, np.float64), "<f8").strip(b"\n")
        return lst_ascii(a).encode("ascii")

    def lst_ascii(a):
        if ncmpt == 1:
            return f"nonuniform List<{typ}> {a.shape[0]}\n(\n" + "\n".join(_fmt(v) for v in a) + "\n)"
        return f"nonuniform List<{typ}> {a.shape[0]}\n(\n" + "\n".join("(" + " ".join(_fmt(x) for x in v) + ")" for v in a) + "\n)"

    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        w = lambda t: f.write(t.encode("ascii") if isinstance(t, str) else t)
        w(_hdr(cls, os.path.basename(os.path.dirname(path)), name, binary))
        w(f"dimensions      {dimensions};\n\ninternalField   ")
        w(lst(internal))
        w(";\n\nboundaryField\n{\n")
        nI = mesh.n_internal
        for p in mesh.patches:
            t = patch_types.get(p.name, "empty" if p.kind == PATCH_EMPTY else "calculated")
            w(f"    {p.name}\n    {{\n        type            {t};\n")
            if t in ("fixedGradient", "qgdFlux", "qhdFlux") and p.kind != PATCH_EMPTY:
                g = None if gradients is None else gradients.get(p.name)
                if g is None:
                    if t == "fixedGradient":
                        raise FoamFormatError(f"write_field {name}: fixedGradient patch {p.name} needs its gradient (gradients=...)")
                    g = np.zeros((p.size, ncmpt) if ncmpt > 1 else p.size)     # qgdFlux / qhdFlux re-evaluate it in updateCoeffs()
                g = np.broadcast_to(np.asarray(g, np.float64), (p.size, ncmpt) if ncmpt > 1 else (p.size,))
                w("        gradient        ")
                w(lst(g) if p.size else f"nonuniform List<{typ}> 0()")
                w(";\n")
            if boundary is not None and p.kind != PATCH_EMPTY:
                if p.size:
                    w("        value           ")
                    w(lst(np.asarray(boundary)[p.start - nI:p.start - nI + p.size]))
                    w(";\n")
                else:           # zero-sized patch of a processor mesh
                    w(f"        value           nonuniform List<{typ}> 0();\n")
            w("    }\n")
        w("}\n")


# ------------------------------------------------------------------------------------------------ solver glue
_BC_CODE = {"fixedValue": 0, "zeroGradient": 1, "fixedGradient": 2, "qgdFlux": 3, "calculated": 4, "qhdFlux": 5, "empty": 1,
            "slip": 6, "symmetryPlane": 6, "symmetry": 6, "wedge": 7}


def bc_arrays(mesh: PolyMesh, fld: VolField) -> Tuple[np.ndarray, np.ndarray]:
    """(kinds per patch, values per boundary face) in the form qgd_qgdfoam_set_bcs / qgd_qhdfoam_set_bcs take:
    fixedValue faces carry the value, fixedGradient / qhdFlux faces the gradient."""
    kinds = np.zeros(len(mesh.patches), np.int32)
    nB, nI = mesh.n_bnd, mesh.n_internal
    vals = np.zeros((nB, fld.ncmpt) if fld.ncmpt > 1 else nB)
    for i, p in enumerate(mesh.patches):
        t = fld.patch_types[p.name]
        if t not in _BC_CODE:
            raise FoamFormatError(f"patch field type {t} of {fld.name} on {p.name} is outside the device-native set")
        kinds[i] = _BC_CODE[t]
        if kinds[i] in (6, 7) and fld.ncmpt == 1:   # basicSymmetry / wedge on a scalar: the patch-internal value [OF-v2312]
            kinds[i] = 1
        sl = slice(p.start - nI, p.start - nI + p.size)
        if t == "fixedValue":
            vals[sl] = fld.patch_values[p.name]
        elif t in ("fixedGradient", "qhdFlux"):
            if p.name in fld.patch_gradients:
                vals[sl] = fld.patch_gradients[p.name]
            elif t == "fixedGradient" and p.size:               # fixedGradientFvPatchField reads `gradient` with a mandatory lookup
                raise FoamFormatError(f"patch {p.name} of {fld.name}: fixedGradient without a `gradient` entry")
    return kinds, vals


def read_case_fields(case_dir: str, mesh: PolyMesh, time: str = "0", names=("U", "T", "p")) -> Dict[str, VolField]:
    out = {n: read_field(os.path.join(case_dir, time, n), mesh) for n in names}
    a = os.path.join(case_dir, time, "alphaQGD")
    if os.path.exists(a):                                   # QGDCoeffs.C:119-143
        out["alphaQGD"] = read_field(a, mesh)
    return out
