"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/qgd_b200.h declares,
and fails loudly (no CPU fallback) when no CUDA device is usable.  No compute calls are made here."""
import ctypes
import os
import re

import pytest

from qgdsolver_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "qgd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(qgd_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/qgd_b200.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == declared, "api.ABI_SYMBOLS out of sync with the header"


def test_version_and_error_string():
    lib = api.load_library()
    assert lib.qgd_version() >= 100
    assert isinstance(lib.qgd_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.QGDError) as e:
        api.init(0)
    assert e.value.code == api.ERR_CUDA and "no CPU fallback" in e.value.message
    # every compute entry point refuses to run before a successful qgd_init
    from qgdsolver_b200 import polymesh
    with pytest.raises(api.QGDError) as e:
        api.Mesh(polymesh.hex_box(2, 2, 2))
    assert e.value.code == api.ERR_STATE


def test_product_does_not_touch_the_oracle():
    """The product package must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "qgdsolver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "qgd_oracle" not in txt, f
    out = os.popen(f"ldd {api.LIB_PATH}").read()
    assert "oracle" not in out


def _struct_fields(hdr, end_marker):
    end = hdr.index(end_marker)
    body = hdr[hdr.rindex("typedef struct {", 0, end):end]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        first, *rest = decl.split(",")
        names.append(re.sub(r"\[\d+\]", "", first.split()[-1].lstrip("*")))
        names += [re.sub(r"\[\d+\]", "", r.strip().lstrip("*")) for r in rest]
    return names


def test_desc_struct_layout_matches_header():
    # field order / count of the ctypes mirrors follows the C declarations
    hdr = open(os.path.join(ROOT, "include", "qgd_b200.h")).read()
    assert _struct_fields(hdr, "} qgd_qgdfoam_desc;") == [f[0] for f in api.QGDFoamDesc._fields_]
    assert _struct_fields(hdr, "} qgd_qhdfoam_desc;") == [f[0] for f in api.QHDFoamDesc._fields_]
    assert _struct_fields(hdr, "} qgd_mesh_desc;") == [f[0] for f in api._MeshDesc._fields_]


def test_header_is_plain_c():
    """the drop-in boundary is a C ABI: include/qgd_b200.h compiles as C99 (no C++ types in any signature)"""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", os.path.join(ROOT, "include", "qgd_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
