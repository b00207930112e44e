"""decomposePar / reconstructPar stand-ins for cases without an OpenFOAM installation at hand (host-side harness):

    python -m qgdsolver_b200.decompose_case <caseDir> -n N [-time 0]        processorN/ meshes, addressing lists, fields, cellDecomposition
    python -m qgdsolver_b200.decompose_case <caseDir> -reconstruct <time>   processorN/<time>/ fields -> <time>/

The split is the deterministic geometric one of decompose.geometric_split (the `simple` method's slabs); a case decomposed by the
real decomposePar (scotch) is consumed as it is - see foamcase.read_cell_decomposition / runcase -parallel.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import decompose, foamcase


def decompose_case(case_dir: str, n: int, time: str = "0") -> int:
    mesh = foamcase.read_polymesh(case_dir)
    rank = decompose.geometric_split(mesh, n)
    procs = foamcase.write_decomposed_case(mesh, rank, case_dir)
    tdir = os.path.join(case_dir, time)
    nI = mesh.n_internal
    for name in sorted(os.listdir(tdir)) if os.path.isdir(tdir) else []:
        path = os.path.join(tdir, name)
        if not os.path.isfile(path):
            continue
        try:
            f = foamcase.read_field(path, mesh)
        except (foamcase.FoamFormatError, KeyError):
            continue                                  # not a vol field this reader knows
        for p in procs:
            types = dict(f.patch_types)
            pm = p.mesh
            for patch in pm.patches:
                if patch.kind == foamcase.PATCH_PROCESSOR:
                    types[patch.name] = "processor"
            # boundary `value` entries: only where the original patch carried one
            gf = np.abs(p.face_addr[pm.n_internal:].astype(np.int64)) - 1
            bnd = np.zeros((pm.n_bnd,) + f.internal.shape[1:])
            have = False
            for gp in mesh.patches:
                if gp.name in f.patch_values:
                    have = True
                    sel = (gf >= gp.start) & (gf < gp.start + gp.size)
                    bnd[sel] = f.patch_values[gp.name][gf[sel] - gp.start]
            proc_faces = gf < nI
            bnd[proc_faces] = f.internal[p.cell_addr][pm.owner[pm.n_internal:][proc_faces]]
            foamcase.write_field(os.path.join(case_dir, f"processor{p.rank}", time, name), pm, name, f.internal[p.cell_addr], types,
                                 bnd if have else None, f.dimensions, gradients=foamcase.proc_patch_gradients(mesh, f.patch_gradients, p))
    return len(procs)


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 3 or argv[0].startswith("-"):
        print(__doc__)
        return 1
    case_dir = argv[0]
    if argv[1] == "-n":
        time = argv[argv.index("-time") + 1] if "-time" in argv else "0"
        n = decompose_case(case_dir, int(argv[2]), time)
        print(f"decomposed {case_dir} into {n} processor directories")
        return 0
    if argv[1] == "-reconstruct":
        time = argv[2]
        mesh = foamcase.read_polymesh(case_dir)
        p0 = os.path.join(case_dir, "processor0", time)
        names = [f for f in sorted(os.listdir(p0)) if os.path.isfile(os.path.join(p0, f))]
        foamcase.reconstruct_fields(case_dir, time, names, mesh)
        print(f"reconstructed {', '.join(names)} at time {time}")
        return 0
    print(__doc__)
    return 1


if __name__ == "__main__":
    raise SystemExit(main())
