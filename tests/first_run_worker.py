"""Runs ONE first-run device case in its own process (tests/test_zzz_*_gpu_*.py spawn it), so that a crash or a hang of code that
has never run on a device stays inside that process: `python tests/first_run_worker.py <kind> [<case>]` prints FIRST_RUN_OK."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from qgdsolver_b200 import polymesh as pm  # noqa: E402


def rel_linf(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check(cond, what):
    if not cond:
        raise AssertionError(what)


def _sutherland(c):
    c.sutherland = dict(As=2.5e-3, Ts=0.4)
    return c


VARSC5 = {
    "hex_mixed": (lambda: cases.case_hex3d(n=(10, 9, 8), perturb=0.15, bcs="mixed", model="varScModel5"), 60),
    "hex_fixed_quality_floor": (lambda: cases.case_hex3d(n=(7, 6, 5), perturb=0.2, bcs="fixed", model="varScModel5",
                                                         varsc=dict(rC=0.35, smoothCoeff=0.15, maxAspectRatio=1.2)), 60),
    "2d_mixed_cellset": (lambda: cases.case_2d((24, 20), perturb=0.2, bcs="mixed", model="varScModel5",
                                               varsc=dict(const_sc_cells=np.array([3, 50, 77, 300], np.int32))), 60),
    "truncoct_zg": (lambda: cases.case_truncoct(n=(5, 4, 4), bcs="zg", model="varScModel5", varsc=dict(smoothCoeff=0.05)), 40),
    "prism_qgdflux_offsets": (lambda: cases.case_prism(bcs="qgdflux", model="varScModel5", gas=dict(cases.GAS, Tref=0.2, Hsref=0.1)), 60),
    "sod_adjust": (lambda: cases.case_sod(200, model="varScModel5", adjust_time_step=True, max_co=0.2, varsc=dict(rC=0.7)), 100),
    "hex_sutherland": (lambda: _sutherland(cases.case_hex3d(perturb=0.1, bcs="mixed", model="varScModel5", gas=dict(cases.GAS, mu=3e-3))), 40),
    "hex_reduced_scheme": (lambda: cases.case_hex3d(perturb=0.1, bcs="mixed", model="varScModel5", scheme="reduced"), 40),
    # implicitDiffusion true (the reference's default): the model's pass follows the closing phase of the implicit step
    "prism_implicit": (lambda: cases.case_prism(bcs="fixed", model="varScModel5", implicit=True), 40),
    "2d_implicit_qgdflux_adjust": (lambda: cases.case_2d(perturb=0.1, bcs="qgdflux", model="varScModel5", implicit=True,
                                                         adjust_time_step=True, max_co=0.1,
                                                         varsc=dict(rC=0.9, minSc=0.002, smoothCoeff=0.3)), 40),
}

WEDGE = {
    "wedge_fixed": lambda: cases.case_wedge(n=(16, 12), bcs="fixed"),
    "wedge_perturbed_mixed": lambda: cases.case_wedge(n=(14, 10), perturb=0.15, bcs="mixed", angle_deg=3.0),
    "wedge_qgdflux_adjust": lambda: cases.case_wedge(n=(12, 10), bcs="qgdflux", adjust_time_step=True, max_co=0.1),
    "wedge_leastSquares": lambda: cases.case_wedge(n=(12, 10), bcs="fixed", scheme="leastSquares"),
    "wedge_model1n": lambda: cases.case_wedge(n=(12, 10), perturb=0.1, bcs="mixed", model="constScPrModel1n"),
    # polyPatch type symmetryPlane: slip velocity + constrained vertices (combined where two planes meet)
    "symplane_hex": lambda: cases.with_symmetry_planes(cases.case_hex3d(n=(9, 8, 7), perturb=0.2, bcs="mixed"), ("yMin", "yMax", "zMax")),
    "symplane_forward_step": lambda: cases.with_symmetry_planes(cases.case_forward_step(n=30), ("yMin", "yMax", "step")),
    "symplane_2d_leastSquares": lambda: cases.with_symmetry_planes(cases.case_2d((20, 16), perturb=0.15, bcs="fixed", scheme="leastSquares"), ("yMin", "xMax")),
}


def run_varsc5(qgd, O, name):
    fn, n_steps = VARSC5[name]
    c = fn()
    o = c.make_oracle(O)
    s = c.make_solver(qgd)
    live = c.mesh.patch_kind_per_bface() != 1
    # start-up: QGDCoeffs::correct has run twice (thermo constructor + thermo.correct(), createFields.H:3-8)
    a, ab = s.get("ScQGD", with_bnd=True)
    b, bb = o.get("ScQGD", with_bnd=True)
    check(rel_linf(a, b) < 1e-12 and rel_linf(ab[live], bb[live]) < 1e-12, "ScQGD after start-up")
    check(rel_linf(s.get("mu"), o.get("mu")) < 1e-12, "mu after start-up")
    c.oracle_step(o, n_steps)
    s.step(n_steps)
    for f in ("rho", "rhoU", "rhoE"):
        e = rel_linf(s.get(f), o.get(f))
        print(f"varsc5 {name} steps={n_steps} {f} relLinf={e:.3e}")
        check(e < 1e-10, f)
    a, ab = s.get("ScQGD", with_bnd=True)
    b, bb = o.get("ScQGD", with_bnd=True)
    check(b.max() > 1.1 * b.min(), "the sensor is active: ScQGD is not a constant field")
    e, eb = rel_linf(a, b), rel_linf(ab[live], bb[live])
    print(f"varsc5 {name} ScQGD relLinf={e:.3e} boundary {eb:.3e} range [{b.min():.4f}, {b.max():.4f}]")
    check(e < 1e-9 and eb < 1e-9, "ScQGD")
    for f in ("mu", "alpha", "tauQGD", "T", "p"):
        x, xb = s.get(f, with_bnd=True)
        y, yb = o.get(f, with_bnd=True)
        check(rel_linf(x, y) < 1e-9, f)
        check(rel_linf(xb[live], yb[live]) < 1e-9, f + " (boundary)")
    if c.opts["adjust_time_step"]:
        check(abs(s.scalars()["deltaT"] - o.deltaT()) < 1e-12 * o.deltaT(), "deltaT")


def run_varsc5_refusals(qgd, O):
    c = cases.case_hex3d(perturb=0.1, bcs="zg", model="varScModel5")
    s = c.make_solver(qgd)
    n0 = s.launch_count()
    s.step(2)
    check(s.launch_count() - n0 >= 2 * (6 + 8), "launch count: the step kernels + the model's own pass")

    def refused(fn, code, word):
        try:
            fn()
        except qgd.QGDError as e:
            check(e.code == code and word in e.message, f"wrong refusal: {e.code} {e.message}")
            return
        raise AssertionError("not refused")
    refused(lambda: s.step_fields_host(1, None, None), qgd.ERR_UNSUPPORTED, "ScQGD")
    refused(lambda: s.set_pipeline(1), qgd.ERR_UNSUPPORTED, "varScModel5")


def run_wedge(qgd, O, name):
    c = WEDGE[name]()
    o = c.make_oracle(O)
    s = c.make_solver(qgd)
    live = c.mesh.patch_kind_per_bface() != 1
    _, ab = s.get("U", with_bnd=True)
    _, bb = o.get("U", with_bnd=True)
    check(rel_linf(ab[live], bb[live]) < 1e-14, "U_b = faceT . U_P at start-up")
    c.oracle_step(o, 100)
    s.step(100)
    for f in ("rho", "rhoU", "rhoE"):
        e = rel_linf(s.get(f), o.get(f))
        print(f"wedge {name} steps=100 {f} relLinf={e:.3e}")
        check(e < 1e-10, f)
    for f in ("U", "p", "T"):
        _, xb = s.get(f, with_bnd=True)
        _, yb = o.get(f, with_bnd=True)
        check(rel_linf(xb[live], yb[live]) < 1e-10, f + " (boundary)")
    if c.opts["adjust_time_step"]:
        check(abs(s.scalars()["deltaT"] - o.deltaT()) < 1e-12 * o.deltaT(), "deltaT")


def run_wedge_ops(qgd, O):
    """operator level: fvsc::grad / fvsc::div of scalar, vector and tensor fields on a wedge mesh (vertex constraint on K = 3, 9)"""
    from test_gpu_parity import _fields
    mesh = pm.wedge_box(9, 7, angle_deg=6.0, perturb=0.15, seed=5)
    o = O.Oracle(mesh)
    dm = qgd.Mesh(mesh)
    # GaussVolPoint: no derivative on wedge faces; reduced: nf*snGrad there too; leastSquares: constraint patches stay zero
    for scheme in ("GaussVolPoint", "reduced", "leastSquares"):
        st = qgd.FvscStencil(dm, scheme)
        osch = O.FVSC_SCHEMES[scheme]
        for k in (1, 3):
            cell, bnd, bsg = _fields(mesh, k, 300 + k)
            e = rel_linf(st.Grad(cell, bnd, bsg), o.fvsc_grad(cell, bnd, bsg, scheme=osch))
            print(f"wedge ops {scheme} grad k={k} relLinf={e:.3e}")
            check(e < 1e-12, f"{scheme} grad k={k}")
        for k in (3, 9):
            cell, bnd, bsg = _fields(mesh, k, 400 + k)
            e = rel_linf(st.Div(cell, bnd, bsg), o.fvsc_div(cell, bnd, bsg, scheme=osch))
            print(f"wedge ops {scheme} div k={k} relLinf={e:.3e}")
            check(e < 1e-12, f"{scheme} div k={k}")


def run_wedge_refusals(qgd, O):
    def refused(fn, code, word):
        try:
            fn()
        except qgd.QGDError as e:
            check(e.code == code and word in e.message, f"wrong refusal: {e.code} {e.message}")
            return
        raise AssertionError("not refused")
    refused(lambda: cases.case_wedge(n=(6, 5), implicit=True).make_solver(qgd), qgd.ERR_UNSUPPORTED, "wedge")
    refused(lambda: cases.case_wedge(n=(6, 5), scheme="reduced").make_solver(qgd), qgd.ERR_UNSUPPORTED, "reduced")
    c2 = cases.case_hex3d(bcs="zg")
    c2.bcU[0] = cases.WEDGE                                      # wedge velocity on an ordinary patch
    refused(lambda: c2.make_solver(qgd), qgd.ERR_INVALID, "wedge")


GRAPH = {
    "hex_zg_forked_boundary_stream": lambda: cases.case_hex3d(n=(12, 10, 9), perturb=0.1, bcs="zg"),
    "hex_mixed_qgdflux": lambda: cases.case_hex3d(perturb=0.2, bcs="mixed"),
    "2d_adjust_dt": lambda: cases.case_2d(perturb=0.1, bcs="fixed", adjust_time_step=True, max_co=0.1),
    "prism_model1n": lambda: cases.case_prism(bcs="fixed", model="constScPrModel1n"),
    "hex_varSc7": lambda: cases.case_hex3d(perturb=0.1, bcs="fixed", model="varScModel7", varsc=dict(cSc1=3.0, minSc=0.02, maxSc=0.4)),
    "forward_step_slip": lambda: cases.case_forward_step(n=30),
    "sod_leastSquares": lambda: cases.case_sod(200, scheme="leastSquares"),
}


def run_graph(qgd, O, name):
    """QGD_STEP_GRAPH=1: the captured-and-replayed step must give bit-identical fields to the stream launches"""
    n = 40
    os.environ.pop("QGD_STEP_GRAPH", None)
    c = GRAPH[name]()
    s1 = c.make_solver(qgd)
    l0 = s1.launch_count()
    s1.step(n)
    l1 = s1.launch_count() - l0
    check(s1.graph_steps() == 0, "graph used without the switch")
    os.environ["QGD_STEP_GRAPH"] = "1"
    try:
        s2 = c.make_solver(qgd)
        l0 = s2.launch_count()
        s2.step(n)
        l2 = s2.launch_count() - l0
        first_outside = 1 if c.model == "constScPrModel1n" else 0
        check(s2.graph_steps() == n - first_outside, f"graph steps {s2.graph_steps()}")
        s2.step(1)                                   # a single step takes the stream path again
        s1.step(1)
    finally:
        os.environ.pop("QGD_STEP_GRAPH", None)
    check(l1 == l2, f"kernel launches per {n} steps differ: {l1} vs {l2}")
    for f in ("rho", "rhoU", "rhoE", "p", "T", "mu"):
        a, ab = s1.get(f, with_bnd=True)
        b, bb = s2.get(f, with_bnd=True)
        check(np.array_equal(a, b) and np.array_equal(ab, bb), f + " differs between graph replay and stream launches")
    check(s1.scalars() == s2.scalars(), "time-step scalars differ")
    o = c.make_oracle(O)
    c.oracle_step(o, n + 1)
    for f in ("rho", "rhoU", "rhoE"):
        e = rel_linf(s2.get(f), o.get(f))
        print(f"graph {name} steps={n + 1} {f} relLinf vs oracle={e:.3e}")
        check(e < 1e-10, f)


def main():
    import signal
    signal.alarm(270)                            # never-run device code: a hang ends here, not in the test session
    kind = sys.argv[1]
    name = sys.argv[2] if len(sys.argv) > 2 else None
    import oracle as O
    from qgdsolver_b200 import api
    O.build()
    api.load_library()
    api.init(0)
    {"varsc5": lambda: run_varsc5(api, O, name), "varsc5_refusals": lambda: run_varsc5_refusals(api, O),
     "wedge": lambda: run_wedge(api, O, name), "wedge_ops": lambda: run_wedge_ops(api, O),
     "wedge_refusals": lambda: run_wedge_refusals(api, O), "graph": lambda: run_graph(api, O, name)}[kind]()
    print("FIRST_RUN_OK", kind, name or "")


if __name__ == "__main__":
    main()
