"""Decomposed QGDFoam runs with N = 2, 4 and 8 ranks on ONE GPU (one process per rank, all on cuda:0) against the serial CPU oracle:
the library runs its N-GPU code path unchanged, the NCCL transport underneath is the loopback stand-in tests/fake_nccl (checked on
CPU by tests/test_fake_nccl_cpu.py).  This is how a single-GPU box can exercise what needs eight GPUs otherwise - in particular
the case that hung at N = 8 in round 2 (a corner rank without a qgdFlux face skipped the mid-step exchange): a rank that skips an
exchange blocks its neighbours on a pipe here exactly as it blocked them in ncclRecv, and the test times out.

NCCL itself is covered by the real multi-GPU tests (tests/test_gpu_multi.py: N = 2 all cases, N = 4 three cases, N = 8 QHDFoam and
PCG; profiles/r02*_multi_*).  Written after the round's GPU budget was spent: first device run = the driver's round-end suite,
sorted last, non-strict xfail."""
import os
import subprocess
import sys
import tempfile
import time

import pytest

from test_fake_nccl_cpu import FAKE, ROOT, build_fake_nccl
from first_run_common import BUDGET, budget_ok, first_run

pytestmark = pytest.mark.gpu

SETS = {
    8: ["perturbed_mixed_serialrule,uniform_adjust_procrule", "2d_qgdflux_serialrule,truncoct_mixed_serialrule,slip_perturbed_serialrule"],
    4: ["perturbed_mixed_serialrule,prism_fixed_serialrule,uniform_zg_procrule", "varSc7_fixed_serialrule,2d_leastSquares_serialrule",
        "perturbed_mixed_implicit_serialrule,2d_qgdflux_implicit_adjust_serialrule"],
    2: ["perturbed_mixed_serialrule,uniform_adjust_procrule,truncoct_mixed_serialrule"],
}


@first_run
@pytest.mark.parametrize("world,names", [(w, n) for w, sets in SETS.items() for n in sets])
def test_decomposed_qgdfoam_on_one_gpu_matches_oracle(world, names):
    assert budget_ok(), "first-run budget used up by earlier cases (time-outs / 20 minutes): not started"
    build_fake_nccl()
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = FAKE + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    env.pop("FAKE_NCCL_HOST", None)
    with tempfile.TemporaryDirectory(prefix="qgd_loopback_") as rdv:
        logs = [open(os.path.join(rdv, f"log_{r}.txt"), "w+") for r in range(world)]
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "loopback_worker.py"), str(r), str(world), rdv, names],
                                  stdout=logs[r], stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
        t0, why = time.time(), ""
        while any(p.poll() is None for p in procs):
            if any(p.poll() not in (None, 0) for p in procs):
                why = "a rank failed"                       # do not let its neighbours wait for it
                break
            if time.time() - t0 > 300:
                why = "a rank blocked in an exchange its neighbours never entered (timeout)"
                BUDGET["timeouts"] += 1
                break
            time.sleep(0.2)
        for p in procs:
            if p.poll() is None:
                p.kill()
            p.wait()
        BUDGET["spent"] += time.time() - t0
        outs = []
        for f in logs:
            f.seek(0)
            outs.append(f.read())
            f.close()
    print(outs[0][-4000:])
    assert not why and all(p.returncode == 0 for p in procs) and "LOOPBACK_ALL_OK" in outs[0], why + "\n" + "\n".join(o[-1500:] for o in outs)
