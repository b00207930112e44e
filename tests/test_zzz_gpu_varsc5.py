"""varScModel5 on the device (varScModel5.C:52-269; qgdsolver_b200/csrc/qgd_varsc5.*) against the CPU oracle through the C ABI.

Written after the round's GPU budget was spent: the per-item device code and its launch sequences are verified on CPU under a
serial executor (tests/test_varsc5_host_cpu.py); the CUDA executor, the C-ABI plumbing and the interplay with the ordinary step
kernels have their first device run in the driver's round-end suite.  Each case therefore runs in its own process
(tests/first_run_worker.py: a crash or hang stays contained), the file sorts after every device-verified test file, and the cases
are non-strict xfail so that a first-run surprise cannot mask the verified suite; an XPASS is the parity evidence."""
import os
import subprocess
import sys
import time

import pytest

from first_run_worker import VARSC5

pytestmark = pytest.mark.gpu
first_run = pytest.mark.xfail(strict=False, reason="first device run is the driver's round-end suite (GPU budget of the round spent)")
WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "first_run_worker.py")


# circuit breaker shared by every first-run file: never-run device code must not be able to stall the session - after two time-outs
# or 20 minutes spent in first-run cases the remaining ones fail at once
BUDGET = {"timeouts": 0, "spent": 0.0}


def budget_ok():
    return BUDGET["timeouts"] < 2 and BUDGET["spent"] < 1200.0


def run_isolated(*args, timeout=300):
    assert budget_ok(), "first-run budget used up by earlier cases (time-outs / 20 minutes): not started"
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, WORKER, *args], capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        BUDGET["timeouts"] += 1
        raise
    finally:
        BUDGET["spent"] += time.time() - t0
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "FIRST_RUN_OK" in r.stdout, (r.stdout[-3000:] + r.stderr[-3000:])


@first_run
@pytest.mark.parametrize("name", list(VARSC5))
def test_varsc5_steps_match_oracle(name):
    run_isolated("varsc5", name)


@first_run
def test_varsc5_refusals_and_bookkeeping():
    run_isolated("varsc5_refusals")
