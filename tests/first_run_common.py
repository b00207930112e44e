"""Shared by the first-run device test files (tests/test_zzz_*_gpu_*.py): code that has never run on a device is started in its own
process, marked non-strict xfail, and guarded by a session-wide circuit breaker."""
import os
import subprocess
import sys
import time

import pytest

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


