"""varScModel5 on the device (varScModel5.C:52-269; qgdsolver_b200/csrc/qgd_varsc5.*) against the CPU oracle through the C ABI.

Written after the round's GPU budget was spent: the per-item device code and its launch sequences are verified on CPU under a
serial executor (tests/test_varsc5_host_cpu.py); the CUDA executor, the C-ABI plumbing and the interplay with the ordinary step
kernels have their first device run in the driver's round-end suite.  Each case therefore runs in its own process
(tests/first_run_worker.py: a crash or hang stays contained), the file sorts after every device-verified test file, and the cases
are non-strict xfail so that a first-run surprise cannot mask the verified suite; an XPASS is the parity evidence."""
import pytest

from first_run_common import first_run, run_isolated
from first_run_worker import VARSC5

pytestmark = pytest.mark.gpu


@first_run
@pytest.mark.parametrize("name", list(VARSC5))
def test_varsc5_steps_match_oracle(name):
    run_isolated("varsc5", name)


@first_run
def test_varsc5_refusals_and_bookkeeping():
    run_isolated("varsc5_refusals")
