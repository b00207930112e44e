"""Axisymmetric (wedge) and symmetryPlane-patch QGDFoam cases on the device against the CPU oracle through the C ABI: the wedge velocity condition
(QGD_BC_WEDGE, k_wedge_bnd), the wedge vertex constraint (k_wedge_points, k_wedge_points_generic) and the 2D
GaussVolPoint path on a wedge mesh whose every vertex is a patch point.

Written after the round's GPU budget was spent (first device run = the driver's round-end suite; CPU-side evidence:
tests/test_wedge_host_cpu.py, tests/test_oracle_wedge.py).  Each case runs in its own process (tests/first_run_worker.py), the
file sorts after every device-verified file, non-strict xfail."""
import pytest

from first_run_worker import WEDGE
from first_run_common import first_run, run_isolated

pytestmark = pytest.mark.gpu


@first_run
@pytest.mark.parametrize("name", list(WEDGE))
def test_wedge_steps_match_oracle(name):
    run_isolated("wedge", name)


@first_run
def test_fvsc_operators_on_a_wedge_mesh_match_oracle():
    run_isolated("wedge_ops")


@first_run
def test_wedge_refusals():
    run_isolated("wedge_refusals")
