"""The QGDFoam step as a CUDA graph (opt-in, QGD_STEP_GRAPH=1; SURVEY 8b "step is a CUDA graph"): captured once per
qgd_qgdfoam_step call from the ordinary launch sequence - both streams - and replayed; fields must be bit-identical to the stream
launches.  Written after the round's GPU budget was spent: first device run = the driver's round-end suite, own process per case,
sorted last, non-strict xfail (see tests/test_zzz_1_gpu_varsc5.py)."""
import pytest

from first_run_worker import GRAPH
from first_run_common import first_run, run_isolated

pytestmark = pytest.mark.gpu


@first_run
@pytest.mark.parametrize("name", list(GRAPH))
def test_graph_replay_is_bit_identical_to_stream_launches(name):
    run_isolated("graph", name)
