"""Index logic of the chunk pipeline of the host calls (ElementAssembly._increment_pipelined), on CPU: the upload pieces tile the dof
vector exactly once and in order, every chunk finds the node planes it reads ([bounds[c] - 1, bounds[c + 1]], include/edelweiss_b200.h)
already uploaded, and the download pieces tile P exactly once."""
import pytest

from edelweissfe_b200.assembly import pipeline_pieces


@pytest.mark.parametrize("bounds", [[0, 34, 68, 101], [0, 17, 34, 51, 68, 85, 101], [0, 9, 18, 26], [0, 5, 9], [0, 4, 8, 12, 13], [0, 7]])
@pytest.mark.parametrize("pd", [3, 3 * 101 * 101])
def test_pieces_cover_and_order(bounds, pd):
    n_planes = bounds[-1]
    pieces = pipeline_pieces(bounds, pd)
    assert len(pieces) == len(bounds) - 1
    uploaded = 0  # dofs [0, uploaded) are on the device
    down = 0
    for c, (lo, hi, a, b) in enumerate(pieces):
        assert lo == uploaded or lo == hi  # contiguous, no gap, nothing twice
        assert hi >= lo
        uploaded = max(uploaded, hi)
        need_hi = min(bounds[c + 1] + 1, n_planes) * pd  # chunk c reads up to node plane bounds[c + 1] (inclusive)
        assert uploaded >= need_hi
        assert (a, b) == (bounds[c] * pd, bounds[c + 1] * pd) and a == down
        down = b
    assert uploaded == n_planes * pd and down == n_planes * pd
