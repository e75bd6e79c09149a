"""CPU-only: bench.py's algorithmic-bytes formula reproduces the per-element figures of SURVEY.md §8(d)."""
import importlib.util
import os

import pytest

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize(
    "nn,nGp,nState,n,nnz,nNode,expected",
    [
        (8, 8, 12, (100, 100, 100), 245438109, 1030301, 3655.0),       # config 2: C3D8 linear elastic
        (8, 8, 13, (200, 100, 100), 490060809, 2050401, 3779.0),       # config 3: C3D8 von Mises
        (20, 27, 12, (100, 100, 50), 1061478009, 2060501, 22742.0),    # config 4: C3D20 linear elastic
        (8, 8, 13, (200, 200, 200), 1953736209, 8120601, 3771.0),      # config 5: C3D8TL Neo-Hooke
    ],
)
def test_algorithmic_bytes_match_survey(nn, nGp, nState, n, nnz, nNode, expected):
    b = _bench()
    nEl = n[0] * n[1] * n[2]
    got = b.algorithmic_bytes_per_element(nn, nGp, nState, nnz, nEl, nNode)
    assert abs(got - expected) < 1.0, got


def test_box_counts_closed_form():
    """nnz of a Hexa8 box: 9 (3nX+1)(3nY+1)(3nZ+1) — the numbers used above."""
    assert 9 * 301 * 301 * 301 == 245438109
    assert 9 * 601 * 301 * 301 == 490060809
    assert 9 * 601 * 601 * 601 == 1953736209
