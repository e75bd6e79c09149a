"""CPU-only: the index rules of the half-block scratch (edelweissfe_b200/csrc/ewb_generic.cuh, HalfLayout) are consistent.

Three device routines share the layout and each derives the location of node pair (a, b) on its own:
  * the scalar element loop (nodeRow + HalfEmit::pass): thread a writes K[a][a+d], d = 0 .. nb(a)-1, at S[a][d];
  * the tensor-pipe phase B (phaseBDmma20LE): every unordered pair a <= b once, directly or transposed under b;
  * the row gather (rowGatherHalfKernel): reads pair (a, b) for every ordered pair.
This restates the three rules in Python and checks that every read hits a written slot with the right orientation."""
import pytest


def nb(a, NN):
    return NN // 2 + (1 if a < NN // 2 else 0)


def writer_circulant(NN):
    """{(row node, slot d): (i-node, j-node)}: stored block is K[i-node][j-node]"""
    out = {}
    for a in range(NN):
        for d in range(nb(a, NN)):
            out[(a, d)] = (a, (a + d) % NN)
    return out


def writer_dmma(NN):
    out = {}
    for a in range(NN):
        for b in range(a, NN):
            d = b - a
            if d < nb(a, NN):
                key, val = (a, d), (a, b)          # stored as K[a][b]
            else:
                key, val = (b, NN - d), (b, a)     # stored transposed: K[b][a]
            assert key not in out
            out[key] = val
    return out


def gather_read(a, b, NN):
    """slot the gather reads for the ordered pair (a, b) and whether it transposes"""
    d = (b - a) % NN
    if d < nb(a, NN):
        return (a, d), False
    return (b, NN - d), True


@pytest.mark.parametrize("NN", [8, 20])
def test_writers_agree_and_gather_finds_every_pair(NN):
    wc, wd = writer_circulant(NN), writer_dmma(NN)
    assert wc == wd, "scalar and tensor-pipe element loops must fill the same slots with the same orientation"
    assert len(wc) == NN * (NN + 1) // 2  # every unordered pair exactly once
    for (a, d) in wc:
        assert 0 <= d < NN // 2 + 1  # inside the SA = 10 (NN/2 + 1) doubles of node a
    for a in range(NN):
        for b in range(NN):
            key, transposed = gather_read(a, b, NN)
            assert key in wc, (a, b, key)
            stored = wc[key]
            assert stored == ((b, a) if transposed else (a, b))
