"""The oracle (oracle/port.py) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
from conftest import golden_names, load_golden, relerr

from oracle import port

TOL = 1e-12  # north_star: K, P, state within 1e-12 relative; pattern/indexing bit-exact


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    elType, material = str(g["elType"]), str(g["material"])
    coords, conn = g["coords"], g["conn"]
    dofs = port.element_dofs(conn)
    assert np.array_equal(dofs, g["element_dofs"])  # DofManager numbering, bit-exact
    I, J = port.vij_pattern(dofs)
    assert np.array_equal(I, g["I"]) and np.array_equal(J, g["J"])  # VIJ layout, bit-exact
    for p in range(int(g["nPasses"])):
        o = port.assemble(elType, material, g["props"], coords, conn, g[f"U{p}"], g[f"dU{p}"], g[f"stateRef{p}"])
        assert o["indptr"].dtype == np.int32 and o["indices"].dtype == np.int32
        assert np.array_equal(o["indptr"], g["indptr"]) and np.array_equal(o["indices"], g["indices"])
        assert relerr(o["data"], g[f"data{p}"]) < TOL
        assert relerr(o["P"], g[f"P{p}"]) < TOL
        assert relerr(o["F"], g[f"F{p}"]) < TOL
        assert relerr(o["stateTemp"], g[f"stateTemp{p}"]) < TOL
        if f"V{p}" in g:
            assert relerr(o["V"], g[f"V{p}"]) < TOL
        assert not o["failed"].any()


@pytest.mark.parametrize("name", [n for n in golden_names() if n.endswith("_box")])
def test_boxgen_restatement(name):
    g = load_golden(name)
    nX, nY, nZ, lX, lY, lZ = g["box"]
    nn = g["conn"].shape[1]
    coords, conn = port.boxgen(int(nX), int(nY), int(nZ), lX, lY, lZ, nnodes=nn)
    assert np.array_equal(conn, g["conn"])
    assert np.array_equal(coords, g["coords"])  # np.linspace layers, bit-exact


def test_hexa8_nnz_closed_form():
    # nnz = 9 (3nX+1)(3nY+1)(3nZ+1)
    for n in [(1, 1, 1), (2, 3, 4), (5, 2, 3)]:
        coords, conn = port.boxgen(*n)
        I, J = port.vij_pattern(port.element_dofs(conn))
        indptr, indices, x = port.csr_pattern(I, J, 3 * coords.shape[0])
        assert indices.size == 9 * (3 * n[0] + 1) * (3 * n[1] + 1) * (3 * n[2] + 1)


def test_von_mises_newton_failure_flag():
    # a hardening law with a huge negative exponent slope cannot converge in 15 updates
    props = [2.1e4, 0.22, 355.0, -5e4, 200.0, 1400.0]
    stress = np.zeros((1, 1, 6))
    de = np.zeros((1, 1, 6))
    de[..., 3] = 0.5
    s, C, k, failed = port.von_mises(props, stress, de, np.zeros((1, 1)))
    assert failed.shape == (1, 1)


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_matches_reference_golden(name):
    """oracle/element_loop.c (the OpenMP CPU baseline) against the reference's golden vectors."""
    import subprocess

    from conftest import ROOT

    subprocess.check_call(["make", "-s", "-C", ROOT + "/oracle"])
    from oracle import cport

    c = cport.load()
    g = load_golden(name)
    for p in range(int(g["nPasses"])):
        o = c.assemble(str(g["elType"]), str(g["material"]), g["props"], g["coords"], g["conn"], g[f"U{p}"], g[f"dU{p}"], g[f"stateRef{p}"])
        assert np.array_equal(o["indptr"], g["indptr"]) and np.array_equal(o["indices"], g["indices"])
        for key in ("data", "P", "F", "stateTemp"):
            assert relerr(o[key], g[f"{key}{p}"]) < TOL, key
        if f"V{p}" in g:
            assert relerr(o["V"], g[f"V{p}"]) < TOL
            # the reference's sequential updateCSR order is reproduced bit for bit on the reference's own V
            data = np.empty(g["indices"].size)
            x = np.ascontiguousarray(o["x"], dtype=np.int32)
            V = np.ascontiguousarray(g[f"V{p}"])
            c.lib.ewo_update_csr(x.size, x.ctypes.data, V.ctypes.data, data.ctypes.data, data.size)
            assert np.array_equal(data, g[f"data{p}"])


@pytest.mark.parametrize("name", golden_names())
def test_oracle_body_force(name):
    g = load_golden(name)
    P, _ = port.body_force(str(g["elType"]), g["coords"], g["conn"], g["bodyforce_load"])
    assert relerr(P, g["bodyforce_PExt"]) < TOL
