"""GPU parity tests proper: the CUDA path (through the C ABI) against the reference's golden vectors
and against the oracle on seeded inputs.  Tolerance: 1e-12 relative (max-abs / max-abs) for K, P, F,
state (north_star); CSR pattern, slot map and dof indexing bit-exact."""
import numpy as np
import pytest
from conftest import golden_names, load_golden, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _assembly(g, box=None):
    from edelweissfe_b200 import ElementAssembly

    return ElementAssembly(str(g["elType"]), g["conn"], g["coords"], str(g["material"]), g["props"], box=box)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_golden(name, path):
    import torch

    from edelweissfe_b200 import _lib

    g = load_golden(name)
    box = None
    if path == "auto":
        if not (bool(g["boxgen_regular"]) or True) or g["conn"].shape[1] != 8:
            pytest.skip("structured sweep is Hexa8 only")
        box = [int(v) for v in g["box"][:3]]  # BoxGen topology (coordinates may be distorted)
    asm = _assembly(g, box)
    indptr, indices = asm.csr_pattern()
    assert indptr.dtype == torch.int32 and indices.dtype == torch.int32
    assert np.array_equal(indptr.cpu().numpy(), g["indptr"])
    assert np.array_equal(indices.cpu().numpy(), g["indices"])
    # COO -> CSR slot map against the reference pattern: indices[x[p]] == J[p] and row(x[p]) == I[p]
    x = asm.slot_map().cpu().numpy()
    assert np.array_equal(g["indices"][x], g["J"].astype(np.int32))
    assert np.array_equal(np.searchsorted(g["indptr"], x, side="right") - 1, g["I"])
    flags = _lib.EWB_FLAG_FORCE_GENERIC if path == "generic" else 0
    for p in range(int(g["nPasses"])):
        asm.U.copy_(torch.as_tensor(g[f"U{p}"]))
        asm.dU.copy_(torch.as_tensor(g[f"dU{p}"]))
        asm.set_state_aos(g[f"stateRef{p}"], "ref")
        asm.assemble(flags)
        asm.poll()
        assert relerr(asm.csr_data.cpu().numpy(), g[f"data{p}"]) < TOL
        assert relerr(asm.P.cpu().numpy(), g[f"P{p}"]) < TOL
        assert relerr(asm.F.cpu().numpy(), g[f"F{p}"]) < TOL
        assert relerr(asm.state_aos("temp").cpu().numpy(), g[f"stateTemp{p}"]) < TOL
        if path == "generic" and f"V{p}" in g:
            V, Pe = asm.compute_elements_vij()
            assert relerr(V.cpu().numpy(), g[f"V{p}"]) < TOL
            # reference-order CSR accumulation of the reference's own V is reproduced bit for bit
            data = asm.update_csr(torch.as_tensor(g[f"V{p}"]).to(asm.device)).cpu().numpy()
            assert np.array_equal(data, g[f"data{p}"])


@pytest.mark.parametrize(
    "elType,material,props,n,scale",
    [
        ("C3D8", "linearelastic", [2.1e4, 0.22], (7, 9, 11), 1e-3),
        ("C3D8", "vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400], (9, 8, 7), 5e-3),
        ("C3D8TL", "neohookewa", [91304.34783, 100000.0], (6, 7, 8), 2e-2),
        ("C3D8", "linearelastic", [2.1e4, 0.22], (25, 16, 9), 1e-3),  # several x-chunks and y/z tiles, ragged last tiles
        ("C3D8", "vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400], (26, 8, 15), 5e-3),
        ("C3D8TL", "neohookewb", [91304.34783, 100000.0], (5, 8, 7), 2e-2),
        ("C3D8TL", "neohookewc", [91304.34783, 100000.0], (24, 7, 6), 2e-2),
        ("C3D20", "linearelastic", [2.1e4, 0.22], (3, 4, 3), 1e-3),
        ("C3D20", "vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400], (3, 3, 4), 5e-3),
    ],
)
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_against_oracle_seeded(elType, material, props, n, scale, path):
    import torch

    from edelweissfe_b200 import ElementAssembly, _lib, box_mesh
    from oracle import port

    nn = 20 if "20" in elType else 8
    if path == "auto" and nn != 8:
        pytest.skip("structured sweep is Hexa8 only")
    coords, conn = box_mesh(*n, lX=float(n[0]), lY=1.1 * n[1], lZ=0.9 * n[2], elType=elType)
    c2, conn2 = port.boxgen(*n, float(n[0]), 1.1 * n[1], 0.9 * n[2], nnodes=nn)
    assert np.array_equal(conn, conn2) and np.array_equal(coords, c2)
    rng = np.random.default_rng(1)
    if nn == 8:
        coords = coords + 0.15 * rng.uniform(-1, 1, coords.shape)
    asm = ElementAssembly(elType, conn, coords, material, props, box=n if path == "auto" else None)
    flags = _lib.EWB_FLAG_FORCE_GENERIC if path == "generic" else 0
    nGp = 27 if nn == 20 else 8
    state = np.zeros((conn.shape[0], nGp, 12 + port.MATERIAL_NSTATE[material]))
    U = np.zeros(3 * coords.shape[0])
    for p in range(2):
        dU = scale * rng.standard_normal(U.size)
        U = U + dU
        o = port.assemble(elType, material, props, coords, conn, U, dU, state, want_vij=False)
        asm.U.copy_(torch.as_tensor(U))
        asm.dU.copy_(torch.as_tensor(dU))
        asm.assemble(flags)
        asm.poll()
        if p == 0:
            ip, ix = asm.csr_pattern()
            assert np.array_equal(ip.cpu().numpy(), o["indptr"]) and np.array_equal(ix.cpu().numpy(), o["indices"])
        assert relerr(asm.csr_data.cpu().numpy(), o["data"]) < TOL
        assert relerr(asm.P.cpu().numpy(), o["P"]) < TOL
        assert relerr(asm.F.cpu().numpy(), o["F"]) < TOL
        assert relerr(asm.state_aos("temp").cpu().numpy(), o["stateTemp"]) < TOL
        asm.accept_last_state()
        state = o["stateTemp"]


def test_dirichlet_rows():
    import torch

    from edelweissfe_b200 import ElementAssembly, box_mesh

    coords, conn = box_mesh(3, 3, 3)
    asm = ElementAssembly("C3D8", conn, coords, "linearelastic", [2.1e4, 0.22])
    asm.dU.copy_(torch.as_tensor(1e-3 * np.random.default_rng(0).standard_normal(asm.nDof)))
    asm.assemble()
    asm.poll()
    K0 = asm.to_scipy()
    dofs = np.array([0, 5, 17, 100], dtype=np.int32)
    asm.apply_dirichlet_k(dofs)
    K1 = asm.to_scipy()
    assert np.array_equal(K0.indices, K1.indices)
    D = K1.toarray()
    K0d = K0.toarray()
    for d in dofs:
        row = np.zeros(asm.nDof)
        row[d] = 1.0
        assert np.array_equal(D[d], row)
    keep = np.setdiff1d(np.arange(asm.nDof), dofs)
    assert np.array_equal(D[keep], K0d[keep])


def test_cutback_flag():
    """A hardening law that cannot converge in 15 updates must surface as CutbackRequest(…, 0.5)."""
    import torch

    from edelweissfe_b200 import CutbackRequest, ElementAssembly, box_mesh
    from oracle import port

    props = [2.1e4, 0.22, 355.0, -5e4, 200.0, 1400.0]
    coords, conn = box_mesh(2, 2, 2)
    rng = np.random.default_rng(0)
    dU = 0.05 * rng.standard_normal(3 * coords.shape[0])
    state = np.zeros((8, 8, 13))
    o = port.assemble("C3D8", "vonmises", props, coords, conn, dU, dU, state, want_vij=False)
    asm = ElementAssembly("C3D8", conn, coords, "vonmises", props)
    asm.dU.copy_(torch.as_tensor(dU))
    asm.assemble()
    if o["failed"].any():
        with pytest.raises(CutbackRequest) as ei:
            asm.poll()
        assert ei.value.cutbackSize == 0.5
    else:
        asm.poll()
