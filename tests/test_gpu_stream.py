"""Task-stream kernel of the arbitrary-mesh path (20-node hexahedra, csrc/ewb_stream.cuh): element loop and row gather as warp
tasks of one persistent launch.  It must reproduce the two-phase path (separate kernels, EWB_FLAG_TWO_PHASE) BIT FOR BIT — same
blocks, same ascending-element summation per node (csrgenerator.pyx:100-115) — for every element order, chunk size and gather
delay, including the degenerate ones that make every gather task wait on the element tasks right before it."""
import ctypes as C

import numpy as np
import pytest
from conftest import entrywise, relerr

pytestmark = pytest.mark.gpu

VM = [2.1e4, 0.22, 355, 1000, 200, 1400]
LE = [2.1e4, 0.22]


def _run(elType, material, props, n, scale, env, monkeypatch, order="morton", flags=0, passes=2, seed=3):
    import torch

    from edelweissfe_b200 import ElementAssembly, box_mesh

    for k in ("EWB_STREAM", "EWB_STREAM_DISCARD", "EWB_STREAM_CHUNK", "EWB_STREAM_DELAY", "EWB_STREAM_EPT", "EWB_STREAM_NPT", "EWB_C3D20_ROWS"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("EWB_STREAM", "1")  # the task-stream kernel is opt-in
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))
    monkeypatch.setenv("EWB_ELEMENT_ORDER", "morton" if order == "morton" else "none")
    coords, conn = box_mesh(*n, lX=float(n[0]), lY=float(n[1]), lZ=float(n[2]), elType=elType)
    rng = np.random.default_rng(seed)
    coords = coords + 0.05 * rng.standard_normal(coords.shape) * (elType == "C3D20R")  # C3D20: affine elements only
    asm = ElementAssembly(elType, conn, coords, material, props)
    if order == "random":
        perm = np.ascontiguousarray(np.random.default_rng(seed + 100).permutation(conn.shape[0]).astype(np.int32))
        assert asm.lib.ewb_plan_set_element_order(asm.plan, perm.ctypes.data_as(C.c_void_p)) == 0
        bad = perm.copy()
        bad[0] = bad[1]
        assert asm.lib.ewb_plan_set_element_order(asm.plan, bad.ctypes.data_as(C.c_void_p)) < 0  # not a permutation
    out = []
    for it in range(passes):
        dU = scale * rng.standard_normal(asm.nDof)
        asm.dU.copy_(torch.as_tensor(dU))
        asm.U.add_(asm.dU)
        if flags & 1:  # EWB_FLAG_ACCUMULATE_PF: caller-provided start values
            asm.P.fill_(0.25)
            asm.F.fill_(0.5)
        asm.csr_data.fill_(float("nan"))
        asm.assemble(flags=flags)
        asm.poll() if material == "linearelastic" else _poll_ok(asm)
        out.append([t.cpu().numpy().copy() for t in (asm.csr_data, asm.P, asm.F, asm.state_temp)])
        asm.accept_last_state()
    return out


def _poll_ok(asm):
    from edelweissfe_b200.assembly import CutbackRequest

    try:
        asm.poll()
    except CutbackRequest:
        pytest.fail("unexpected cut-back request")


CASES = [
    ("C3D20", "linearelastic", LE, (5, 4, 3), 1e-3),
    ("C3D20", "vonmises", VM, (3, 4, 3), 4e-3),
    ("C3D20R", "linearelastic", LE, (4, 3, 5), 1e-3),
    ("C3D20R", "vonmises", VM, (3, 3, 4), 4e-3),
]


@pytest.mark.parametrize("elType,material,props,n,scale", CASES)
@pytest.mark.parametrize(
    "env,order",
    [
        ({}, "morton"),
        ({}, "none"),
        ({"EWB_STREAM_CHUNK": 1, "EWB_STREAM_DELAY": 0, "EWB_STREAM_EPT": 1, "EWB_STREAM_NPT": 1}, "random"),
        ({"EWB_STREAM_CHUNK": 7, "EWB_STREAM_DELAY": 1, "EWB_STREAM_EPT": 3, "EWB_STREAM_NPT": 5, "EWB_STREAM_DISCARD": 0}, "random"),
        ({"EWB_STREAM_CHUNK": 4096, "EWB_STREAM_DELAY": 3}, "morton"),
        ({"EWB_STREAM": 0, "EWB_C3D20_ROWS": 1}, "none"),  # the two task bodies as two ordinary launches over the row scratch
    ],
)
def test_stream_equals_two_phase_bitwise(elType, material, props, n, scale, env, order, monkeypatch):
    ref = _run(elType, material, props, n, scale, {"EWB_STREAM": 0}, monkeypatch, order="none")
    got = _run(elType, material, props, n, scale, env, monkeypatch, order=order)
    for a, b in zip(ref, got):
        for x, y, what in zip(a, b, ("csr_data", "P", "F", "state")):
            if material == "linearelastic":
                assert np.array_equal(x, y), f"{what}: max diff {np.abs(x - y).max()}"
            else:
                # von Mises blocks come from another instantiation of the same routine (the compiler contracts its products into FMAs
                # differently): identical up to rounding, entry-wise
                assert relerr(x, y) < 1e-14 and entrywise(x, y, rel=1e-12, floor=1e-15) <= 1.0, f"{what}: {relerr(x, y)}"


@pytest.mark.parametrize("flags", [1, 4, 5, 16])
def test_stream_flags(flags, monkeypatch):
    """ACCUMULATE_PF, NO_STIFFNESS and the explicit two-phase flag."""
    ref = _run("C3D20", "linearelastic", LE, (4, 3, 3), 1e-3, {"EWB_STREAM": 0}, monkeypatch, order="none", flags=flags & ~16)
    got = _run("C3D20", "linearelastic", LE, (4, 3, 3), 1e-3, {}, monkeypatch, flags=flags)
    for a, b in zip(ref, got):
        for x, y, what in zip(a, b, ("csr_data", "P", "F", "state")):
            if what == "csr_data" and flags & 4:
                assert np.isnan(y).all()  # untouched
                continue
            assert np.array_equal(x, y), what


def test_stream_against_oracle(monkeypatch):
    """The stream kernel against the CPU oracle directly (not only against the other CUDA path)."""
    import torch

    from edelweissfe_b200 import ElementAssembly, box_mesh
    from oracle import port

    monkeypatch.setenv("EWB_STREAM", "1")
    n = (4, 5, 3)
    coords, conn = box_mesh(*n, lX=4.0, lY=5.0, lZ=3.0, elType="C3D20")
    rng = np.random.default_rng(11)
    asm = ElementAssembly("C3D20", conn, coords, "linearelastic", LE)
    dU = 1e-3 * rng.standard_normal(asm.nDof)
    asm.dU.copy_(torch.as_tensor(dU))
    asm.U.copy_(asm.dU)
    asm.assemble()
    asm.poll()
    o = port.assemble("C3D20", "linearelastic", LE, coords, conn, dU, dU, np.zeros((conn.shape[0], 27, 12)), want_vij=False)
    K = asm.csr_data.cpu().numpy()
    assert relerr(K, o["data"]) < 1e-12 and entrywise(K, o["data"]) <= 1.0
    assert relerr(asm.state_aos("temp").cpu().numpy(), o["stateTemp"]) < 1e-12
    assert relerr(asm.P.cpu().numpy(), o["P"]) < 1e-12
    assert relerr(asm.F.cpu().numpy(), o["F"]) < 1e-12


def test_stream_is_one_launch(monkeypatch):
    """After the first call (slot table) an assembly is exactly one kernel launch."""
    from edelweissfe_b200 import ElementAssembly, box_mesh

    monkeypatch.setenv("EWB_STREAM", "1")
    coords, conn = box_mesh(3, 3, 3, elType="C3D20")
    asm = ElementAssembly("C3D20", conn, coords, "linearelastic", LE)
    asm.assemble()
    asm.poll()
    n0 = asm.lib.ewb_launch_count()
    asm.assemble()
    asm.poll()
    assert asm.lib.ewb_launch_count() - n0 == 1
