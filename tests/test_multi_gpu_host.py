"""N>1 path, host-side logic, on CPU with the gloo backend (world_size 2 and 3): slab layout, the
neighbour exchange of the ghost-plane tail and the interface add reproduce the single-domain assembly.
Local assemblies come from the oracle (tests may use it); the CUDA interface-add kernel is mirrored by
partition.interface_add_host, which shares the index logic (SlabLayout) with the product path."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from edelweissfe_b200.boxgen import box_mesh
from edelweissfe_b200.partition import SlabLayout, exchange_tails, interface_add_host, slab_ranges
from oracle import port

N = (5, 3, 2)
PROPS = [2.1e4, 0.22]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coords, conn = box_mesh(*N, lX=5.0, lY=3.3, lZ=2.2)
        rng = np.random.default_rng(0)
        coords = coords + 0.1 * rng.uniform(-1, 1, coords.shape)
        dU = 1e-3 * rng.standard_normal(3 * coords.shape[0])
        lay = SlabLayout(N[0], N[1], N[2], rank, world)
        # local mesh = slice of the global one (exact same coordinates)
        _, lconn = box_mesh(lay.nXloc, N[1], N[2])
        n0 = lay.node_offset()
        lcoords = coords[n0 : n0 + lay.nNodeLoc]
        assert np.array_equal(lconn + n0, conn[lay.a * N[1] * N[2] : lay.b * N[1] * N[2]])
        ldU = dU[3 * n0 : 3 * (n0 + lay.nNodeLoc)]
        state = np.zeros((lconn.shape[0], 8, 12))
        o = port.assemble("C3D8", "linearelastic", PROPS, lcoords, lconn, ldU, ldU, state, want_vij=False)
        data, P, F = (torch.from_numpy(o[k].copy()) for k in ("data", "P", "F"))
        indptr = o["indptr"]
        head = lay.head_nnz(indptr)
        recv = torch.zeros(head if lay.has_lower else 1, dtype=torch.float64)
        rP = torch.zeros(lay.planeDofs if lay.has_lower else 1, dtype=torch.float64)
        rF = torch.zeros_like(rP)
        exchange_tails(lay, data, P, F, recv, rP, rF, indptr, dist)
        # test-only: also ship the column indices of the tail so that the halo block can be placed globally
        idx = torch.from_numpy(o["indices"].astype(np.int64) + 3 * n0)
        ridx = torch.zeros(head if lay.has_lower else 1, dtype=torch.int64)
        ops = []
        if lay.has_upper:
            ops.append(dist.P2POp(dist.isend, idx[lay.tail_start(indptr):].contiguous(), rank + 1))
        if lay.has_lower:
            ops.append(dist.P2POp(dist.irecv, ridx, rank - 1))
        for r in dist.batch_isend_irecv(ops) if ops else []:
            r.wait()
        data, P, F = data.numpy(), P.numpy(), F.numpy()
        if lay.has_lower:
            interface_add_host(indptr, lay.planeDofs, data, recv.numpy(), P, F, rP.numpy(), rF.numpy())
        # owned rows in global numbering
        nG = 3 * coords.shape[0]
        rows, cols, vals = [], [], []
        for r in range(lay.ownedDofs):
            sl = slice(indptr[r], indptr[r + 1])
            rows += [r + 3 * n0] * (indptr[r + 1] - indptr[r])
            cols += list(o["indices"][sl].astype(np.int64) + 3 * n0)
            vals += list(data[sl])
            if lay.has_lower and r < lay.planeDofs:  # lower halo block = first half of the received row
                half = (indptr[r + 1] - indptr[r]) // 2
                hs = slice(indptr[r], indptr[r] + half)
                rows += [r + 3 * n0] * half
                cols += list(ridx.numpy()[hs])
                vals += list(recv.numpy()[hs])
        q.put((rank, np.array(rows), np.array(cols), np.array(vals), P[: lay.ownedDofs].copy(), F[: lay.ownedDofs].copy(), 3 * n0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_exchange_reproduces_global_assembly(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    coords, conn = box_mesh(*N, lX=5.0, lY=3.3, lZ=2.2)
    rng = np.random.default_rng(0)
    coords = coords + 0.1 * rng.uniform(-1, 1, coords.shape)
    dU = 1e-3 * rng.standard_normal(3 * coords.shape[0])
    g = port.assemble("C3D8", "linearelastic", PROPS, coords, conn, dU, dU, np.zeros((conn.shape[0], 8, 12)), want_vij=False)
    nG = 3 * coords.shape[0]
    Kg = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(nG, nG))
    Kd = sp.csr_matrix((nG, nG))
    Pd, Fd = np.zeros(nG), np.zeros(nG)
    covered = np.zeros(nG, dtype=int)
    for rank, rows, cols, vals, P, F, off in results:
        Kd = Kd + sp.csr_matrix((vals, (rows, cols)), shape=(nG, nG))
        Pd[off : off + P.size] = P
        Fd[off : off + F.size] = F
        covered[off : off + P.size] += 1
    assert (covered == 1).all()  # every dof owned exactly once
    scale = np.abs(g["data"]).max()
    assert abs(Kd - Kg).max() / scale < 1e-13
    assert np.abs(Pd - g["P"]).max() / np.abs(g["P"]).max() < 1e-13
    assert np.abs(Fd - g["F"]).max() / np.abs(g["F"]).max() < 1e-13


def test_slab_ranges():
    assert slab_ranges(100, 8) == [(0, 13), (13, 26), (26, 39), (39, 52), (52, 64), (64, 76), (76, 88), (88, 100)]
    assert slab_ranges(8, 8)[-1] == (7, 8)
    lay = SlabLayout(10, 3, 2, 1, 3)
    assert (lay.a, lay.b, lay.nXloc) == (4, 7, 3)
    assert lay.ownedDofs == lay.nDofLoc - lay.planeDofs and lay.node_offset() == 4 * 12


def test_morton_order_is_a_locality_preserving_permutation():
    """Host helper behind ewb_plan_set_gather_order: a permutation, and consecutive nodes stay spatially close."""
    import numpy as np

    from edelweissfe_b200 import box_mesh
    from edelweissfe_b200.assembly import morton_order

    coords, _ = box_mesh(7, 6, 5, lX=7.0, lY=6.0, lZ=5.0, elType="C3D20")
    order = morton_order(coords)
    assert order.dtype == np.int32 and sorted(order.tolist()) == list(range(coords.shape[0]))
    step = np.linalg.norm(np.diff(coords[order], axis=0), axis=1)
    natural = np.linalg.norm(np.diff(coords, axis=0), axis=1)
    assert np.median(step) <= np.median(natural) * 1.5 and step.mean() < 0.5 * np.linalg.norm(coords.max(0) - coords.min(0))
    # degenerate input (all nodes in a plane) must not divide by zero
    flat = coords.copy()
    flat[:, 2] = 1.0
    assert sorted(morton_order(flat).tolist()) == list(range(coords.shape[0]))
