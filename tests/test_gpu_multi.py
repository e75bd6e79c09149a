"""2- and 4-GPU parity of the slab-partitioned assembly (NCCL exchange + interface add) against the single-GPU
result.  Skipped on boxes with one GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = (9, 8, 7)
PROPS = [2.1e4, 0.22, 355.0, 1000.0, 200.0, 1400.0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    from edelweissfe_b200 import box_mesh

    coords, conn = box_mesh(*N, lX=9.0, lY=8.0, lZ=7.0)
    rng = np.random.default_rng(3)
    dU = 4e-3 * rng.standard_normal(3 * coords.shape[0])
    return coords, conn, dU


def _worker(rank, world, port_no, q, exchange):
    import torch
    import torch.distributed as dist

    from edelweissfe_b200.partition import SlabAssembly

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        coords, conn, dU = _inputs()
        slab = SlabAssembly(N, (9.0, 8.0, 7.0), "C3D8", "vonmises", PROPS, rank, world, torch.device("cuda", rank), exchange=exchange)
        assert slab.exchange == exchange, "requested exchange mode is not available: %s" % slab.exchange
        lay, asm = slab.layout, slab.asm
        n0 = lay.node_offset()
        asm.coords.copy_(torch.as_tensor(coords[n0 : n0 + lay.nNodeLoc]))
        ldU = dU[3 * n0 : 3 * (n0 + lay.nNodeLoc)]
        asm.U.copy_(torch.as_tensor(ldU))
        asm.dU.copy_(torch.as_tensor(ldU))
        for _ in range(3):  # repeated assemblies: the double-buffered receive side must stay consistent
            slab.assemble()
        asm.poll()
        rows, nnzs = slab.owned_slices()
        q.put((rank, 3 * n0, slab.indptr_host[: lay.ownedDofs + 1].copy(), slab.indices.cpu().numpy()[nnzs], asm.csr_data.cpu().numpy()[nnzs],
               asm.P.cpu().numpy()[rows], asm.F.cpu().numpy()[rows], slab.recv.cpu().numpy(), lay.planeDofs, lay.has_lower))
        slab.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])  # 4: ranks with BOTH neighbours (receive from below, store to the rank above)
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_two_gpu_slabs_match_single_gpu(exchange, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import scipy.sparse as sp
    import torch.multiprocessing as mp

    from edelweissfe_b200 import ElementAssembly

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        results = [q.get(timeout=240) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:  # never leave a rank behind in a collective
            if p.is_alive():
                p.kill()
    coords, conn, dU = _inputs()
    ref = ElementAssembly("C3D8", conn, coords, "vonmises", PROPS, box=N)
    ref.U.copy_(torch.as_tensor(dU))
    ref.dU.copy_(torch.as_tensor(dU))
    ref.assemble()
    ref.poll()
    Kg = ref.to_scipy()
    Pg, Fg = ref.P.cpu().numpy(), ref.F.cpu().numpy()
    nG = Kg.shape[0]
    scale = abs(Kg).max()
    for rank, off, indptr, indices, data, P, F, recv, planeDofs, has_lower in results:
        nrows = indptr.size - 1
        Kl = sp.csr_matrix((data, indices.astype(np.int64) + off, indptr), shape=(nrows, nG))
        Kref = Kg[off : off + nrows]
        if has_lower:  # the dx=-1 columns live in the halo block: compare them with the received rows
            lowcols = np.arange(off - planeDofs, off)
            halo = Kref[:planeDofs][:, lowcols]
            Kref = Kref.tolil()
            Kref[:planeDofs, lowcols] = 0
            Kref = Kref.tocsr()
            got = []
            for r in range(planeDofs):
                half = (indptr[r + 1] - indptr[r]) // 2
                got.append(recv[indptr[r] : indptr[r] + half])
            halo.sort_indices()
            assert np.abs(np.concatenate(got) - halo.data).max() / scale < 1e-12
        assert abs(Kl - Kref).max() / scale < 1e-12
        assert np.abs(P - Pg[off : off + nrows]).max() / np.abs(Pg).max() < 1e-12
        assert np.abs(F - Fg[off : off + nrows]).max() / np.abs(Fg).max() < 1e-12
