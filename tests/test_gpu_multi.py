"""2- and 4-GPU parity of the slab-partitioned assembly (fused peer transfer / NCCL exchange + interface add) against the
single-GPU result (edelweissfe_b200.partition.slab_parity_check, also run by bench.py --gpus N before timing).
Skipped on boxes with fewer GPUs."""
import os
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, q, exchange, material, n=(9, 8, 7), chunks=None):
    import torch
    import torch.distributed as dist

    from edelweissfe_b200.partition import slab_parity_check

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    os.environ["EWB_EXCHANGE"] = exchange
    if chunks:
        os.environ["EWB_CHUNKS"] = str(chunks)  # x-chunks of the fused kernel (read at plan creation): exercises the chunk pipeline
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        props = {"vonmises": (2.1e4, 0.22, 355.0, 1000.0, 200.0, 1400.0), "linearelastic": (2.1e4, 0.22)}[material]
        err = slab_parity_check(world, rank, torch.device("cuda", rank), n=n, material=material, props=props)
        q.put((rank, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,chunks", [(2, (9, 8, 7), None), (4, (9, 8, 7), None), (2, (19, 8, 7), 2), (4, (35, 8, 7), 2)])
# 4: ranks with BOTH neighbours (receive from below, store to the rank above); chunks=2: the host calls' transfers are pipelined
# over two x-chunks per rank (slab_parity_check compares them with the device-resident assembly)
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("material", ["vonmises", "linearelastic"])  # first-generation sweep / row-pipelined kernel
def test_slabs_match_single_gpu(exchange, world, n, chunks, material):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q, exchange, material, n, chunks)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        results = dict(q.get(timeout=240) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:  # never leave a rank behind in a collective
            if p.is_alive():
                p.kill()
    assert results[0] is not None and results[0] < 1e-12
