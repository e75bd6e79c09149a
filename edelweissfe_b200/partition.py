"""Multi-GPU slab decomposition of a BoxGen box.

BoxGen numbers elements and nodes ix-major (generators/boxgen.py:133-136,168-170), so contiguous
element-plane ranges are slabs in x and every slab's nodes are a contiguous CSR row range.  Rank g owns
element planes [a_g, a_g+1) and node planes [a_g, a_g+1) (the last rank also the final plane).  Its local
mesh holds one more node plane — the ghost plane a_g+1 owned by rank g+1 — whose partial rows are the
contiguous TAIL of the local CSR value array (and of P, F).  One neighbour exchange per assembly:

    rank g  --(tail of csr_data, P, F)-->  rank g+1

On GPUs the transfer is fused into the sweep kernel ("peer" exchange, the default): the CTAs that finish ghost-plane rows
store them straight into rank g+1's receive buffer (CUDA-IPC mapped peer memory, NVLink) while the rest of the slab is
still being computed; the only collective is a 4-byte max-all-reduce of the status words, which orders the receiver's
interface add after the sender's kernel and gives every rank the same cut-back decision.  The "nccl" exchange (send/recv
of the finished tail, torch.distributed P2P) is the fallback when peer mapping is not available, and what the CPU tests
run over gloo.  Then `ewb_interface_add` adds the dx=0 half of the received rows onto the receiver's first node plane
(own contribution first, neighbour second: deterministic).  The dx=-1 half of the received rows is the
receiver's lower halo block.  No all-reduce is needed for K.

The reference has no multi-process mode; this layout is the B200-native extension of its prange over
elements (solvers/nonlinearimplicitstaticparallelmk2.pyx:157-160).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin this process to the CPU cores that NVML reports as local to the GPU (same NUMA node / PCIe root), BEFORE pinned host
    buffers are allocated: with one rank per GPU, every rank's host<->device traffic then stays on its own socket instead of
    all ranks sharing one memory controller.  Returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:  # noqa: BLE001 - a placement hint only
        return False


def slab_ranges(nX: int, world: int):
    """Element-plane ranges [a_g, a_g+1) per rank, as even as possible."""
    base, rem = divmod(nX, world)
    out, a = [], 0
    for g in range(world):
        b = a + base + (1 if g < rem else 0)
        out.append((a, b))
        a = b
    return out


@dataclass
class SlabLayout:
    """Pure index logic of one rank's slab (shared by the CUDA path and the CPU tests)."""

    nX: int
    nY: int
    nZ: int
    rank: int
    world: int

    def __post_init__(self):
        self.a, self.b = slab_ranges(self.nX, self.world)[self.rank]
        self.nXloc = self.b - self.a
        if self.nXloc < 1:
            raise ValueError("every rank needs at least one element plane")
        self.planeNodes = (self.nY + 1) * (self.nZ + 1)
        self.planeDofs = 3 * self.planeNodes
        self.nNodeLoc = (self.nXloc + 1) * self.planeNodes
        self.nDofLoc = 3 * self.nNodeLoc
        self.has_lower = self.rank > 0  # receives interface rows from rank-1
        self.has_upper = self.rank < self.world - 1  # sends its ghost-plane rows to rank+1
        # owned dofs: all local dofs except the ghost plane (kept by the last rank, which has none)
        self.ownedDofs = self.nDofLoc - (self.planeDofs if self.has_upper else 0)

    def node_offset(self):
        """global node index = local node index + node_offset()."""
        return self.a * self.planeNodes

    def tail_start(self, indptr):
        """First CSR slot of the ghost plane's rows (the tail sent upwards)."""
        return int(indptr[self.nDofLoc - self.planeDofs])

    def head_nnz(self, indptr):
        """Number of CSR slots of the first node plane's rows (size of the message received from below)."""
        return int(indptr[self.planeDofs])

    def local_mesh(self, lX, lY, lZ, elType="C3D8", x0=0.0, y0=0.0, z0=0.0):
        from .boxgen import box_mesh

        h = lX / self.nX
        return box_mesh(self.nXloc, self.nY, self.nZ, lX=h * self.nXloc, lY=lY, lZ=lZ, x0=x0 + h * self.a, y0=y0, z0=z0, elType=elType)


def exchange_tails(layout: SlabLayout, data, P, F, recv, recvP, recvF, indptr_host, dist, group=None):
    """Send this rank's ghost-plane tail upwards, receive the lower neighbour's tail (blocking P2P batch)."""
    ops = []
    if layout.has_upper:
        ts = layout.tail_start(indptr_host)
        d0 = layout.nDofLoc - layout.planeDofs
        ops += [dist.P2POp(dist.isend, data[ts:], layout.rank + 1, group), dist.P2POp(dist.isend, P[d0:], layout.rank + 1, group),
                dist.P2POp(dist.isend, F[d0:], layout.rank + 1, group)]
    if layout.has_lower:
        ops += [dist.P2POp(dist.irecv, recv, layout.rank - 1, group), dist.P2POp(dist.irecv, recvP, layout.rank - 1, group),
                dist.P2POp(dist.irecv, recvF, layout.rank - 1, group)]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


class _RawDeviceArray:
    """__cuda_array_interface__ view of a raw device allocation (so torch can wrap it without owning it)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None}


class SlabAssembly:
    """One rank's share of a distributed BoxGen assembly: local ElementAssembly + interface exchange.

    exchange: "peer" (ghost-plane rows stored into the neighbour's memory by the sweep kernel), "nccl" (send/recv of the
    finished tail) or None = $EWB_EXCHANGE, default "peer" with a collective fallback to "nccl"."""

    def __init__(self, n, lengths, elType, material, props, rank, world, device, exchange=None, group=None):
        import os

        import torch

        from .assembly import ElementAssembly

        self.layout = SlabLayout(n[0], n[1], n[2], rank, world)
        coords, conn = self.layout.local_mesh(*lengths, elType=elType)
        self.asm = ElementAssembly(elType, conn, coords, material, props, device=device, box=(self.layout.nXloc, n[1], n[2]))
        self.indptr, self.indices = self.asm.csr_pattern()
        self.indptr_host = self.indptr.cpu().numpy()
        lay = self.layout
        f64 = dict(dtype=torch.float64, device=self.asm.device)
        head = lay.head_nnz(self.indptr_host)
        self.recv = torch.zeros(head if lay.has_lower else 1, **f64)
        self.recvP = torch.zeros(lay.planeDofs if lay.has_lower else 1, **f64)
        self.recvF = torch.zeros(lay.planeDofs if lay.has_lower else 1, **f64)
        self.interface_bytes = 8 * (head + 2 * lay.planeDofs) if world > 1 else 0
        self.step = 0
        self._own = self._peer = None
        self.exchange = (exchange or os.environ.get("EWB_EXCHANGE", "peer")) if world > 1 else "none"
        if self.exchange not in ("peer", "nccl", "none"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        if self.exchange == "peer":
            self._setup_peer(head, group)

    # ---- fused transfer: the upper neighbour's receive buffers mapped into this process ----------------------------
    def _setup_peer(self, head, group):
        import warnings

        import torch
        import torch.distributed as dist

        from ._lib import check

        a, lay = self.asm, self.layout
        seg = head + 2 * lay.planeDofs  # doubles per parity: [rows | P | F]
        handle, err = None, None
        try:
            if lay.has_lower:
                ptr, h = C.c_void_p(), C.create_string_buffer(64)
                check(a.lib.ewb_peer_alloc(2 * seg * 8, C.byref(ptr), h))
                self._own, handle = ptr.value, h.raw
        except Exception as e:  # noqa: BLE001 - any failure -> collective fallback below
            err = e
        handles = [None] * lay.world
        dist.all_gather_object(handles, handle, group=group)
        try:
            if err is None and lay.has_upper:
                if handles[lay.rank + 1] is None:
                    raise RuntimeError("upper neighbour has no receive buffer")
                ptr = C.c_void_p()
                check(a.lib.ewb_peer_open(handles[lay.rank + 1], C.byref(ptr)))
                self._peer = ptr.value
                upHead = self._peer_head = int(self.indptr_host[-1]) - lay.tail_start(self.indptr_host)
                self._peer_seg = upHead + 2 * lay.planeDofs
        except Exception as e:  # noqa: BLE001
            err = e
        oks = [None] * lay.world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            if err is not None:
                warnings.warn(f"peer exchange unavailable on rank {lay.rank} ({err}); using the NCCL send/recv exchange")
            self.close()
            self.exchange = "nccl"
            return
        st = C.c_void_p()
        check(a.lib.ewb_plan_status_ptr(a.plan, C.byref(st)))
        self.status = torch.as_tensor(_RawDeviceArray(st.value, 1, "<i4"), device=a.device)
        self._views = []
        if lay.has_lower:
            whole = torch.as_tensor(_RawDeviceArray(self._own, 2 * seg, "<f8"), device=a.device)
            for par in range(2):
                b = whole[par * seg : (par + 1) * seg]
                self._views.append((b[:head], b[head : head + lay.planeDofs], b[head + lay.planeDofs :]))

    def close(self):
        """Unmap / free the peer buffers (safe to call twice)."""
        lib = self.asm.lib
        if self._peer:
            lib.ewb_plan_set_peer(self.asm.plan, None, None, None)
            lib.ewb_peer_close(C.c_void_p(self._peer))
            self._peer = None
        if self._own:
            lib.ewb_peer_free(C.c_void_p(self._own))
            self._own = None

    def assemble(self, flags=0, group=None):
        """Local fused assembly, then the neighbour exchange and the interface add (all on the current stream)."""
        par = self._before_kernel()
        self.asm.assemble(flags)
        self._after_kernel(par, group)

    def _before_kernel(self):
        """"peer" exchange: point the fused kernel at this assembly's half of the upper neighbour's (double buffered) receive buffer."""
        from ._lib import check

        a, lay = self.asm, self.layout
        par = self.step & 1  # receive buffers are double buffered: a fast sender may already be one assembly ahead
        if self.exchange == "peer" and lay.has_upper:
            base = self._peer + 8 * par * self._peer_seg
            check(a.lib.ewb_plan_set_peer(a.plan, C.c_void_p(base), C.c_void_p(base + 8 * self._peer_head),
                                          C.c_void_p(base + 8 * (self._peer_head + lay.planeDofs))))
        return par

    def _after_kernel(self, par, group=None):
        """Neighbour exchange (or, for "peer", the status all-reduce that orders it) and the interface add, on the current stream."""
        import torch.distributed as dist

        from ._lib import check

        a, lay = self.asm, self.layout
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        if self.exchange == "peer":
            # orders the neighbour's peer stores before the interface add; every rank sees the same cut-back request
            # NCCL has no bitwise reductions: MAX gives every rank the SAME word, in which an error bit (4: internal, 8: inverted
            # element) outranks the cut-back bit (1) — errors abort the step on all ranks, a cut-back alone reaches all ranks as 1
            dist.all_reduce(self.status, op=dist.ReduceOp.MAX, group=group)
            if lay.has_lower:
                self.recv, self.recvP, self.recvF = self._views[par]
        elif lay.world > 1:
            exchange_tails(lay, a.csr_data, a.P, a.F, self.recv, self.recvP, self.recvF, self.indptr_host, dist, group)
        if lay.has_lower:
            check(a.lib.ewb_interface_add(p(self.indptr), lay.planeDofs, p(a.csr_data), p(self.recv), p(a.P), p(a.F), p(self.recvP), p(self.recvF),
                                          a._stream()))
        self.step += 1

    def compute_host(self, U, dU, flags=0, group=None):
        """ElementAssembly.compute_host for one rank's slab: pinned host U, dU (local numbering) in, (P, F) out; the matrix (owned
        rows + lower halo block) and the Gauss-point state stay on the device."""
        a = self.asm
        hU, hdU = a._pinned("U", a.nDof), a._pinned("dU", a.nDof)
        if U is not None:  # None: the pinned input buffers (ElementAssembly.host_io) were filled in place
            hU.numpy()[:] = U
        if dU is not None:
            hdU.numpy()[:] = dU
        a.U.copy_(hU, non_blocking=True)
        a.dU.copy_(hdU, non_blocking=True)
        self.assemble(flags, group)
        hP, hF = a._pinned("P", a.nDof), a._pinned("F", a.nDof)
        hP.copy_(a.P, non_blocking=True)
        hF.copy_(a.F, non_blocking=True)
        self.poll(group)
        return hP.numpy(), hF.numpy()

    def compute_host_increment(self, dU, flags=0, group=None):
        """ElementAssembly.compute_host_increment for one rank's slab (after asm.begin_increment): host dU (local numbering) in,
        (P, sum|F| over this rank's owned dofs) out."""
        a = self.asm
        a._increment_upload(dU)
        self.assemble(flags, group)
        hP = a._increment_download(self.layout.ownedDofs)  # the ghost plane's entries are the upper neighbour's
        self.poll(group)
        return hP.numpy(), float(a._pin_fsum[0])

    def compute_host_increment_pipelined(self, dU, flags=0, group=None):
        """compute_host_increment with the host transfers overlapped with the kernel x-chunk by x-chunk
        (ElementAssembly._increment_pipelined); the exchange and the interface add follow the last chunk, and the first node
        plane of P — the only part the interface add changes — is downloaded again after it.  Same results, bitwise."""
        import torch

        a, lay = self.asm, self.layout
        bounds = a.x_chunks(flags)
        if bounds is None or len(bounds) < 3 or getattr(a, "Un", None) is None:
            return self.compute_host_increment(dU, flags, group)
        par = self._before_kernel()
        hP = a._increment_pipelined(dU, bounds, flags)
        self._after_kernel(par, group)
        cur = torch.cuda.current_stream(a.device)
        cur.wait_stream(a._pipe_streams["d2h"])
        if lay.has_lower:
            hP[: lay.planeDofs].copy_(a.P[: lay.planeDofs], non_blocking=True)
        torch.sum(a.F[: lay.ownedDofs], dim=0, keepdim=True, out=a._fsum)
        a._pin_fsum.copy_(a._fsum, non_blocking=True)
        self.poll(group)
        return hP.numpy(), float(a._pin_fsum[0])

    def poll(self, group=None):
        """Synchronise; every rank takes the same cut-back decision (nonlinearimplicitstatic.py:253-262): in the "peer" exchange the
        status words were combined by the all-reduce of assemble(); the "nccl" exchange reduces a flag here."""
        import torch
        import torch.distributed as dist

        from .assembly import CutbackRequest

        if self.exchange != "nccl" or self.layout.world == 1:
            return self.asm.poll()
        failed = 0
        try:
            self.asm.poll()
        except CutbackRequest:
            failed = 1
        flag = torch.tensor([failed], dtype=torch.int32, device=self.asm.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()):
            raise CutbackRequest("Von Mises Newton failed.", 0.5)

    # owned part of the distributed system (rows of owned dofs, local column numbering + lower halo block in self.recv)
    def owned_slices(self):
        lay = self.layout
        nnz_owned = int(self.indptr_host[lay.ownedDofs])
        return slice(0, lay.ownedDofs), slice(0, nnz_owned)


def interface_add_host(indptr, n_rows, data, recv, P, F, rP, rF):
    """NumPy mirror of ewb_interface_add (host-side index logic; used by the CPU gloo tests)."""
    for row in range(n_rows):
        r0, r1 = int(indptr[row]), int(indptr[row + 1])
        half = (r1 - r0) // 2
        data[r0 : r0 + half] += recv[r0 + half : r1]
    P[:n_rows] += rP
    F[:n_rows] += rF


def slab_parity_check(world, rank, device, n=None, elType="C3D8", material="vonmises", props=(2.1e4, 0.22, 355.0, 1000.0, 200.0, 1400.0), group=None):
    """Slab-partitioned assembly of a small box on `world` ranks against the single-GPU assembly of the same box (computed on
    rank 0): owned CSR rows, the lower halo block, P and F.  Returns the largest relative error on rank 0 (None elsewhere).
    Collective: every rank of the group must call it.  Used by bench.py (--gpus N > 1) and tests/test_gpu_multi.py."""
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist

    from .assembly import ElementAssembly
    from .boxgen import box_mesh

    n = n or (3 * world + 1, 6, 5)
    lengths = (float(n[0]), float(n[1]), float(n[2]))
    coords, conn = box_mesh(*n, lX=lengths[0], lY=lengths[1], lZ=lengths[2], elType=elType)
    rng = np.random.default_rng(3)
    coords = coords + 0.1 * rng.uniform(-1, 1, coords.shape)
    dU = 4e-3 * rng.standard_normal(3 * coords.shape[0])
    slab = SlabAssembly(n, lengths, elType, material, list(props), rank, world, device, group=group)
    lay, asm = slab.layout, slab.asm
    n0 = lay.node_offset()
    asm.coords.copy_(torch.as_tensor(coords[n0 : n0 + lay.nNodeLoc]))
    ldU = dU[3 * n0 : 3 * (n0 + lay.nNodeLoc)]
    asm.U.copy_(torch.as_tensor(ldU))
    asm.dU.copy_(torch.as_tensor(ldU))
    for _ in range(3):  # repeated assemblies: the double-buffered receive side must stay consistent
        slab.assemble(group=group)
    slab.poll(group)
    # the host-facing calls of the same slab: full (U, dU in / P, F out), lean (dU in / P, sum|F| out) and lean with the transfers
    # pipelined over the kernel's x-chunks must agree bit for bit with the device-resident assembly above
    host_ok = 1
    Pd, Fd = asm.P.cpu().numpy().copy(), asm.F.cpu().numpy().copy()
    _, own_nnz = slab.owned_slices()
    Kd = asm.csr_data[own_nnz].clone()
    Pf, Ff = slab.compute_host(ldU, ldU, group=group)
    host_ok &= int(np.array_equal(Pf[: lay.ownedDofs], Pd[: lay.ownedDofs]) and np.array_equal(Ff[: lay.ownedDofs], Fd[: lay.ownedDofs]))
    asm.begin_increment(np.zeros_like(ldU))
    Pl, fl = slab.compute_host_increment(ldU, group=group)
    host_ok &= int(np.array_equal(Pl[: lay.ownedDofs], Pd[: lay.ownedDofs]))
    Pl = Pl.copy()
    asm.csr_data.fill_(float("nan"))
    Pp, fp = slab.compute_host_increment_pipelined(ldU, group=group)
    host_ok &= int(np.array_equal(Pp[: lay.ownedDofs], Pl[: lay.ownedDofs]) and fp == fl and bool(torch.equal(asm.csr_data[own_nnz], Kd)))
    host_ok &= int(abs(fl - Fd[: lay.ownedDofs].sum()) <= 1e-12 * abs(fl))
    chunks = asm.x_chunks()
    rows, nnzs = slab.owned_slices()
    mine = (rank, 3 * n0, slab.indptr_host[: lay.ownedDofs + 1].copy(), slab.indices.cpu().numpy()[nnzs], asm.csr_data.cpu().numpy()[nnzs],
            asm.P.cpu().numpy()[rows], asm.F.cpu().numpy()[rows], slab.recv.cpu().numpy(), lay.planeDofs, lay.has_lower,
            (slab.exchange, host_ok, 0 if chunks is None else len(chunks) - 1))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    slab.close()
    if rank != 0:
        return None
    ref = ElementAssembly(elType, conn, coords, material, list(props), device=device, box=n)
    ref.U.copy_(torch.as_tensor(dU))
    ref.dU.copy_(torch.as_tensor(dU))
    ref.assemble()
    ref.poll()
    Kg = ref.to_scipy()
    Pg, Fg = ref.P.cpu().numpy(), ref.F.cpu().numpy()
    nG = Kg.shape[0]
    scale = abs(Kg).max()
    worst = 0.0
    for _r, off, indptr, indices, data, P, F, recv, planeDofs, has_lower, (_ex, host_ok, _nch) in gathered:
        if not host_ok:
            worst = max(worst, 1.0)  # a host-facing call disagreed with the device-resident assembly
        nrows = indptr.size - 1
        Kl = sp.csr_matrix((data, indices.astype(np.int64) + off, indptr), shape=(nrows, nG))
        Kref = Kg[off : off + nrows]
        if has_lower:  # the dx=-1 columns live in the halo block: compare them with the received rows
            lowcols = np.arange(off - planeDofs, off)
            halo = Kref[:planeDofs][:, lowcols]
            Kref = Kref.tolil()
            Kref[:planeDofs, lowcols] = 0
            Kref = Kref.tocsr()
            got = [recv[indptr[r] : indptr[r] + (indptr[r + 1] - indptr[r]) // 2] for r in range(planeDofs)]
            halo.sort_indices()
            worst = max(worst, float(np.abs(np.concatenate(got) - halo.data).max() / scale))
        worst = max(worst, float(abs(Kl - Kref).max() / scale))
        worst = max(worst, float(np.abs(P - Pg[off : off + nrows]).max() / np.abs(Pg).max()))
        worst = max(worst, float(np.abs(F - Fg[off : off + nrows]).max() / np.abs(Fg).max()))
    return worst
