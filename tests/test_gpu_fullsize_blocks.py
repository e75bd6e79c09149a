"""Oracle parity on SAMPLED SUB-BLOCKS of the full-size BASELINE meshes (configs 2-5; SURVEY §0 Gotcha 2: the reference and the
oracle cannot hold these meshes, so the check is made where it is affordable — blocks of 4x4x4 elements at the box corners, on
faces, across tile / x-chunk boundaries and in the interior).  For every block the oracle assembles the block's own sub-mesh with
the same U, dU and state; compared ENTRY-WISE (relative with an absolute floor, next to the norm-wise 1e-12): the CSR rows of the
nodes whose elements all lie inside the block, their P and F, and the updated state of every element of the block."""
import numpy as np
import pytest
from conftest import entrywise, relerr

pytestmark = pytest.mark.gpu


def _blocks(n, bs):
    nx, ny, nz = n
    cand = [(0, 0, 0), (nx - bs, ny - bs, nz - bs), (0, ny - bs, nz // 2), (nx // 2, 0, 0),  # corners / edges / faces
            (nx // 2 - 1, ny // 2 - 2, nz // 2 - 3),                                       # interior
            (15, 5, 5), (33, 11, 6), (nx - bs, 3, 12),                                      # across x-chunk ends and y / z tile edges
            (49, 6, nz - bs), (66, ny - bs, 20), (2, 13, 27)]
    out = []
    for b in cand:
        b = tuple(int(min(max(v, 0), m - bs)) for v, m in zip(b, n))
        if b not in out:
            out.append(b)
    return out


@pytest.mark.parametrize("workload", ["le100", "vm_200x100x100", "nh100", "c3d20_100x100x50"])
def test_sampled_blocks_against_oracle(workload):
    import torch

    from edelweissfe_b200 import ElementAssembly, box_mesh
    from oracle import port

    elType, bs = "C3D8", 4
    if workload == "le100":
        n, material, props, scale = (100, 100, 100), "linearelastic", [2.1e4, 0.22], 1e-3
    elif workload == "vm_200x100x100":
        n, material, props, scale = (200, 100, 100), "vonmises", [2.1e4, 0.22, 355.0, 1000.0, 200.0, 1400.0], 1e-6
    elif workload == "nh100":
        elType, n, material, props, scale = "C3D8TL", (100, 100, 100), "neohookewa", [91304.34783, 100000.0], 1e-2
    else:
        elType, n, material, props, scale, bs = "C3D20", (100, 100, 50), "linearelastic", [2.1e4, 0.22], 1e-3, 2
    nn = 20 if elType == "C3D20" else 8
    nGp = 27 if nn == 20 else 8
    coords, conn = box_mesh(*n, lX=float(n[0]), lY=float(n[1]), lZ=float(n[2]), elType=elType)
    asm = ElementAssembly(elType, conn, coords, material, props, box=n if nn == 8 else None)
    rng = np.random.default_rng(0)
    dU = scale * rng.standard_normal(asm.nDof)
    if material == "vonmises":  # the bench recipe: shear ramp, about half of the Gauss points yield
        G = 2.1e4 / (2 * 1.22)
        dU[0::3] += 0.5 * (2.0 * 355.0 / (np.sqrt(3.0) * G)) * coords[:, 1] ** 2 / n[1]
    asm.U.copy_(torch.as_tensor(dU))
    asm.dU.copy_(torch.as_tensor(dU))
    asm.assemble()
    asm.poll()
    indptr, indices = asm.csr_pattern()
    indptr_h = indptr.cpu().numpy().astype(np.int64)
    P, F = asm.P.cpu().numpy(), asm.F.cpu().numpy()
    worst = dict(K=0.0, P=0.0, F=0.0, state=0.0, Kn=0.0)
    nBlocks = 0
    for bx, by, bz in _blocks(n, bs):
        ex, ey, ez = np.meshgrid(np.arange(bx, bx + bs), np.arange(by, by + bs), np.arange(bz, bz + bs), indexing="ij")
        elems = ((ex * n[1] + ey) * n[2] + ez).reshape(-1)
        sub = conn[elems]
        nodes = np.unique(sub)
        loc = {g: i for i, g in enumerate(nodes)}
        subconn = np.vectorize(loc.get)(sub).astype(np.int32)
        gd = (3 * nodes[:, None] + np.arange(3)[None, :]).reshape(-1)
        state = np.zeros((elems.size, nGp, asm.nState))
        o = port.assemble(elType, material, props, coords[nodes], subconn, dU[gd], dU[gd], state, want_vij=False)
        # nodes whose incident elements all lie inside the block (count their incidences in the whole mesh: box faces included)
        inc_blk = np.bincount(subconn.reshape(-1), minlength=nodes.size)
        deg = (indptr_h[3 * nodes + 1] - indptr_h[3 * nodes]) // 3  # neighbours incl. itself: fixes the expected incidence
        ldeg = (o["indptr"][3 * np.arange(nodes.size) + 1] - o["indptr"][3 * np.arange(nodes.size)]) // 3
        complete = np.where(deg == ldeg)[0] if nn == 8 else np.where((deg == ldeg) & (inc_blk > 0))[0]
        assert complete.size >= (bs - 1) ** 3
        for ln in complete:
            g = int(nodes[ln])
            for i in range(3):
                a0, a1 = int(indptr_h[3 * g + i]), int(indptr_h[3 * g + i + 1])
                cols = indices[a0:a1].cpu().numpy().astype(np.int64)
                vals = asm.csr_data[a0:a1].cpu().numpy()
                b0, b1 = int(o["indptr"][3 * ln + i]), int(o["indptr"][3 * ln + i + 1])
                lcols = gd[o["indices"][b0:b1]]
                assert np.array_equal(cols, lcols)  # pattern of the row, bit-exact
                worst["K"] = max(worst["K"], entrywise(vals, o["data"][b0:b1]))
                worst["Kn"] = max(worst["Kn"], relerr(vals, o["data"][b0:b1]))
        cd = (3 * complete[:, None] + np.arange(3)[None, :]).reshape(-1)
        worst["P"] = max(worst["P"], entrywise(P[gd[cd]], o["P"][cd], rel=1e-8, floor=1e-12))
        worst["F"] = max(worst["F"], entrywise(F[gd[cd]], o["F"][cd]))
        st = asm.state_temp[:, torch.as_tensor(elems, device=asm.device), :].permute(1, 2, 0).cpu().numpy()
        worst["state"] = max(worst["state"], entrywise(st, o["stateTemp"], rel=1e-9, floor=1e-13))
        nBlocks += 1
    assert nBlocks >= 8
    assert worst["Kn"] < 1e-12, worst
    assert max(worst["K"], worst["P"], worst["F"], worst["state"]) <= 1.0, worst
