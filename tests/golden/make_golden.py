"""Generate tests/golden/*.npz from the UNMODIFIED reference (imported from /root/reference
via tools/refshim.py).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds the inputs (coords, conn, props, U, dU, stateRef) and what the reference's
own objects produced for them: DofManager element dofs, VIJ I/J, CSRGenerator pattern
(indptr/indices, int32), CSR data, P, F, per-element stateTemp (and the VIJ values V).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402  (only for the BoxGen coordinates used as *input*)
from tools.refdriver import RefModel  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, elType, material, props, box, distort, scaleU per pass
    ("c3d8_le_box", "C3D8", "linearelastic", [2.1e4, 0.22], dict(nX=3, nY=4, nZ=2, lX=3.0, lY=4.5, lZ=2.2), 0.0, [1e-3]),
    ("c3d8_le_distorted", "C3D8", "linearelastic", [2.1e4, 0.22], dict(nX=3, nY=4, nZ=2, lX=3.0, lY=4.5, lZ=2.2), 0.2, [1e-3, 1e-3]),
    ("c3d8_vm_distorted", "C3D8", "vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400], dict(nX=3, nY=4, nZ=2, lX=3.0, lY=4.5, lZ=2.2), 0.2, [6e-3, 4e-3, 0.0]),
    ("c3d20_le_box", "C3D20", "linearelastic", [2.1e4, 0.22], dict(nX=2, nY=3, nZ=2, lX=2.0, lY=3.3, lZ=2.2), 0.0, [1e-3, 1e-3]),
    ("c3d20_le_affine", "C3D20", "linearelastic", [2.1e4, 0.22], dict(nX=2, nY=2, nZ=1, lX=2.0, lY=2.4, lZ=1.1), "affine", [1e-3]),
    ("c3d8tl_nha_distorted", "C3D8TL", "neohookewa", [91304.34783, 100000.0], dict(nX=3, nY=4, nZ=2, lX=3.0, lY=4.5, lZ=2.2), 0.2, [3e-2, 2e-2]),
    ("c3d8tl_nhb_distorted", "C3D8TL", "neohookewb", [91304.34783, 100000.0], dict(nX=3, nY=4, nZ=2, lX=3.0, lY=4.5, lZ=2.2), 0.2, [3e-2, 2e-2]),
    # integration variants of the same formulation (library.py:228-259, 276-291)
    ("c3d8r_le_distorted", "C3D8R", "linearelastic", [2.1e4, 0.22], dict(nX=3, nY=3, nZ=2, lX=3.0, lY=3.3, lZ=2.2), 0.2, [1e-3, 1e-3]),
    ("c3d8r_vm_distorted", "C3D8R", "vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400], dict(nX=3, nY=3, nZ=2, lX=3.0, lY=3.3, lZ=2.2), 0.2, [1.5e-2, 1e-2]),
    ("c3d8e_vm_distorted", "C3D8E", "vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400], dict(nX=2, nY=3, nZ=2, lX=2.0, lY=3.3, lZ=2.2), 0.2, [6e-3, 4e-3]),
    ("c3d20r_le_box", "C3D20R", "linearelastic", [2.1e4, 0.22], dict(nX=2, nY=2, nZ=2, lX=2.0, lY=2.2, lZ=2.4), 0.0, [1e-3, 1e-3]),
    # total-Lagrange element with hypo-elastic materials: B^T C B + geometric stiffness (displacementtlelement/element.py:415-425)
    ("c3d8tl_le_distorted", "C3D8TL", "linearelastic", [2.1e4, 0.22], dict(nX=2, nY=3, nZ=2, lX=2.0, lY=3.3, lZ=2.2), 0.15, [2e-2, 1e-2, 1e-2]),
    ("c3d8tl_vm_distorted", "C3D8TL", "vonmises", [2.1e4, 0.22, 355, 1000, 200, 1400], dict(nX=2, nY=3, nZ=2, lX=2.0, lY=3.3, lZ=2.2), 0.15, [1.2e-2, 8e-3, 5e-3]),
    ("c3d8tl_nhc_distorted", "C3D8TL", "neohookewc", [91304.34783, 100000.0], dict(nX=3, nY=4, nZ=2, lX=3.0, lY=4.5, lZ=2.2), 0.2, [3e-2, 2e-2]),
]


def make(name, elType, material, props, box, distort, scales, seed=0):
    nn = 20 if "20" in elType else 8
    coords, conn = port.boxgen(box["nX"], box["nY"], box["nZ"], box["lX"], box["lY"], box["lZ"], nnodes=nn)
    rng = np.random.default_rng(seed)
    if distort == "affine":  # parallelepiped elements: shear + stretch the whole box
        A = np.array([[1.0, 0.2, 0.1], [0.05, 0.9, 0.15], [0.1, -0.1, 1.1]])
        coords = coords @ A.T
        m = RefModel(elType, material, props, nodes=coords, conn=conn)
    elif distort:
        h = min(box["lX"] / box["nX"], box["lY"] / box["nY"], box["lZ"] / box["nZ"])
        coords = coords + distort * h * rng.uniform(-1, 1, coords.shape)
        m = RefModel(elType, material, props, nodes=coords, conn=conn)
    else:
        m = RefModel(elType, material, props, box=box)
        assert np.array_equal(m.coords(), coords)
    assert np.array_equal(m.connectivity(), conn)
    n = 3 * coords.shape[0]
    out = dict(
        elType=elType, material=material, props=np.asarray(props, float), coords=coords, conn=conn,
        box=np.array([box[k] for k in ("nX", "nY", "nZ", "lX", "lY", "lZ")], float), boxgen_regular=not bool(distort),
        element_dofs=m.element_dofs(), nPasses=len(scales),
    )
    U = np.zeros(n)
    for p, sc in enumerate(scales):
        dU = sc * rng.standard_normal(n)
        U = U + dU
        stateRef = np.array([np.array(el._stateVarsRef) for el in m.elements])
        r = m.assemble(U, dU)
        if p == 0:
            out.update(I=r["I"].astype(np.int64), J=r["J"].astype(np.int64), indptr=r["indptr"], indices=r["indices"])
        out.update({f"U{p}": U.copy(), f"dU{p}": dU, f"stateRef{p}": stateRef, f"data{p}": r["data"], f"P{p}": r["P"],
                    f"F{p}": r["F"], f"stateTemp{p}": r["stateTemp"]})
        if nn == 8:
            out[f"V{p}"] = r["V"]
        if material == "vonmises":
            k0, k1 = stateRef[..., 12], r["stateTemp"][..., 12]
            print(f"   {name} pass {p}: plastic GP fraction {(k1 > k0).mean():.3f}")
        m.accept()
    # body force: the reference's own computeBodyForce on every element (element.py:348-371)
    load = np.array([0.3, -0.000077, 1.7])
    PExt = np.zeros(n)
    for el, d in zip(m.elements, m.element_dofs()):
        Pe = np.zeros(el.nDof)
        el.computeBodyForce(Pe, np.zeros(el.nDof**2), load, np.zeros(el.nDof), np.array([0.0, 0.0]), 1.0)
        PExt[d] += Pe
    out.update(bodyforce_load=load, bodyforce_PExt=PExt)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for c in CASES:
        make(*c)
