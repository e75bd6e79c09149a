"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

NumPy restatement of EdelweissFE's element-loop hot path (NIST.computeElements +
CSRGenerator.updateCSR) for provider `edelweiss` C3D8 / C3D20 / C3D8TL with
LinearElastic / VonMises / NeoHooke-W{a,b,c}.  It follows the reference's *dense*
formulation (explicit 6 x nDof B operator, Bt C B products, the 3^4 dtau/dF tensor for TL),
batched over elements, so that it is an independent check of the structured CUDA kernels.

Parity status: PINNED.  tests/golden/*.npz hold outputs of the unmodified reference
(generated in the build container by tests/golden/make_golden.py, which imports
/root/reference) and tests/test_oracle_golden.py checks this file against them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product path (edelweissfe_b200/) never does.

All `ref:` citations are paths below /root/reference/edelweissfe/.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# element tables
# --------------------------------------------------------------------------------------

# ref: generators/boxgen.py:172-185 (Hexa8 connectivity offsets dx,dy,dz)
HEXA8_OFFSETS = np.array(
    [(0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0), (0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)], dtype=np.int64
)
# ref: generators/boxgen.py:187-299 (Hexa20 offsets on the doubled grid)
HEXA20_OFFSETS = np.array(
    [(0, 0, 0), (0, 0, 2), (2, 0, 2), (2, 0, 0), (0, 2, 0), (0, 2, 2), (2, 2, 2), (2, 2, 0),
     (0, 0, 1), (1, 0, 2), (2, 0, 1), (1, 0, 0), (0, 2, 1), (1, 2, 2), (2, 2, 1), (1, 2, 0),
     (0, 1, 0), (0, 1, 2), (2, 1, 2), (2, 1, 0)], dtype=np.int64,
)


def _local_coords(offsets, span):
    """Local (xi, eta, zeta) of every node in the reference's naming: the library's `eta`
    runs along BoxGen x, `xi` along y, `zeta` along z (ref: elements/displacementtlelement/
    _elementcomputationmatrices.py:389-427)."""
    o = offsets.astype(float) * (2.0 / span) - 1.0
    return np.stack([o[:, 1], o[:, 0], o[:, 2]], axis=1)  # columns: xi, eta, zeta


def gauss_points(eltype: str):
    """(xi, eta, zeta, w) per Gauss point.  ref: elements/library.py:34-47, 212-227, 260-275."""
    t = eltype.upper()
    if t in ("C3D8", "C3D8TL", "C3D8N", "C3D8NTL", "C3D20R"):
        g = 1.0 / np.sqrt(3.0)
        s8 = g * np.array([-1, 1, 1, -1.0])
        t8 = g * np.array([-1, -1, 1, 1.0])
        xi = g * np.hstack([-np.ones(4), np.ones(4)])
        return xi, np.hstack([t8, t8]), np.hstack([s8, s8]), np.ones(8)
    if t == "C3D8R":  # reduced integration: one point, weight 8 (library.py:228-243)
        z = np.zeros(1)
        return z, z.copy(), z.copy(), np.array([8.0])
    if t in ("C3D20", "C3D8E", "C3D20TL"):
        r = np.sqrt(0.6)
        s20 = np.array([-1, 0, 1.0])
        t20 = np.array([-1, -1, -1, 0, 0, 0, 1, 1, 1.0])
        xi = r * np.hstack([-np.ones(9), np.zeros(9), np.ones(9)])
        eta = r * np.hstack([t20, t20, t20])
        zeta = r * np.hstack([s20 for _ in range(9)])
        w1 = {-1.0: 5.0 / 9.0, 0.0: 8.0 / 9.0, 1.0: 5.0 / 9.0}
        w = np.array([w1[round(a / r)] * w1[round(b / r)] * w1[round(c / r)] for a, b, c in zip(xi, eta, zeta)])
        return xi, eta, zeta, w
    raise ValueError(eltype)


def shape_derivatives(nnodes: int, xi, eta, zeta):
    """dN[gp, r, a]; rows r ordered (d/d eta, d/d xi, d/d zeta) like the reference tables
    (ref: displacementtlelement/_elementcomputationmatrices.py:370-498).  Written from the
    standard trilinear / serendipity formulas, not from the reference's expanded polynomials."""
    xi, eta, zeta = (np.atleast_1d(np.asarray(v, dtype=float)) for v in (xi, eta, zeta))
    if nnodes == 8:
        lc = _local_coords(HEXA8_OFFSETS, 1)
    else:
        lc = _local_coords(HEXA20_OFFSETS, 2)
    a, b, c = lc[:, 0][None, :], lc[:, 1][None, :], lc[:, 2][None, :]  # node xi, eta, zeta
    X, E, Z = xi[:, None], eta[:, None], zeta[:, None]
    if nnodes == 8:
        dxi = a * (1 + b * E) * (1 + c * Z) / 8
        deta = b * (1 + a * X) * (1 + c * Z) / 8
        dzeta = c * (1 + a * X) * (1 + b * E) / 8
    else:
        fx, fe, fz = 1 + a * X, 1 + b * E, 1 + c * Z
        s = a * X + b * E + c * Z - 2
        # corners
        dxi_c = a * fe * fz * (s + fx) / 8
        deta_c = b * fx * fz * (s + fe) / 8
        dzeta_c = c * fx * fe * (s + fz) / 8
        # mid-edge nodes with one zero local coordinate
        dxi = np.where(a == 0, -2 * X * fe * fz / 4, np.where(b == 0, a * (1 - E**2) * fz / 4, np.where(c == 0, a * fe * (1 - Z**2) / 4, dxi_c)))
        deta = np.where(a == 0, (1 - X**2) * b * fz / 4, np.where(b == 0, -2 * E * fx * fz / 4, np.where(c == 0, fx * b * (1 - Z**2) / 4, deta_c)))
        dzeta = np.where(a == 0, (1 - X**2) * fe * c / 4, np.where(b == 0, fx * (1 - E**2) * c / 4, np.where(c == 0, -2 * Z * fx * fe / 4, dzeta_c)))
    return np.stack([deta, dxi, dzeta], axis=1)


ELEMENT_INFO = {
    "C3D8": dict(nnodes=8, tl=False),
    "C3D20": dict(nnodes=20, tl=False),
    "C3D8TL": dict(nnodes=8, tl=True),
    # integration variants (library.py:228-259, 276-291): same formulation, other Gauss rule
    "C3D8R": dict(nnodes=8, tl=False),
    "C3D8E": dict(nnodes=8, tl=False),
    "C3D20R": dict(nnodes=20, tl=False),
}

# number of material state variables; ref: materials/linearelastic/linearelastic.py:49-57,
# materials/vonmises/vonmises.py:95-103, materials/neohooke/neohookepencegouformulationa.py:55-63
MATERIAL_NSTATE = {"linearelastic": 0, "vonmises": 1, "neohookewa": 1, "neohookewb": 1, "neohookewc": 1}


# --------------------------------------------------------------------------------------
# mesh / numbering / patterns
# --------------------------------------------------------------------------------------

def boxgen(nX, nY, nZ, lX=1.0, lY=1.0, lZ=1.0, x0=0.0, y0=0.0, z0=0.0, nnodes=8):
    """Node coordinates [nNode,3] and 0-based connectivity [nEl,nnodes] of a BoxGen mesh.
    ref: generators/boxgen.py:116-141 (nodes, z fastest; Hexa20 keeps grid nodes with
    (ix%2)+(iy%2)+(iz%2) < 2), :168-299 (elements, ix-major)."""
    m = 1 if nnodes == 8 else 2
    NX, NY, NZ = m * nX + 1, m * nY + 1, m * nZ + 1
    xs, ys, zs = np.linspace(x0, x0 + lX, NX), np.linspace(y0, y0 + lY, NY), np.linspace(z0, z0 + lZ, NZ)
    ix, iy, iz = np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij")
    keep = np.ones(ix.shape, dtype=bool) if nnodes == 8 else ((ix % 2) + (iy % 2) + (iz % 2) < 2)
    gridToNode = np.full(NX * NY * NZ, -1, dtype=np.int64)
    gridToNode[keep.ravel()] = np.arange(int(keep.sum()))
    coords = np.stack([xs[ix.ravel()], ys[iy.ravel()], zs[iz.ravel()]], axis=1)[keep.ravel()]
    ex, ey, ez = np.meshgrid(np.arange(nX), np.arange(nY), np.arange(nZ), indexing="ij")
    off = HEXA8_OFFSETS if nnodes == 8 else HEXA20_OFFSETS
    gx = m * ex.ravel()[:, None] + off[None, :, 0]
    gy = m * ey.ravel()[:, None] + off[None, :, 1]
    gz = m * ez.ravel()[:, None] + off[None, :, 2]
    conn = gridToNode[gx * (NY * NZ) + gy * NZ + gz]
    assert (conn >= 0).all()
    return coords, conn.astype(np.int32)


def element_dofs(conn):
    """dof(node i, component c) = 3 i + c, node-major per element.
    ref: numerics/dofmanager.py:280-292, 445-471."""
    conn = np.asarray(conn, dtype=np.int64)
    return (3 * conn[:, :, None] + np.arange(3)[None, None, :]).reshape(conn.shape[0], -1)


def vij_pattern(dofs):
    """COO pattern: element e owns [e n^2, (e+1) n^2); I[off+p] = dof[p % n], J[off+p] = dof[p // n].
    ref: numerics/dofmanager.py:543-555."""
    nEl, n = dofs.shape
    I = np.tile(dofs, (1, n)).reshape(-1)  # noqa: E741
    J = np.repeat(dofs, n, axis=1).reshape(-1)
    return I.astype(np.int64), J.astype(np.int64)


def csr_pattern(I, J, nDof):  # noqa: E741
    """SciPy-canonical CSR pattern of the union of COO pairs (rows ascending, columns sorted,
    duplicates merged, explicit zeros kept), int32 indptr/indices, and the COO->CSR slot map x.
    ref: numerics/csrgenerator.pyx:47-98."""
    key = I.astype(np.int64) * np.int64(nDof) + J.astype(np.int64)
    uniq, inv = np.unique(key, return_inverse=True)
    rows = uniq // nDof
    indices = (uniq % nDof).astype(np.int32)
    indptr = np.zeros(nDof + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    return indptr, indices, inv.astype(np.int32)


def update_csr(x, V, nnz):
    """data[x[p]] += V[p] in ascending p.  ref: numerics/csrgenerator.pyx:100-115.
    (np.bincount accumulates sequentially in input order, i.e. the same summation order.)"""
    return np.bincount(x, weights=V, minlength=nnz)


# --------------------------------------------------------------------------------------
# kinematics
# --------------------------------------------------------------------------------------

def jacobians(dN, X):
    """J[e,gp] = dN[gp] . X_e  (3x3; rows = local (eta,xi,zeta), cols = x,y,z).
    ref: displacementelement/_elementcomputationmatrices.py:306-366 (_J3D8 polynomial form,
    identical); for C3D20 see DESIGN.md (reference quirk on non-affine elements)."""
    return np.einsum("gra,eac->egrc", dN, X)


def nabla_n(dN, J):
    """nablaN[e,gp,:,a] = inv(J) . dN[:,a].  ref: displacementtlelement/_elementcomputationmatrices.py:247-281."""
    invJ = np.linalg.inv(J)
    return np.einsum("egcr,gra->egca", invJ, dN)


def b_operator(gradN):
    """Small-strain B[e,gp,6,3n]; Voigt rows 11,22,33,12,13,23, engineering shear.
    ref: displacementelement/_elementcomputationmatrices.py:700-817 (_B3D8), :820-1060 (_B3D20)."""
    nEl, nGp, _, n = gradN.shape
    B = np.zeros((nEl, nGp, 6, 3 * n))
    Nx, Ny, Nz = gradN[:, :, 0, :], gradN[:, :, 1, :], gradN[:, :, 2, :]
    B[:, :, 0, 0::3] = Nx
    B[:, :, 1, 1::3] = Ny
    B[:, :, 2, 2::3] = Nz
    B[:, :, 3, 0::3] = Ny
    B[:, :, 3, 1::3] = Nx
    B[:, :, 4, 0::3] = Nz
    B[:, :, 4, 2::3] = Nx
    B[:, :, 5, 1::3] = Nz
    B[:, :, 5, 2::3] = Ny
    return B


# --------------------------------------------------------------------------------------
# materials
# --------------------------------------------------------------------------------------

def elasticity_matrix(E, v):
    """ref: materials/linearelastic/linearelastic.py:95-117 (same in vonmises.py)."""
    return (
        E / ((1 + v) * (1 - 2 * v))
        * np.array(
            [
                [(1 - v), v, v, 0, 0, 0],
                [v, (1 - v), v, 0, 0, 0],
                [v, v, (1 - v), 0, 0, 0],
                [0, 0, 0, (1 - 2 * v) / 2, 0, 0],
                [0, 0, 0, 0, (1 - 2 * v) / 2, 0],
                [0, 0, 0, 0, 0, (1 - 2 * v) / 2],
            ]
        )
    )


def linear_elastic(props, stress, dstrain):
    """stress += C dstrain; tangent = C.  ref: materials/linearelastic/linearelastic.py:185-210."""
    Ei = elasticity_matrix(props[0], props[1])
    C = np.broadcast_to(Ei, stress.shape[:-1] + (6, 6)).copy()
    return stress + dstrain @ Ei.T, C, np.zeros(stress.shape[:-1] + (0,)), np.zeros(stress.shape[:-1], dtype=bool)


_IDEV = np.array(
    [[2 / 3, -1 / 3, -1 / 3, 0, 0, 0], [-1 / 3, 2 / 3, -1 / 3, 0, 0, 0], [-1 / 3, -1 / 3, 2 / 3, 0, 0, 0],
     [0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1.0]]
)
_IDEV_HALF = _IDEV.copy()
_IDEV_HALF[3:, 3:] *= 0.5


def von_mises(props, stress, dstrain, kappa):
    """J2 radial return with scalar Newton on d-kappa (start 0, |R|>1e-12, at most 15 updates)
    and the consistent tangent.  ref: materials/vonmises/vonmises.py:35-64, 105-124, 186-254.
    Returns (stress, C, kappa_new[...,1], failed) where failed marks Gauss points at which the
    reference would raise CutbackRequest("Von Mises Newton failed.", 0.5)."""
    E, v, fy0, HLin, dfy, delta = (float(p) for p in props[:6])
    G = E / (2 * (1.0 + v))
    Ei = elasticity_matrix(E, v)
    shp = stress.shape[:-1]
    s_in = stress.reshape(-1, 6)
    de = dstrain.reshape(-1, 6)
    k0 = kappa.reshape(-1)
    n_pts = s_in.shape[0]
    fy = lambda k: fy0 + HLin * k + dfy * (1.0 - np.exp(-delta * k))  # noqa: E731
    dfy_dk = lambda k: HLin + dfy * delta * np.exp(-delta * k)  # noqa: E731

    out_s = s_in.copy()
    out_C = np.broadcast_to(Ei, (n_pts, 6, 6)).copy()
    out_k = k0.copy()
    failed = np.zeros(n_pts, dtype=bool)

    nonzero = np.sqrt(np.sum(de * de, axis=1)) >= 10**-14  # :211
    trial = s_in + de @ Ei  # :215
    dev = trial @ _IDEV.T
    devn = np.sqrt(np.sum(np.square(dev[:, 0:3]), axis=1) + 2 * np.sum(np.square(dev[:, 3:6]), axis=1))
    f = devn - np.sqrt(2 / 3) * fy(k0)
    plastic = nonzero & (f > 0.0)
    elastic = nonzero & ~plastic
    out_s[elastic] = trial[elastic]
    if plastic.any():
        idx = np.nonzero(plastic)[0]
        dn, kk = devn[idx], k0[idx]
        R = lambda dk: dn - np.sqrt(6) * G * dk - np.sqrt(2 / 3) * fy(kk + dk)  # noqa: E731
        dk = np.zeros(idx.size)
        active = np.abs(R(dk)) > 1e-12
        counter = 0
        fail_local = np.zeros(idx.size, dtype=bool)
        while active.any():
            if counter == 15:
                fail_local |= active
                break
            dR = -np.sqrt(6) * G - np.sqrt(2 / 3) * dfy_dk(kk + dk)
            dk = np.where(active, dk - R(dk) / dR, dk)
            counter += 1
            active = active & (np.abs(R(dk)) > 1e-12)
        nvec = dev[idx] / dn[:, None]
        dLambda = np.sqrt(3 / 2) * dk
        knew = kk + dk
        out_k[idx] = knew
        out_s[idx] = trial[idx] - 2.0 * G * dLambda[:, None] * nvec
        fac_n = 2.0 * G * (1.0 / (1.0 + dfy_dk(knew) / (3.0 * G)) - 2.0 * G * dLambda / dn)
        fac_d = 4.0 * G**2 * dLambda / dn
        out_C[idx] = Ei[None] - fac_n[:, None, None] * np.einsum("pi,pj->pij", nvec, nvec) - fac_d[:, None, None] * _IDEV_HALF[None]
        failed[idx] = fail_local
    return out_s.reshape(shp + (6,)), out_C.reshape(shp + (6, 6)), out_k.reshape(shp + (1,)), failed.reshape(shp)


def _voigt_stress(T):
    """Voigt order 11,22,33,12,23,13.  ref: utils/voigtnotation.py:74-93."""
    return np.stack([T[..., 0, 0], T[..., 1, 1], T[..., 2, 2], T[..., 0, 1], T[..., 1, 2], T[..., 0, 2]], axis=-1)


def _voigt_strain(E):
    """Voigt strain with doubled shear, order 11,22,33,12,23,13.  ref: utils/voigtnotation.py:32-52."""
    return np.stack([E[..., 0, 0], E[..., 1, 1], E[..., 2, 2], 2 * E[..., 0, 1], 2 * E[..., 1, 2], 2 * E[..., 0, 2]], axis=-1)


def neo_hooke(kind, props, F):
    """Kirchhoff stress tau (3x3), A = d tau / d F (3x3x3x3) and energy.
    ref: materials/neohooke/neohookepencegouformulation{a,b,c}.py:130-145."""
    mu, K = float(props[0]), float(props[1])
    I3 = np.eye(3)
    invF = np.linalg.inv(F)
    B = F @ np.swapaxes(F, -1, -2)
    J = np.linalg.det(F)
    J_ = J[..., None, None]
    e1 = np.einsum("ik,...jl->...ijkl", I3, F) + np.einsum("...il,jk->...ijkl", F, I3)
    e_inv = np.einsum("ij,...lk->...ijkl", I3, invF)
    if kind == "neohookewa":
        I1 = np.trace(B, axis1=-2, axis2=-1)
        lambdaBar = (K - 2 / 3 * mu) * (J**2 - J) - mu
        muBar = (K - 2 / 3 * mu) * (2 * J**2 - J)
        tau = mu * B + lambdaBar[..., None, None] * I3
        A = mu * e1 + muBar[..., None, None, None, None] * e_inv
        energy = mu / 2 * (I1 - 3) + (K / 2 - mu / 3) * (J - 1) ** 2 - mu * np.log(J)
    elif kind == "neohookewb":
        I1 = np.trace(B, axis1=-2, axis2=-1)
        lambdaHat = K / 2 * (J**2 + 1 / J**2)
        muBar = mu / (3 * J ** (2 / 3))
        lambdaBar = K / 4 * (J**2 - 1 / J**2) - muBar * I1
        tau = mu / J_ ** (2 / 3) * B + lambdaBar[..., None, None] * I3
        m5 = muBar[..., None, None, None, None]
        A = (
            3 * m5 * e1
            - 2 * m5 * np.einsum("...ij,...lk->...ijkl", B, invF)
            + (lambdaHat + 2 / 3 * I1 * muBar)[..., None, None, None, None] * e_inv
            - 2 * m5 * np.einsum("ij,...kl->...ijkl", I3, F)
        )
        energy = mu / 2 * (I1 / J ** (2 / 3) - 3) + K / 8 * (J**2 + 1 / J**2 - 2)
    elif kind == "neohookewc":
        I1 = np.trace(F, axis1=-2, axis2=-1)  # sic: the reference uses trace(F) here (…c.py:134)
        muBar = mu * J ** (2 / 3 - K / mu)
        lambdaBar = (K / mu - 2 / 3) * muBar
        tau = mu * B - muBar[..., None, None] * I3
        A = mu * e1 + lambdaBar[..., None, None, None, None] * e_inv
        energy = mu / 2 * (I1 - 3) + 3 * mu**2 / (3 * K - 2 * mu) * (J ** (2 / 3 - K / mu) - 1)
    else:
        raise ValueError(kind)
    return tau, A, energy


# --------------------------------------------------------------------------------------
# element loops
# --------------------------------------------------------------------------------------

def compute_small_strain(eltype, material, props, X, Ue, dUe, stateRef):
    """DisplacementElement.computeYourself, batched.
    ref: elements/displacementelement/element.py:290-346.
    X [nEl,n,3], Ue/dUe [nEl,3n], stateRef [nEl,nGp,12+m] -> Ke [nEl,3n,3n] (row-major as the
    element writes it), Pe [nEl,3n], stateTemp, failed[nEl,nGp]."""
    nn = ELEMENT_INFO[eltype]["nnodes"]
    xi, eta, zeta, w = gauss_points(eltype)
    dN = shape_derivatives(nn, xi, eta, zeta)
    J = jacobians(dN, X)
    detJ = np.linalg.det(J)
    B = b_operator(nabla_n(dN, J))
    dstrain = np.einsum("egvd,ed->egv", B, dUe)
    state = stateRef.copy()
    stress0 = state[..., 0:6]
    if material == "linearelastic":
        stress, C, mstate, failed = linear_elastic(props, stress0, dstrain)
    elif material == "vonmises":
        stress, C, mstate, failed = von_mises(props, stress0, dstrain, state[..., 12])
    else:
        raise ValueError("small-strain element needs a hypo-elastic material (element.py:333-334)")
    scale = detJ * w[None, :]
    Ke = np.einsum("egvi,egvw,egwj,eg->eij", B, C, B, scale, optimize=True)
    Pe = -np.einsum("egvi,egv,eg->ei", B, stress, scale, optimize=True)
    state[..., 0:6] = stress
    state[..., 6:12] += dstrain
    if mstate.shape[-1]:
        state[..., 12:] = mstate
    return Ke, Pe, state, failed


def compute_tl_hyperelastic(eltype, material, props, X, Ue, dUe, stateRef):
    """DisplacementTLElement.computeYourself, hyperelastic branch, batched.
    ref: elements/displacementtlelement/element.py:346-427 (branch :391-414)."""
    nn = ELEMENT_INFO[eltype]["nnodes"]
    xi, eta, zeta, w = gauss_points(eltype)
    dN = shape_derivatives(nn, xi, eta, zeta)
    J = jacobians(dN, X)
    detJ = np.linalg.det(J)
    gN = nabla_n(dN, J)  # [e,g,3,a]
    u = Ue.reshape(Ue.shape[0], nn, 3)
    F = np.eye(3)[None, None] + np.einsum("eai,egja->egij", u, gN)  # _elementcomputationmatrices.py:99-105
    H = F - np.eye(3)
    Egl = 0.5 * (H + np.swapaxes(H, -1, -2) + np.swapaxes(H, -1, -2) @ H)
    invF = np.linalg.inv(F)
    NAi = np.einsum("egja,egjk->egak", gN, invF)  # nablaN^T . invF
    tau, A, energy = neo_hooke(material, props, F)
    PK1 = invF @ tau
    Hk = np.einsum("egai,egijkl,eglb->egajbk", NAi, A, gN, optimize=True) - np.einsum(
        "egak,egbi,egij->egajbk", NAi, NAi, tau, optimize=True
    )
    scale = detJ * w[None, :]
    nd = 3 * nn
    Ke = np.einsum("egajbk,eg->eajbk", Hk, scale).reshape(-1, nd, nd)
    Pe = -np.einsum("egja,egjk,eg->eak", gN, PK1, scale).reshape(-1, nd)
    state = stateRef.copy()
    state[..., 0:6] = _voigt_stress(tau)
    state[..., 6:12] = _voigt_strain(Egl)
    state[..., 12] = energy
    return Ke, Pe, state, np.zeros(state.shape[:2], dtype=bool)


def _unvoigt_stress(s):
    """ref: utils/voigtnotation.py:95-113 (order 11,22,33,12,23,13)."""
    T = np.empty(s.shape[:-1] + (3, 3))
    T[..., 0, 0], T[..., 1, 1], T[..., 2, 2] = s[..., 0], s[..., 1], s[..., 2]
    T[..., 0, 1] = T[..., 1, 0] = s[..., 3]
    T[..., 1, 2] = T[..., 2, 1] = s[..., 4]
    T[..., 0, 2] = T[..., 2, 0] = s[..., 5]
    return T


def b_operator_tl(F, gN):
    """Nonlinear B [e,g,6,3n], Voigt rows 11,22,33,12,23,13 with engineering shear.
    ref: elements/displacementtlelement/_elementcomputationmatrices.py:867-916 (_B03D)."""
    e, g, _, nn = gN.shape
    B = np.zeros((e, g, 6, nn, 3))
    dX, dY, dZ = gN[:, :, 0, :, None], gN[:, :, 1, :, None], gN[:, :, 2, :, None]  # [e,g,a,1]
    Fc = lambda c: F[:, :, None, :, c]  # noqa: E731  column c of F as [e,g,1,k]
    B[:, :, 0] = dX * Fc(0)
    B[:, :, 1] = dY * Fc(1)
    B[:, :, 2] = dZ * Fc(2)
    B[:, :, 3] = dX * Fc(1) + dY * Fc(0)
    B[:, :, 4] = dY * Fc(2) + dZ * Fc(1)
    B[:, :, 5] = dX * Fc(2) + dZ * Fc(0)
    return B.reshape(e, g, 6, 3 * nn)


def compute_tl_hypoelastic(eltype, material, props, X, Ue, dUe, stateRef):
    """DisplacementTLElement.computeYourself, non-hyperelastic branch (material tangent B^T C B plus the geometric
    stiffness Hgeo), batched.  ref: elements/displacementtlelement/element.py:373-427 (branch :415-425), Hgeo :48-73.
    The element keeps the last accepted Green-Lagrange strain in `_Eold` (:215, :461); it equals the strain part of the
    accepted state, [6:12] = sum of dStrain = Voigt(E_old), which is what this restatement (and the device) uses."""
    nn = ELEMENT_INFO[eltype]["nnodes"]
    xi, eta, zeta, w = gauss_points(eltype)
    dN = shape_derivatives(nn, xi, eta, zeta)
    J = jacobians(dN, X)
    detJ = np.linalg.det(J)
    gN = nabla_n(dN, J)  # [e,g,3,a]
    u = Ue.reshape(Ue.shape[0], nn, 3)
    F = np.eye(3)[None, None] + np.einsum("eai,egja->egij", u, gN)
    H = F - np.eye(3)
    Egl = 0.5 * (H + np.swapaxes(H, -1, -2) + np.swapaxes(H, -1, -2) @ H)
    state = stateRef.copy()
    dstrain = _voigt_strain(Egl) - state[..., 6:12]
    stress0 = state[..., 0:6]
    if material == "linearelastic":
        stress, C, mstate, failed = linear_elastic(props, stress0, dstrain)
    elif material == "vonmises":
        stress, C, mstate, failed = von_mises(props, stress0, dstrain, state[..., 12])
    else:
        raise ValueError(material)
    B = b_operator_tl(F, gN)
    S = _unvoigt_stress(stress)
    Hsub = np.einsum("egia,egij,egjb->egab", gN, S, gN)  # nablaN^T S nablaN
    scale = detJ * w[None, :]
    nd = 3 * nn
    Ke = np.einsum("egvi,egvw,egwj,eg->eij", B, C, B, scale, optimize=True)
    Ke = Ke + np.einsum("egab,kl,eg->eakbl", Hsub, np.eye(3), scale).reshape(-1, nd, nd)
    Pe = -np.einsum("egvi,egv,eg->ei", B, stress, scale, optimize=True)
    state[..., 0:6] = stress
    state[..., 6:12] += dstrain
    if mstate.shape[-1]:
        state[..., 12:] = mstate
    return Ke, Pe, state, failed


def shape_functions_bodyforce(nnodes: int, xi, eta, zeta):
    """N[gp, a] as `computeNOperator` evaluates it (ref: displacementelement/_elementcomputationmatrices.py:106-212).
    Its node table has xi and eta swapped relative to the derivative tables: node a sits at
    (xi, eta, zeta) = (eta_a, xi_a, zeta_a) of the derivative convention.  Replicated, not "fixed"."""
    xi, eta, zeta = (np.atleast_1d(np.asarray(v, dtype=float)) for v in (xi, eta, zeta))
    lc = _local_coords(HEXA8_OFFSETS, 1) if nnodes == 8 else _local_coords(HEXA20_OFFSETS, 2)
    a, b, c = lc[:, 1][None, :], lc[:, 0][None, :], lc[:, 2][None, :]  # swapped: a multiplies xi, b multiplies eta
    X, E, Z = xi[:, None], eta[:, None], zeta[:, None]
    fx, fe, fz = 1 + a * X, 1 + b * E, 1 + c * Z
    if nnodes == 8:
        return fx * fe * fz / 8
    corner = fx * fe * fz * (a * X + b * E + c * Z - 2) / 8
    return np.where(a == 0, (1 - X**2) * fe * fz / 4, np.where(b == 0, fx * (1 - E**2) * fz / 4, np.where(c == 0, fx * fe * (1 - Z**2) / 4, corner)))


def body_force(eltype, coords, conn, load):
    """PExt[el] += sum_gp outer(N[gp], load) detJ w for every element.
    ref: elements/displacementelement/element.py:348-371 (computeBodyForce), solvers/nonlinearimplicitstatic.py:545-553."""
    nn = ELEMENT_INFO[eltype.upper()]["nnodes"]
    xi, eta, zeta, w = gauss_points(eltype.upper())
    dN = shape_derivatives(nn, xi, eta, zeta)
    N = shape_functions_bodyforce(nn, xi, eta, zeta)
    detJ = np.linalg.det(jacobians(dN, coords[conn]))
    Pe = np.einsum("ga,i,eg,g->eai", N, np.asarray(load, dtype=float), detJ, w).reshape(conn.shape[0], -1)
    dofs = element_dofs(conn)
    return np.bincount(dofs.reshape(-1), weights=Pe.reshape(-1), minlength=3 * coords.shape[0]), Pe


# Abaqus face numbering of the 8-node hexahedron (faces 1..6), local nodes ordered so that the right-hand rule gives the OUTWARD normal.
HEX8_FACES = np.array([[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7]])


def surface_pressure(coords, conn, elems, faces, pressure):
    """PExt of a pressure load (positive = pushing into the element) on (element, face id 1..6) pairs of 8-node hexahedra:
    f_a = -p int N_a n dA with the bilinear face interpolation and 2x2 Gauss points; dead load on the reference geometry.
    The algorithm is the Marmot displacement element's (un-vendored third-party code, SURVEY §8c: pinned only by
    testfiles/LinearElasticIsotropic/U.ref); called from solvers/nonlinearimplicitstatic.py:460-508."""
    P = np.zeros(3 * coords.shape[0])
    g = 1.0 / np.sqrt(3.0)
    rs = np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]])
    for e, f in zip(np.asarray(elems), np.asarray(faces)):
        nodes = conn[e][HEX8_FACES[f - 1]]
        X = coords[nodes]
        for r, s_ in g * rs:
            N = 0.25 * (1 + rs[:, 0] * r) * (1 + rs[:, 1] * s_)
            dr = 0.25 * rs[:, 0] * (1 + rs[:, 1] * s_)
            ds = 0.25 * rs[:, 1] * (1 + rs[:, 0] * r)
            nA = np.cross(dr @ X, ds @ X)
            for a in range(4):
                P[3 * nodes[a] : 3 * nodes[a] + 3] -= pressure * N[a] * nA
    return P


def compute_elements(eltype, material, props, coords, conn, U, dU, stateRef, chunk=4096):
    """Batched element evaluation in element order; returns Ke, Pe, stateTemp, failed."""
    eltype = eltype.upper()
    material = material.lower()
    dofs = element_dofs(conn)
    if ELEMENT_INFO[eltype]["tl"]:
        fn = compute_tl_hyperelastic if material.startswith("neohooke") else compute_tl_hypoelastic  # element.py:333-334
    else:
        fn = compute_small_strain
    outs = []
    for s in range(0, conn.shape[0], chunk):
        sl = slice(s, s + chunk)
        outs.append(fn(eltype, material, props, coords[conn[sl]], U[dofs[sl]], dU[dofs[sl]], stateRef[sl]))
    return tuple(np.concatenate([o[i] for o in outs]) for i in range(4))


def assemble(eltype, material, props, coords, conn, U, dU, stateRef, chunk=4096, want_vij=True):
    """One NIST.computeElements pass + CSRGenerator.updateCSR.
    ref: solvers/nonlinearimplicitstatic.py:794-849, numerics/csrgenerator.pyx:100-115.
    Returns dict(V, I, J, indptr, indices, x, data, P, F, stateTemp, failed)."""
    dofs = element_dofs(conn)
    nDof = 3 * coords.shape[0]
    Ke, Pe, state, failed = compute_elements(eltype, material, props, coords, conn, U, dU, stateRef, chunk)
    V = Ke.reshape(-1)  # row-major element write into the slice (element.py:318)
    I, J = vij_pattern(dofs)  # noqa: E741
    indptr, indices, x = csr_pattern(I, J, nDof)
    data = update_csr(x, V, indices.size)
    P = np.bincount(dofs.reshape(-1), weights=Pe.reshape(-1), minlength=nDof)
    Fv = np.bincount(dofs.reshape(-1), weights=np.abs(Pe).reshape(-1), minlength=nDof)
    out = dict(indptr=indptr, indices=indices, x=x, data=data, P=P, F=Fv, stateTemp=state, failed=failed)
    if want_vij:
        out.update(V=V, I=I, J=J)
    return out
