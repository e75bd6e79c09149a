"""Array-native BoxGen: the structured hexahedral box generator of the reference
(generators/boxgen.py:91-448) without per-node / per-element Python objects, so that the
1M-8M element benchmark meshes can be built in seconds.

Conventions (verified against the reference generator in tests/golden): nodes numbered with z
fastest, `n = ix*NY*NZ + iy*NZ + iz`; elements ix-major; Hexa20 keeps the grid nodes with
(ix%2)+(iy%2)+(iz%2) < 2 of the doubled grid (boxgen.py:139) in grid order.
"""
import numpy as np

_OFF8 = np.array([(0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0), (0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)], dtype=np.int64)
_OFF20 = np.array(
    [(0, 0, 0), (0, 0, 2), (2, 0, 2), (2, 0, 0), (0, 2, 0), (0, 2, 2), (2, 2, 2), (2, 2, 0),
     (0, 0, 1), (1, 0, 2), (2, 0, 1), (1, 0, 0), (0, 2, 1), (1, 2, 2), (2, 2, 1), (1, 2, 0),
     (0, 1, 0), (0, 1, 2), (2, 1, 2), (2, 1, 0)], dtype=np.int64,
)


def box_mesh(nX, nY, nZ, lX=1.0, lY=1.0, lZ=1.0, x0=0.0, y0=0.0, z0=0.0, elType="C3D8"):
    """Returns (coords float64 [nNode,3], conn int32 [nEl,nNodesPerElement])."""
    n20 = "20" in elType
    m = 2 if n20 else 1
    NX, NY, NZ = m * nX + 1, m * nY + 1, m * nZ + 1
    xs, ys, zs = (np.linspace(a, a + l, N) for a, l, N in ((x0, lX, NX), (y0, lY, NY), (z0, lZ, NZ)))
    coords = np.empty((NX, NY, NZ, 3))
    coords[..., 0] = xs[:, None, None]
    coords[..., 1] = ys[None, :, None]
    coords[..., 2] = zs[None, None, :]
    coords = coords.reshape(-1, 3)
    off = _OFF20 if n20 else _OFF8
    e = np.arange(nX * nY * nZ, dtype=np.int64)
    ex, ey, ez = e // (nY * nZ), (e // nZ) % nY, e % nZ
    grid = (m * ex[:, None] + off[None, :, 0]) * (NY * NZ) + (m * ey[:, None] + off[None, :, 1]) * NZ + (m * ez[:, None] + off[None, :, 2])
    if n20:
        ix, iy, iz = np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij")
        keep = ((ix % 2) + (iy % 2) + (iz % 2) < 2).ravel()
        renum = np.cumsum(keep) - 1
        coords = coords[keep]
        grid = renum[grid]
    return np.ascontiguousarray(coords), np.ascontiguousarray(grid.astype(np.int32))
