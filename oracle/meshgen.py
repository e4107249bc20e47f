"""oracle/meshgen.py -- TEST INFRASTRUCTURE ONLY.

numpy generator of the synthetic hex box in OpenFOAM blockMesh ordering
(cell id = i + nx*(j + ny*k), x fastest), independent of
the product's own C++ generator so that the two can be checked against each
other.
"""
import numpy as np


def hex_box(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, origin=(0.0, 0.0, 0.0)):
    """Returns dict(C [N,3], V [N], points [M,3], boxN int32[3], boxGeom f64[6])."""
    hx, hy, hz = lx / nx, ly / ny, lz / nz
    i = np.arange(nx, dtype=np.float64)
    j = np.arange(ny, dtype=np.float64)
    k = np.arange(nz, dtype=np.float64)
    cx = origin[0] + (i + 0.5) * hx
    cy = origin[1] + (j + 0.5) * hy
    cz = origin[2] + (k + 0.5) * hz
    C = np.empty((nz, ny, nx, 3), dtype=np.float64)
    C[..., 0] = cx[None, None, :]
    C[..., 1] = cy[None, :, None]
    C[..., 2] = cz[:, None, None]
    C = C.reshape(-1, 3)
    V = np.full(nx * ny * nz, hx * hy * hz, dtype=np.float64)
    px = origin[0] + np.arange(nx + 1, dtype=np.float64) * hx
    py = origin[1] + np.arange(ny + 1, dtype=np.float64) * hy
    pz = origin[2] + np.arange(nz + 1, dtype=np.float64) * hz
    P = np.empty((nz + 1, ny + 1, nx + 1, 3), dtype=np.float64)
    P[..., 0] = px[None, None, :]
    P[..., 1] = py[None, :, None]
    P[..., 2] = pz[:, None, None]
    return dict(C=np.ascontiguousarray(C), V=V, points=np.ascontiguousarray(P.reshape(-1, 3)),
                boxN=np.array([nx, ny, nz], dtype=np.int32),
                boxGeom=np.array([origin[0], origin[1], origin[2], hx, hy, hz], dtype=np.float64),
                n=(nx, ny, nz), h=(hx, hy, hz))
