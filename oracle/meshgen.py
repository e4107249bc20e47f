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


SIDES = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")
BC_FIXED_VALUE, BC_ZERO_GRADIENT, BC_EMPTY = 0, 1, 2
BC_FIXED_FLUX_PRESSURE = 3          # p only; honoured by the pimpleFoamYade restatement (oracle/fv_oracle.cc pimpleSolve)


def hex_box_ldu(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, origin=(0.0, 0.0, 0.0), patches=None):
    """hex_box() plus OpenFOAM LDU addressing and boundary patches.

    Internal faces in upper-triangular order (per owner cell its +x, +y, +z face).  `patches` is a list of
    (name, [sides...]) in boundary order, default one patch per side in SIDES order; the faces of a patch are
    its sides in the listed order, each side by increasing owner cell.  BCs default to U fixedValue 0 /
    p zeroGradient; change them with set_bc().
    """
    m = hex_box(nx, ny, nz, lx, ly, lz, origin)
    hx, hy, hz = m["h"]
    N = nx * ny * nz
    own, nei, dr = [], [], []
    idx = np.arange(N, dtype=np.int64).reshape(nz, ny, nx)
    # candidate (+x, +y, +z) neighbour of every cell, -1 where there is none; flattening [cell][dir] gives owner-major order
    cand = np.full((nz, ny, nx, 3), -1, dtype=np.int64)
    cand[:, :, :-1, 0] = idx[:, :, 1:]
    cand[:, :-1, :, 1] = idx[:, 1:, :]
    cand[:-1, :, :, 2] = idx[1:, :, :]
    cand = cand.reshape(N, 3)
    cell, d = np.nonzero(cand >= 0)
    owner = cell.astype(np.int32)
    neigh = cand[cell, d].astype(np.int32)
    area = np.array([hy * hz, hx * hz, hx * hy])
    dist = np.array([hx, hy, hz])
    Fi = owner.size
    Sf = np.zeros((Fi, 3))
    Sf[np.arange(Fi), d] = area[d]
    m.update(nCells=N, nInternalFaces=Fi, owner=owner, neighbour=neigh, Sf=Sf, magSf=area[d].copy(),
             weights=np.full(Fi, 0.5), deltaCoeffs=(1.0 / dist)[d].copy())
    side_cells = dict(xmin=idx[:, :, 0], xmax=idx[:, :, -1], ymin=idx[:, 0, :], ymax=idx[:, -1, :],
                      zmin=idx[0, :, :], zmax=idx[-1, :, :])
    if patches is None:
        patches = [(s, [s]) for s in SIDES]
    out = []
    for name, sides in patches:
        fc, sf, ms, dc = [], [], [], []
        for s in sides:
            ax = SIDES.index(s) // 2
            sign = -1.0 if s.endswith("min") else 1.0
            c = np.sort(side_cells[s].reshape(-1)).astype(np.int32)
            v = np.zeros((c.size, 3))
            v[:, ax] = sign * area[ax]
            fc.append(c)
            sf.append(v)
            ms.append(np.full(c.size, area[ax]))
            dc.append(np.full(c.size, 1.0 / (0.5 * dist[ax])))
        out.append(dict(name=name, faceCells=np.concatenate(fc), Sf=np.concatenate(sf), magSf=np.concatenate(ms),
                        deltaCoeffs=np.concatenate(dc), bcU=BC_FIXED_VALUE, valueU=(0.0, 0.0, 0.0),
                        bcP=BC_ZERO_GRADIENT, valueP=0.0))
    m["patches"] = out
    m["bbox"] = np.array([origin[0], origin[1], origin[2], origin[0] + lx, origin[1] + ly, origin[2] + lz])
    return m


def set_bc(mesh, name, bcU=None, valueU=None, bcP=None, valueP=None):
    for p in mesh["patches"]:
        if p["name"] == name:
            if bcU is not None:
                p["bcU"] = bcU
            if valueU is not None:
                p["valueU"] = tuple(float(v) for v in valueU)
            if bcP is not None:
                p["bcP"] = bcP
            if valueP is not None:
                p["valueP"] = float(valueP)
            return
    raise KeyError(name)
