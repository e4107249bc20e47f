"""Synthetic hex-box mesh in OpenFOAM blockMesh ordering, as the LDU description fy_create takes.

Cells: id = i + nx*(j + ny*k) (x fastest).  Internal faces in upper-triangular order: for every
cell in index order its +x, +y, +z neighbour faces (owner < neighbour, sorted by owner then
neighbour) -- the order OpenFOAM's lduAddressing requires.  Six boundary patches in the order
xmin xmax ymin ymax zmin zmax, faces within a patch by increasing owner cell.
"""
import numpy as np

PATCH_NAMES = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")
BC_FIXED_VALUE = 0
BC_ZERO_GRADIENT = 1
BC_EMPTY = 2
BC_FIXED_FLUX_PRESSURE = 3      # p only: gradient set by constrainPressure (pimpleFoamYade/pEqn.H:21)


def box_mesh(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, origin=(0.0, 0.0, 0.0), faces=True, patches=None):
    """`patches`: optional list of (name, [sides...]) grouping the six sides into boundary patches in boundary
    order (e.g. the cavity tutorial's movingWall / fixedWalls / frontAndBack); default one patch per side."""
    hx, hy, hz = lx / nx, ly / ny, lz / nz
    N = nx * ny * nz
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i = i.reshape(-1)
    j = j.reshape(-1)
    k = k.reshape(-1)
    C = np.empty((N, 3), dtype=np.float64)
    C[:, 0] = origin[0] + (i + 0.5) * hx
    C[:, 1] = origin[1] + (j + 0.5) * hy
    C[:, 2] = origin[2] + (k + 0.5) * hz
    V = np.full(N, hx * hy * hz, dtype=np.float64)
    m = dict(nCells=N, n=(nx, ny, nz), h=(hx, hy, hz), C=C, V=V,
             boxN=np.array([nx, ny, nz], dtype=np.int32),
             boxGeom=np.array([origin[0], origin[1], origin[2], hx, hy, hz], dtype=np.float64),
             bbox=np.array([origin[0], origin[1], origin[2], origin[0] + nx * hx, origin[1] + ny * hy,
                            origin[2] + nz * hz], dtype=np.float64),
             nInternalFaces=0, patches=[])
    if not faces:
        return m
    cid = np.arange(N, dtype=np.int64)
    # candidate faces per cell in the order +x, +y, +z; keep the existing ones, row-major => sorted by owner
    has = np.stack([i < nx - 1, j < ny - 1, k < nz - 1], axis=1)
    nb = np.stack([cid + 1, cid + nx, cid + nx * ny], axis=1)
    own = np.repeat(cid[:, None], 3, axis=1)
    direc = np.tile(np.arange(3)[None, :], (N, 1))
    sel = has.reshape(-1)
    owner = own.reshape(-1)[sel].astype(np.int32)
    neigh = nb.reshape(-1)[sel].astype(np.int32)
    d = direc.reshape(-1)[sel]
    area = np.array([hy * hz, hx * hz, hx * hy])
    dist = np.array([hx, hy, hz])
    Fi = owner.shape[0]
    Sf = np.zeros((Fi, 3), dtype=np.float64)
    Sf[np.arange(Fi), d] = area[d]
    m.update(nInternalFaces=Fi, owner=owner, neighbour=neigh, Sf=Sf, magSf=area[d].copy(),
             weights=np.full(Fi, 0.5, dtype=np.float64), deltaCoeffs=(1.0 / dist)[d].copy())
    side = dict(xmin=(i == 0, 0, -1.0), xmax=(i == nx - 1, 0, 1.0), ymin=(j == 0, 1, -1.0), ymax=(j == ny - 1, 1, 1.0),
                zmin=(k == 0, 2, -1.0), zmax=(k == nz - 1, 2, 1.0))
    groups = patches if patches is not None else [(nm, [nm]) for nm in PATCH_NAMES]
    patches = []
    for name, sides in groups:
        fcs, sfs, mss, dcs = [], [], [], []
        for sd in sides:
            mask, ax, sign = side[sd]
            fc = cid[mask].astype(np.int32)
            nf = fc.shape[0]
            psf = np.zeros((nf, 3), dtype=np.float64)
            psf[:, ax] = sign * area[ax]
            fcs.append(fc)
            sfs.append(psf)
            mss.append(np.full(nf, area[ax]))
            dcs.append(np.full(nf, 1.0 / (0.5 * dist[ax])))
        patches.append(dict(name=name, faceCells=np.concatenate(fcs), Sf=np.concatenate(sfs), magSf=np.concatenate(mss),
                            deltaCoeffs=np.concatenate(dcs), bcU=BC_FIXED_VALUE, valueU=(0.0, 0.0, 0.0),
                            bcP=BC_ZERO_GRADIENT, valueP=0.0))
    m["patches"] = patches
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
