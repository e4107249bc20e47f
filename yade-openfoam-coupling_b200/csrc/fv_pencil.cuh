// fv_pencil.cuh -- the "pencil" layout the sequential LDU recurrences (DIC factorisation, DIC forward /
// backward substitution, Gauss-Seidel sweeps) and the Krylov kernels around them run in.
//
// WHY.  OpenFOAM's DIC / Gauss-Seidel loops are sequential over the face list; on the lexicographic hex
// box cell (i,j,k) depends on (i-1,j,k), (i,j-1,k), (i,j,k-1).  A hyperplane wavefront with one grid
// barrier per plane costs ~3.5 us x (nx+ny+nz-2) planes = 1.3 ms per sweep at 128^3 against ~20 us of
// HBM time.  Here one WARP owns a pencil of cells instead -- 32 consecutive j (one per lane), all i, one
// k-plane -- and walks it in nx+31 steps: at step t lane l works on i = t-l.  The x-dependency stays in a
// register, the y-dependency is one __shfl from the neighbouring lane.  Only two values cross warps: the
// z-neighbour (the same row of the plane behind) and the y-neighbour of the warp's edge lane.  They travel
// as 8-byte words that are data and flag at once: every slot is pre-armed with a signalling-NaN sentinel and
// the consumer polls it until it is not the sentinel (no fence, no grid barrier) -- through shared-memory
// channels inside a CTA, distributed shared memory inside a thread-block cluster, and, between clusters and
// j-blocks, through the output array itself in L2 (read by helper warps).  The operation order inside a cell is
// exactly OpenFOAM's, so results stay bit-identical to the sequential loops.  See fv_pencil.cu (k_pencil).
//
// LAYOUT (HBM).  So that every warp access is one contiguous 256-byte row, the solver's vectors and
// matrix coefficients live in a skewed layout: slab (k, jb = j/32) holds Tp rows ("slots") of 32 lanes,
//     pos(i,j,k) = ((k*nJB + jb)*Tp + i + (j&31))*32 + (j&31),       Tp = roundup(nx+31, 32)
// i.e. row m of a slab holds the 32 mutually independent cells i = m-lane that a warp processes in one
// step (for both sweep directions).  Every lane prefetches its own words with 8-byte cp.async into a private
// shared-memory ring 8 rows ahead.  The layout costs (nx+31)/nx extra storage; pads are zero, never armed, and
// every array carries guard rows so that the sweeps need no bounds tests.
#pragma once
#include <cstdint>

#include "fv_box.cuh"

constexpr unsigned long long PEN_SENT = 0x7FF4DEADBEEF5A5AULL;   // signalling NaN: arithmetic never produces it

struct PencilGeom {
    int nx, ny, nz, N;
    int nJB;               // j-blocks of 32 lanes
    int Tp;                // rows per slab
    long long nRows;       // nz*nJB*Tp
    long long NP;          // nRows*32 doubles per vector
    long long zStride;     // nJB*Tp*32: distance between (i,j,k) and (i,j,k+1)
    // the part of the box this process works on (everything, unless the solve is decomposed into z slabs: fv_dist.cu)
    int kLo, kHi;          // k-planes [kLo, kHi)
    int jbLo, jbHi;        // j-blocks [jbLo, jbHi)
    long long nLoc;        // local rows: (kHi-kLo) * (jbHi-jbLo) * Tp, enumerated plane by plane, j-block by j-block
};

// local row l -> row of the global layout
__host__ __device__ __forceinline__ long long penGlobalRow(const PencilGeom& g, long long l)
{
    const int nJl = g.jbHi - g.jbLo;
    const int q = (int)(l / g.Tp), m = (int)(l - (long long)q * g.Tp);
    const int kk = q / nJl, jj = q - kk * nJl;
    return ((long long)(g.kLo + kk) * g.nJB + g.jbLo + jj) * g.Tp + m;
}

struct PenCell {
    bool valid;
    int i, j, k, lane;
    long long pos;
};

__device__ __forceinline__ PenCell penDecode(const PencilGeom& g, long long row, int lane)
{
    PenCell c;
    const int r32 = (int)row;                              // nRows < 2^31 (checked at creation): 32-bit divisions
    const int sb = r32 / g.Tp;
    const int m = r32 - sb * g.Tp;
    c.k = sb / g.nJB;
    const int jb = sb - c.k * g.nJB;
    c.lane = lane;
    c.i = m - lane;
    c.j = jb * 32 + lane;
    c.valid = c.i >= 0 && c.i < g.nx && c.j < g.ny;
    c.pos = row * 32 + lane;
    return c;
}
__host__ __device__ __forceinline__ long long penPos(const PencilGeom& g, int i, int j, int k)
{
    return (((long long)k * g.nJB + (j >> 5)) * g.Tp + i + (j & 31)) * 32 + (j & 31);
}
// position of (i, j-1, k) / (i, j+1, k)
__device__ __forceinline__ long long penYm(const PencilGeom& g, const PenCell& c)
{
    return c.lane > 0 ? c.pos - 33 : c.pos - (long long)g.Tp * 32 + 31 * 32 + 31;
}
__device__ __forceinline__ long long penYp(const PencilGeom& g, const PenCell& c)
{
    return c.lane < 31 ? c.pos + 33 : c.pos + (long long)g.Tp * 32 - 31 * 32 - 31;
}

// an LDU matrix in pencil layout: low[d] = coefficient of the face towards the lower neighbour in
// direction d (lduMatrix::lower of the face owned by that neighbour -- or ::upper of it for the DIC
// recurrences of a symmetric matrix, where both are the same array), up[d] = lduMatrix::upper of the
// cell's own +d face.  Zero where there is no such neighbour.
struct PenMatrix {
    double* dg;
    double* low[3];
    double* up[3];
};
