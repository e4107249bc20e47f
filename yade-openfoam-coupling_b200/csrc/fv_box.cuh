// fv_box.cuh -- device-side description of the uniform hex box the finite-volume kernels run on, and the
// helpers every FV kernel shares (cell <-> (i,j,k), face-slot layout, boundary coefficients, deterministic
// grid reductions).
//
// LAYOUT (HBM).  Cells are x-fastest, c = i + nx*(j + ny*k) (OpenFOAM blockMesh order).  OpenFOAM's LDU
// face list (owner/neighbour int32 pairs, upper-triangular order) is NOT kept on the device: on the box the
// "+d" face of cell c is implicit, so every face field / matrix coefficient is stored in OWNER SLOTS
//     F[d*N + c]  = value on the +d face of cell c      (d = 0,1,2)
// which for cells on a max side (i = nx-1, ...) is that side's boundary face.  The min-side boundary faces
// follow at  3N + (j + ny*k)  (xmin),  3N + ny*nz + (i + nx*k)  (ymin),  3N + ny*nz + nx*nz + (i + nx*j)
// (zmin).  A stencil therefore reads coalesced streams with no index arrays (Amul: 56 B/cell instead of
// ~120 B/cell with explicit LDU addressing).
//
// ARITHMETIC ORDER.  Every kernel visits the faces of a cell in the order OpenFOAM's face loops would reach
// them: the faces on which the cell is the neighbour (owners c-nx*ny, c-nx, c-1, i.e. z-, y-, x-), then
// the faces it owns (x+, y+, z+), then its boundary faces in boundary (patch) order.  With -fmad=false
// this makes the per-cell results bit-identical to the sequential CPU loops; only the global sums
// (dot products, residual norms) are associated differently.
#pragma once
#include <cuda_runtime.h>

#include "fv_peer.cuh"

enum { FV_FIXED_VALUE = 0, FV_ZERO_GRADIENT = 1, FV_EMPTY = 2, FV_FIXED_FLUX_PRESSURE = 3 };   // the last one: p only (pimpleFoamYade/pEqn.H:21)
enum { FV_PRECOND_DIC = 0, FV_PRECOND_DIAGONAL = 1, FV_PRECOND_NONE = 2 };

struct BoxGeom {
    int nx, ny, nz, N, sy, sz;
    int off[3];              // first min-side slot of direction d
    int nSlots;
    double Sf[3], magSf[3], dc[3], w[3], V;     // internal faces of direction d (uniform box)
    // boundary sides: 0 xmin 1 xmax 2 ymin 3 ymax 4 zmin 5 zmax
    double bSf[6];           // signed normal component of the outward area vector
    double bMagSf[6], bDc[6];
    int kindU[6], kindP[6];
    double valU[6][3], valP[6];
    int seq[6];              // sides in boundary-list (patch) order
    int valid[3];            // solved vector components (an EMPTY side pair removes its direction)
    int pNeedRef;
    int reconRm[3];          // fvc::reconstruct: directions tensorField inv() removes (no faces: empty patch pair)
    double sumV;             // sum of cell volumes, accumulated sequentially like the CPU loop
    // fixedFluxPressure sides: snGrad(p) of every boundary face, set by constrainPressure (owner-slot layout); else null
    const double* bGradP;
};

__device__ __forceinline__ void fvIJK(const BoxGeom& g, int c, int& i, int& j, int& k)
{
    i = c % g.nx;
    const int r = c / g.nx;
    j = r % g.ny;
    k = r / g.ny;
}
__device__ __forceinline__ int fvIdx(int d, int i, int j, int k) { return d == 0 ? i : (d == 1 ? j : k); }
__device__ __forceinline__ int fvN(const BoxGeom& g, int d) { return d == 0 ? g.nx : (d == 1 ? g.ny : g.nz); }
__device__ __forceinline__ int fvStride(const BoxGeom& g, int d) { return d == 0 ? 1 : (d == 1 ? g.sy : g.sz); }

// is cell (i,j,k) on boundary side s?
__device__ __forceinline__ bool fvOnSide(const BoxGeom& g, int s, int i, int j, int k)
{
    const int d = s >> 1, v = fvIdx(d, i, j, k);
    return (s & 1) ? (v == fvN(g, d) - 1) : (v == 0);
}
// slot of the boundary face of cell c on side s
__device__ __forceinline__ int fvSideSlot(const BoxGeom& g, int s, int c, int i, int j, int k)
{
    const int d = s >> 1;
    if (s & 1) return d * g.N + c;
    return g.off[d] + (d == 0 ? j + g.ny * k : (d == 1 ? i + g.nx * k : i + g.nx * j));
}

// linear interpolation as OpenFOAM writes it: lambda*(P - N) + N
__device__ __forceinline__ double fvLerp(double w, double P, double N) { return w * (P - N) + N; }

// UEqn boundary coefficients of one boundary face (side s, face flux phib), component j:
//   convection  ic = phib*valueInternalCoeffs      bc = -phib*valueBoundaryCoeffs
//   laplacian   ic = (nu magSf)*gradientInternalCoeffs   bc = -(nu magSf)*gradientBoundaryCoeffs
//   UEqn = ddt + div - laplacian  =>  ic = icC - icL, bc = bcC - bcL
__device__ __forceinline__ void fvBCoefU(const BoxGeom& g, int s, double phib, double nu, int j, double& ic, double& bc)
{
    const double gMagSf = nu * g.bMagSf[s];
    double icC, bcC, icL, bcL;
    if (g.kindU[s] == FV_FIXED_VALUE) {
        icC = phib * 0.0;
        bcC = -phib * g.valU[s][j];
        icL = gMagSf * (-1.0 * g.bDc[s]);
        bcL = -gMagSf * (g.bDc[s] * g.valU[s][j]);
    } else {
        icC = phib * 1.0;
        bcC = -phib * 0.0;
        icL = gMagSf * 0.0;
        bcL = -gMagSf * 0.0;
    }
    ic = icC - icL;
    bc = bcC - bcL;
}

// snGrad(p) the fixedFluxPressure patch holds on the boundary face of cell c on side s (0 for every other patch type)
__device__ __forceinline__ double fvFluxGradP(const BoxGeom& g, int s, int c, int i, int j, int k)
{
    return g.kindP[s] == FV_FIXED_FLUX_PRESSURE ? g.bGradP[fvSideSlot(g, s, c, i, j, k)] : 0.0;
}

// pEqn boundary coefficients (fvm::laplacian(gamma, p)) of one boundary face; gammaB = boundary value of gamma;
// gradP = the face's fixedFluxPressure gradient (fixedGradient: gradientInternalCoeffs 0, gradientBoundaryCoeffs = gradient)
__device__ __forceinline__ void fvBCoefP(const BoxGeom& g, int s, double gammaB, double& ic, double& bc, double gradP = 0.0)
{
    ic = 0.0;
    bc = 0.0;
    if (g.kindP[s] == FV_FIXED_FLUX_PRESSURE) {
        const double pGamma = gammaB * g.bMagSf[s];
        ic = pGamma * 0.0;
        bc = -pGamma * gradP;
        return;
    }
    if (g.kindP[s] != FV_FIXED_VALUE) return;
    const double pGamma = gammaB * g.bMagSf[s];
    ic = pGamma * (-1.0 * g.bDc[s]);
    bc = -pGamma * (g.bDc[s] * g.valP[s]);
}

// ---------------------------------------------------------------------------------------------
// deterministic grid reduction: every block leaves NV partial sums, the last block to finish adds
// them up in block order and hands the totals to `fin` (one thread).  For a fixed launch geometry
// the result is bit-reproducible run to run.
// ---------------------------------------------------------------------------------------------
struct FvRed {
    double* partial;         // [NV][gridDim.x]
    unsigned int* ticket;    // zero-initialised; reset by the last block
    double* distOut;         // decomposed solve, NCCL path: the totals go here (this rank's partial sums) instead of to
                             // `fin`; the host all-reduces them over the ranks and runs the finishing kernel
    PeerDev* peer;           // decomposed solve, peer-memory path (fv_peer.cuh): the last block all-reduces the totals over
                             // the ranks itself and then runs `fin`
};

template <int NV, bool MAXFIRST, int BLOCK, class Fin>
__device__ __forceinline__ void fvGridReduce(double (&v)[NV], const FvRed& r, Fin fin)
{
    __shared__ double sh[NV][BLOCK / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = (MAXFIRST && q == 0) ? fmax(x, y) : x + y;
        }
        if (lane == 0) sh[q][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            double x = sh[q][0];
            for (int w = 1; w < BLOCK / 32; ++w) x = (MAXFIRST && q == 0) ? fmax(x, sh[q][w]) : x + sh[q][w];
            r.partial[q * gridDim.x + blockIdx.x] = x;
        }
        __threadfence();
        const unsigned int t = atomicAdd(r.ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double tot[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const volatile double* p = r.partial + q * gridDim.x;
        const bool mx = MAXFIRST && q == 0;
        double x = mx ? -1.7976931348623157e308 : 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += BLOCK) x = mx ? fmax(x, p[b]) : x + p[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = mx ? fmax(x, y) : x + y;
        }
        __syncthreads();
        if (lane == 0) sh[q][warp] = x;
        __syncthreads();
        x = sh[q][0];
        for (int w = 1; w < BLOCK / 32; ++w) x = mx ? fmax(x, sh[q][w]) : x + sh[q][w];
        tot[q] = x;
    }
    if constexpr (NV <= 2 && !MAXFIRST)
        if (r.peer && threadIdx.x < 32) peerAllReduce<NV>(r.peer, tot);      // (warp 0; `tot` of thread 0 counts)
    if (threadIdx.x == 0) {
        *r.ticket = 0u;
        if (r.distOut) {
#pragma unroll
            for (int q = 0; q < NV; ++q) r.distOut[q] = tot[q];
        } else {
            fin(tot);
        }
    }
}
