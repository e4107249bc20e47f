// fv_box.cu -- the fluid half of the hot path as sm_100a kernels on the uniform hex box:
// icoFoamYade's time step (icoFoamYade/icoFoamYade.C:65-149) = CourantNo, vGrad = fvc::grad(U), UEqn
// assembly, segregated momentum predictor (smoothSolver/symGaussSeidel), PISO correctors with the
// pressure-Poisson PCG (DIC / diagonal / none) inner loop, flux and velocity correction.
//
// The arithmetic each kernel implements is OpenFOAM-6's (the reference only CALLS fvm::/fvc::/solve);
// the per-cell operation order follows OpenFOAM's face loops so that results are bit-identical to the
// CPU restatement in oracle/fv_oracle.cc apart from the association of global sums (see fv_box.cuh).
//
// Sequential recurrences (DIC factorisation and its forward/backward substitutions, Gauss-Seidel sweeps)
// keep OpenFOAM's dependency order: on the lexicographic box cell (i,j,k) depends on (i-1,j,k), (i,j-1,k),
// (i,j,k-1), so all cells of a hyperplane i+j+k = s are independent and are processed together
// (wavefront schedule, one grid-wide barrier per plane inside one cooperative kernel).
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "fv_solver.h"

namespace cg = cooperative_groups;

namespace {
constexpr int BLK = 256;
constexpr double FV_SMALL = 1e-15;     // OpenFOAM `small`
constexpr double FV_VSMALL = 1e-300;   // OpenFOAM `vSmall`

#define FV_CELL_LOOP(g, c) for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < (g).N; c += gridDim.x * blockDim.x)

__device__ __forceinline__ bool fvInterior(const BoxGeom& g, int i, int j, int k)
{
    return i > 0 && i < g.nx - 1 && j > 0 && j < g.ny - 1 && k > 0 && k < g.nz - 1;
}

// ---------------------------------------------------------------------------------------------
// face-order conversion (OpenFOAM LDU face list <-> owner slots), fills
// ---------------------------------------------------------------------------------------------
__global__ void k_faces_to_slots(int n, const int* __restrict__ slotOf, const double* __restrict__ faces,
                                 double* __restrict__ slots)
{
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) slots[slotOf[f]] = faces[f];
}
__global__ void k_slots_to_faces(int n, const int* __restrict__ slotOf, const double* __restrict__ slots,
                                 double* __restrict__ faces)
{
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) faces[f] = slots[slotOf[f]];
}
__global__ void k_fill_d(double* __restrict__ a, size_t n, double v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = v;
}

// ---------------------------------------------------------------------------------------------
// createPhi.H: phi = linearInterpolate(U) & Sf
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK) k_create_phi(BoxGeom g, const double* __restrict__ U, double* __restrict__ phi)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
            if (v < nd - 1) {
                phi[d * g.N + c] = g.Sf[d] * fvLerp(g.w[d], U[3 * (size_t)c + d], U[3 * (size_t)(c + sd) + d]);
            } else {
                const int s = 2 * d + 1;
                const double ub = g.kindU[s] == FV_FIXED_VALUE ? g.valU[s][d] : (g.kindU[s] == FV_EMPTY ? 0.0 : U[3 * (size_t)c + d]);
                phi[d * g.N + c] = g.bSf[s] * ub;
            }
            if (v == 0) {
                const int s = 2 * d;
                const double ub = g.kindU[s] == FV_FIXED_VALUE ? g.valU[s][d] : (g.kindU[s] == FV_EMPTY ? 0.0 : U[3 * (size_t)c + d]);
                phi[fvSideSlot(g, s, c, i, j, k)] = g.bSf[s] * ub;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CourantNo.H: sumPhi = fvc::surfaceSum(mag(phi)); CoNum = 0.5 max(sumPhi/V) dt; mean = 0.5 sum(sumPhi)/sum(V) dt
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK) k_courant(BoxGeom g, const double* __restrict__ phi, double dt, FvRed red,
                                                 FvStepDev* __restrict__ out)
{
    double v[2] = {-1.7976931348623157e308, 0.0};
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        double s = 0.0;
        if (k > 0) s += fabs(phi[2 * g.N + c - g.sz]);
        if (j > 0) s += fabs(phi[1 * g.N + c - g.sy]);
        if (i > 0) s += fabs(phi[c - 1]);
        if (i < g.nx - 1) s += fabs(phi[c]);
        if (j < g.ny - 1) s += fabs(phi[g.N + c]);
        if (k < g.nz - 1) s += fabs(phi[2 * g.N + c]);
        if (!fvInterior(g, i, j, k)) {
            for (int q = 0; q < 6; ++q) {
                const int sd = g.seq[q];
                if (g.kindU[sd] == FV_EMPTY || !fvOnSide(g, sd, i, j, k)) continue;
                s += fabs(phi[fvSideSlot(g, sd, c, i, j, k)]);
            }
        }
        v[0] = fmax(v[0], s / g.V);
        v[1] += s;
    }
    const double sumV = g.sumV;
    fvGridReduce<2, true, BLK>(v, red, [=](const double* t) {
        out->CoNum = 0.5 * t[0] * dt;
        out->meanCoNum = 0.5 * (t[1] / sumV) * dt;
    });
}

// ---------------------------------------------------------------------------------------------
// fvc::grad (Gauss linear).  Vector -> tensor [N][9] (T_ij = Sf_i U_j), scalar -> vector [N][3].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fvPatchP(const BoxGeom& g, int s, double pc)
{
    return g.kindP[s] == FV_FIXED_VALUE ? g.valP[s] : pc;
}

__device__ __forceinline__ void fvGradP(const BoxGeom& g, const double* __restrict__ p, int c, int i, int j, int k, double* gp)
{
    const double pc = p[c];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
        double a = 0.0;
        if (v > 0) a -= g.Sf[d] * fvLerp(g.w[d], p[c - sd], pc);
        if (v < nd - 1) a += g.Sf[d] * fvLerp(g.w[d], pc, p[c + sd]);
        if (v == 0 || v == nd - 1) {
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if ((s >> 1) != d || g.kindP[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                a += g.bSf[s] * fvPatchP(g, s, pc);
            }
        }
        gp[d] = a / g.V;
    }
}

__global__ void __launch_bounds__(BLK) k_grad_scalar(BoxGeom g, const double* __restrict__ p, double* __restrict__ out)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        double gp[3];
        fvGradP(g, p, c, i, j, k, gp);
        out[3 * (size_t)c] = gp[0];
        out[3 * (size_t)c + 1] = gp[1];
        out[3 * (size_t)c + 2] = gp[2];
    }
}

__global__ void __launch_bounds__(BLK) k_grad_vector(BoxGeom g, const double* __restrict__ U, double* __restrict__ out)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double uc[3] = {U[3 * (size_t)c], U[3 * (size_t)c + 1], U[3 * (size_t)c + 2]};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
            double a[3] = {0.0, 0.0, 0.0};
            if (v > 0) {
                const double* un = U + 3 * (size_t)(c - sd);
#pragma unroll
                for (int m = 0; m < 3; ++m) a[m] -= g.Sf[d] * fvLerp(g.w[d], un[m], uc[m]);
            }
            if (v < nd - 1) {
                const double* un = U + 3 * (size_t)(c + sd);
#pragma unroll
                for (int m = 0; m < 3; ++m) a[m] += g.Sf[d] * fvLerp(g.w[d], uc[m], un[m]);
            }
            if (v == 0 || v == nd - 1) {
                for (int q = 0; q < 6; ++q) {
                    const int s = g.seq[q];
                    if ((s >> 1) != d || g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
#pragma unroll
                    for (int m = 0; m < 3; ++m) a[m] += g.bSf[s] * (g.kindU[s] == FV_FIXED_VALUE ? g.valU[s][m] : uc[m]);
                }
            }
#pragma unroll
            for (int m = 0; m < 3; ++m) out[9 * (size_t)c + 3 * d + m] = a[m] / g.V;
        }
    }
}

// fvc::div(phi) of one cell (surfaceIntegrate), used by the pEqn source, continuityErrs and the parity hook
__device__ __forceinline__ double fvDivCell(const BoxGeom& g, const double* __restrict__ phi, int c, int i, int j, int k)
{
    double d = 0.0;
    if (k > 0) d -= phi[2 * g.N + c - g.sz];
    if (j > 0) d -= phi[g.N + c - g.sy];
    if (i > 0) d -= phi[c - 1];
    if (i < g.nx - 1) d += phi[c];
    if (j < g.ny - 1) d += phi[g.N + c];
    if (k < g.nz - 1) d += phi[2 * g.N + c];
    if (!fvInterior(g, i, j, k)) {
        for (int q = 0; q < 6; ++q) {
            const int s = g.seq[q];
            if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
            d += phi[fvSideSlot(g, s, c, i, j, k)];
        }
    }
    return d / g.V;
}

__global__ void __launch_bounds__(BLK) k_div_flux(BoxGeom g, const double* __restrict__ phi, double* __restrict__ out)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        out[c] = fvDivCell(g, phi, c, i, j, k);
    }
}

// ---------------------------------------------------------------------------------------------
// UEqn = fvm::ddt(U) + fvm::div(phi,U) - fvm::laplacian(nu,U) == uSource           icoFoamYade.C:79-85
// per cell: diag, lower/upper of its own faces, source; also rAU = 1/A() and the per-component solve
// diagonals (diag + internalCoeffs), which depend on the matrix only.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK)
k_assemble_U(BoxGeom g, double nu, double rDeltaT, const double* __restrict__ phi, const double* __restrict__ U0,
             const double* __restrict__ uSource, double* __restrict__ diagU, double* __restrict__ loU,
             double* __restrict__ upU, double* __restrict__ srcU, double* __restrict__ dgU, double* __restrict__ rAU)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const int N = g.N;
        const double diagD = rDeltaT * g.V;
        double diagC = 0.0, diagL = 0.0;
        // faces on which c is the neighbour: Diag[u] -= Upper
#pragma unroll
        for (int d = 2; d >= 0; --d) {
            const int v = fvIdx(d, i, j, k), sd = fvStride(g, d);
            if (v > 0) {
                const double ph = phi[d * N + c - sd];
                const double lowerC = -g.w[d] * ph;
                const double upperC = lowerC + ph;
                diagC -= upperC;
                diagL -= g.dc[d] * (nu * g.magSf[d]);
            }
        }
        // faces c owns: Diag[l] -= Lower
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d);
            double lo = 0.0, up = 0.0;
            if (v < nd - 1) {
                const double ph = phi[d * N + c];
                const double lowerC = -g.w[d] * ph;
                const double upperC = lowerC + ph;
                const double upperL = g.dc[d] * (nu * g.magSf[d]);
                diagC -= lowerC;
                diagL -= upperL;
                lo = lowerC - upperL;
                up = upperC - upperL;
            }
            loU[d * N + c] = lo;
            upU[d * N + c] = up;
        }
        const double diag = (diagD + diagC) - diagL;
        diagU[c] = diag;
        double D = diag, dg[3] = {diag, diag, diag};
        if (!fvInterior(g, i, j, k)) {
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                const double phib = phi[fvSideSlot(g, s, c, i, j, k)];
                double ic[3], bc;
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    fvBCoefU(g, s, phib, nu, m, ic[m], bc);
                    dg[m] += ic[m];                                   // addBoundaryDiag
                }
                D += (ic[0] + ic[1] + ic[2]) / 3.0;                   // addCmptAvBoundaryDiag
            }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            dgU[m * (size_t)N + c] = dg[m];
            double s = rDeltaT * U0[3 * (size_t)c + m] * g.V;        // fvm::ddt source
            s += g.V * uSource[3 * (size_t)c + m];                    // == uSource
            srcU[3 * (size_t)c + m] = s;
        }
        rAU[c] = 1.0 / (D / g.V);                                     // 1.0/UEqn.A()
    }
}

// source of solve(UEqn == -fvc::grad(p)) with the boundary source, split by component (SoA); psi = U
__global__ void __launch_bounds__(BLK)
k_usolve_setup(BoxGeom g, double nu, const double* __restrict__ phi, const double* __restrict__ srcU,
               const double* __restrict__ gradP, const double* __restrict__ U, double* __restrict__ bU,
               double* __restrict__ psiU)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        double b[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) b[m] = srcU[3 * (size_t)c + m] + g.V * (-gradP[3 * (size_t)c + m]);
        if (!fvInterior(g, i, j, k)) {
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                const double phib = phi[fvSideSlot(g, s, c, i, j, k)];
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    double ic, bc;
                    fvBCoefU(g, s, phib, nu, m, ic, bc);
                    b[m] += bc;                                       // addBoundarySource
                }
            }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            bU[m * (size_t)g.N + c] = b[m];
            psiU[m * (size_t)g.N + c] = U[3 * (size_t)c + m];
        }
    }
}

__global__ void __launch_bounds__(BLK) k_store_component(int N, const double* __restrict__ psi, int m, double* __restrict__ U)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) U[3 * (size_t)c + m] = psi[c];
}

// ---------------------------------------------------------------------------------------------
// lduMatrix kernels on owner-slot coefficients (lo/up [3N]; for a symmetric matrix lo == up)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fvAmulCell(const BoxGeom& g, const double* __restrict__ dg, const double* __restrict__ lo,
                                             const double* __restrict__ up, const double* __restrict__ x, int c, int i,
                                             int j, int k)
{
    const int N = g.N;
    double a = dg[c] * x[c];
    if (k > 0) a += lo[2 * N + c - g.sz] * x[c - g.sz];
    if (j > 0) a += lo[N + c - g.sy] * x[c - g.sy];
    if (i > 0) a += lo[c - 1] * x[c - 1];
    if (i < g.nx - 1) a += up[c] * x[c + 1];
    if (j < g.ny - 1) a += up[N + c] * x[c + g.sy];
    if (k < g.nz - 1) a += up[2 * N + c] * x[c + g.sz];
    return a;
}
__device__ __forceinline__ double fvSumACell(const BoxGeom& g, const double* __restrict__ dg, const double* __restrict__ lo,
                                             const double* __restrict__ up, int c, int i, int j, int k)
{
    const int N = g.N;
    double a = dg[c];
    if (k > 0) a += lo[2 * N + c - g.sz];
    if (j > 0) a += lo[N + c - g.sy];
    if (i > 0) a += lo[c - 1];
    if (i < g.nx - 1) a += up[c];
    if (j < g.ny - 1) a += up[N + c];
    if (k < g.nz - 1) a += up[2 * N + c];
    return a;
}

__global__ void k_solve_begin(FvSolveDev* st, double tol, double relTol, int maxIter, int precond)
{
    st->tol = tol; st->relTol = relTol; st->maxIter = maxIter; st->precond = precond;
    st->avg = 0; st->normFactor = 0; st->initRes = 0; st->finalRes = 0;
    st->wArA = 1e20; st->wArAold = 1e20; st->wApA = 0; st->alpha = 0; st->beta = 0;
    st->nIter = 0; st->done = 0; st->singular = 0;
}

__device__ __forceinline__ bool fvConverged(const FvSolveDev* st)
{
    return st->finalRes < st->tol || (st->relTol > 1e-20 && st->finalRes < st->relTol * st->initRes);
}

// gAverage(psi)
__global__ void __launch_bounds__(BLK) k_avg(int N, const double* __restrict__ psi, FvRed red, FvSolveDev* st)
{
    double v[1] = {0.0};
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) v[0] += psi[c];
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) { st->avg = t[0] / N; });
}

// rA = source - A psi;  normFactor = sum(|Apsi - sumA avg| + |source - sumA avg|) + 1e-20;
// initialResidual = sum|rA| / normFactor                              [OF-6 lduMatrix::solver::normFactor]
__global__ void __launch_bounds__(BLK)
k_solve_init(BoxGeom g, const double* __restrict__ dg, const double* __restrict__ lo, const double* __restrict__ up,
             const double* __restrict__ b, const double* __restrict__ psi, double* __restrict__ rA, FvRed red,
             FvSolveDev* st)
{
    double v[2] = {0.0, 0.0};
    const double avg = st->avg;
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double Apsi = fvAmulCell(g, dg, lo, up, psi, c, i, j, k);
        const double t = fvSumACell(g, dg, lo, up, c, i, j, k) * avg;
        const double r = b[c] - Apsi;
        if (rA) rA[c] = r;
        v[0] += fabs(Apsi - t) + fabs(b[c] - t);
        v[1] += fabs(r);
    }
    fvGridReduce<2, false, BLK>(v, red, [=](const double* t) {
        st->normFactor = t[0] + 1e-20;
        st->initRes = t[1] / st->normFactor;
        st->finalRes = st->initRes;
        st->done = fvConverged(st) ? 1 : 0;
    });
}

// lduMatrix::residual + gSumMag (smoothSolver's convergence test after each sweep)
__global__ void __launch_bounds__(BLK)
k_residual(BoxGeom g, const double* __restrict__ dg, const double* __restrict__ lo, const double* __restrict__ up,
           const double* __restrict__ b, const double* __restrict__ psi, FvRed red, FvSolveDev* st)
{
    double v[1] = {0.0};
    const int N = g.N;
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        double r = b[c] - dg[c] * psi[c];
        if (k > 0) r -= lo[2 * N + c - g.sz] * psi[c - g.sz];
        if (j > 0) r -= lo[N + c - g.sy] * psi[c - g.sy];
        if (i > 0) r -= lo[c - 1] * psi[c - 1];
        if (i < g.nx - 1) r -= up[c] * psi[c + 1];
        if (j < g.ny - 1) r -= up[N + c] * psi[c + g.sy];
        if (k < g.nz - 1) r -= up[2 * N + c] * psi[c + g.sz];
        v[0] += fabs(r);
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) {
        st->finalRes = t[0] / st->normFactor;
        st->nIter += 1;                                               // nSweeps = 1
        st->done = (!(st->nIter < st->maxIter) || fvConverged(st)) ? 1 : 0;
    });
}

// ---------------------------------------------------------------------------------------------
// wavefront engine: cells of the hyperplane i+j+k = s are independent in every sequential LDU
// recurrence; one cooperative kernel walks the planes with a grid barrier between them.
// ---------------------------------------------------------------------------------------------
struct OpDicD {            // DIC calcReciprocalD, before the final 1/x: rD[u] -= upper^2 / rD[l]
    const double* dg; const double* up; double* D;
    __device__ __forceinline__ void cell(const BoxGeom& g, int c, int i, int j, int k, double&) const
    {
        const int N = g.N;
        double r = dg[c];
        if (k > 0) { const double u = up[2 * N + c - g.sz]; r -= u * u / D[c - g.sz]; }
        if (j > 0) { const double u = up[N + c - g.sy]; r -= u * u / D[c - g.sy]; }
        if (i > 0) { const double u = up[c - 1]; r -= u * u / D[c - 1]; }
        D[c] = r;
    }
};
struct OpDicFwd {          // wA = rD rA;  wA[u] -= rD[u] upper wA[l]   (faces ascending)
    const double* rD; const double* up; const double* rA; double* wA;
    __device__ __forceinline__ void cell(const BoxGeom& g, int c, int i, int j, int k, double&) const
    {
        const int N = g.N;
        const double rd = rD[c];
        double w = rd * rA[c];
        if (k > 0) w -= rd * up[2 * N + c - g.sz] * wA[c - g.sz];
        if (j > 0) w -= rd * up[N + c - g.sy] * wA[c - g.sy];
        if (i > 0) w -= rd * up[c - 1] * wA[c - 1];
        wA[c] = w;
    }
};
struct OpDicBwd {          // wA[l] -= rD[l] upper wA[u]   (faces descending); accumulates wA.rA
    const double* rD; const double* up; const double* rA; double* wA;
    __device__ __forceinline__ void cell(const BoxGeom& g, int c, int i, int j, int k, double& acc) const
    {
        const int N = g.N;
        const double rd = rD[c];
        double w = wA[c];
        if (k < g.nz - 1) w -= rd * up[2 * N + c] * wA[c + g.sz];
        if (j < g.ny - 1) w -= rd * up[N + c] * wA[c + g.sy];
        if (i < g.nx - 1) w -= rd * up[c] * wA[c + 1];
        wA[c] = w;
        acc += w * rA[c];
    }
};
struct OpGsFwd {           // symGaussSeidel forward sweep; leaves bPrime (source minus the lower-side products)
    const double* dg; const double* lo; const double* up; const double* b; double* psi; double* bPrime;
    __device__ __forceinline__ void cell(const BoxGeom& g, int c, int i, int j, int k, double&) const
    {
        const int N = g.N;
        double bp = b[c];
        if (k > 0) bp -= lo[2 * N + c - g.sz] * psi[c - g.sz];
        if (j > 0) bp -= lo[N + c - g.sy] * psi[c - g.sy];
        if (i > 0) bp -= lo[c - 1] * psi[c - 1];
        bPrime[c] = bp;
        double x = bp;
        if (i < g.nx - 1) x -= up[c] * psi[c + 1];
        if (j < g.ny - 1) x -= up[N + c] * psi[c + g.sy];
        if (k < g.nz - 1) x -= up[2 * N + c] * psi[c + g.sz];
        psi[c] = x / dg[c];
    }
};
struct OpGsBwd {           // symGaussSeidel backward sweep
    const double* dg; const double* up; const double* bPrime; double* psi;
    __device__ __forceinline__ void cell(const BoxGeom& g, int c, int i, int j, int k, double&) const
    {
        const int N = g.N;
        double x = bPrime[c];
        if (i < g.nx - 1) x -= up[c] * psi[c + 1];
        if (j < g.ny - 1) x -= up[N + c] * psi[c + g.sy];
        if (k < g.nz - 1) x -= up[2 * N + c] * psi[c + g.sz];
        psi[c] = x / dg[c];
    }
};

// DOT: after the sweep, reduce the per-thread accumulators into st->wArA (PCG search-direction update)
template <class Op, bool REV, bool DOT>
__global__ void __launch_bounds__(BLK) k_wave(BoxGeom g, Op op, FvRed red, FvSolveDev* st)
{
    if (st && st->done) return;                 // uniform across the grid: nobody reaches the barrier
    cg::grid_group grid = cg::this_grid();
    const int nJK = g.ny * g.nz, S = g.nx + g.ny + g.nz - 2;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int step = 0; step < S; ++step) {
        const int s = REV ? S - 1 - step : step;
        for (int jk = tid; jk < nJK; jk += nth) {
            const int j = jk % g.ny, k = jk / g.ny, i = s - j - k;
            if (i >= 0 && i < g.nx) op.cell(g, i + g.nx * (j + g.ny * k), i, j, k, acc);
        }
        grid.sync();
    }
    if (DOT) {
        double v[1] = {acc};
        fvGridReduce<1, false, BLK>(v, red, [=](const double* t) {
            st->wArAold = st->wArA;
            st->wArA = t[0];
            st->beta = st->wArA / st->wArAold;
        });
    }
}

__global__ void __launch_bounds__(BLK) k_recip(int N, const double* __restrict__ a, double* __restrict__ out, const FvSolveDev* st)
{
    if (st && st->done) return;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) out[c] = 1.0 / a[c];
}

// diagonal / no preconditioner: wA = rD rA (or rA) with the wA.rA dot
__global__ void __launch_bounds__(BLK)
k_precond_diag(int N, const double* __restrict__ rD, const double* __restrict__ rA, double* __restrict__ wA, FvRed red,
               FvSolveDev* st)
{
    if (st->done) return;
    double v[1] = {0.0};
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
        const double w = rD ? rD[c] * rA[c] : rA[c];
        wA[c] = w;
        v[0] += w * rA[c];
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) {
        st->wArAold = st->wArA;
        st->wArA = t[0];
        st->beta = st->wArA / st->wArAold;
    });
}

// pA = wA (first iteration) | wA + beta pA
__global__ void __launch_bounds__(BLK) k_pcg_dir(int N, const double* __restrict__ wA, double* __restrict__ pA, const FvSolveDev* st)
{
    if (st->done) return;
    const bool first = st->nIter == 0;
    const double beta = st->beta;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x)
        pA[c] = first ? wA[c] : wA[c] + beta * pA[c];
}

// wA = A pA; wApA = wA.pA; alpha = wArA/wApA (with the singularity test of PCG.C)
__global__ void __launch_bounds__(BLK)
k_pcg_amul(BoxGeom g, const double* __restrict__ dg, const double* __restrict__ up, const double* __restrict__ pA,
           double* __restrict__ wA, FvRed red, FvSolveDev* st)
{
    if (st->done) return;
    double v[1] = {0.0};
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double a = fvAmulCell(g, dg, up, up, pA, c, i, j, k);
        wA[c] = a;
        v[0] += a * pA[c];
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) {
        st->wApA = t[0];
        if (fabs(t[0]) / st->normFactor < FV_VSMALL) { st->singular = 1; st->done = 1; }
        else st->alpha = st->wArA / t[0];
    });
}

// psi += alpha pA; rA -= alpha wA; finalResidual = sum|rA|/normFactor; PCG.C's loop condition
__global__ void __launch_bounds__(BLK)
k_pcg_update(int N, const double* __restrict__ pA, const double* __restrict__ wA, double* __restrict__ psi,
             double* __restrict__ rA, FvRed red, FvSolveDev* st)
{
    if (st->done) return;
    const double alpha = st->alpha;
    double v[1] = {0.0};
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
        psi[c] += alpha * pA[c];
        const double r = rA[c] - alpha * wA[c];
        rA[c] = r;
        v[0] += fabs(r);
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) {
        st->finalRes = t[0] / st->normFactor;
        const bool cont = st->nIter < st->maxIter;                    // nIterations++ < maxIter_
        st->nIter += 1;
        if (!cont || fvConverged(st)) st->done = 1;
    });
}

// ---------------------------------------------------------------------------------------------
// PISO corrector kernels
// ---------------------------------------------------------------------------------------------
// HbyA = rAU*UEqn.H()                                                               icoFoamYade.C:100
__global__ void __launch_bounds__(BLK)
k_HbyA(BoxGeom g, double nu, const double* __restrict__ phi, const double* __restrict__ loU, const double* __restrict__ upU,
       const double* __restrict__ srcU, const double* __restrict__ U, const double* __restrict__ rAU,
       double* __restrict__ HbyA)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const int N = g.N;
        const bool inner = fvInterior(g, i, j, k);
        double H[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            // (cmptAv(internalCoeffs) - internalCoeffs.component(m)) * psi
            double bd = 0.0;
            if (!inner) {
                for (int q = 0; q < 6; ++q) {
                    const int s = g.seq[q];
                    if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                    double ic, bc;
                    fvBCoefU(g, s, phi[fvSideSlot(g, s, c, i, j, k)], nu, m, ic, bc);
                    bd += ic;
                }
                bd = -bd;
                for (int q = 0; q < 6; ++q) {
                    const int s = g.seq[q];
                    if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                    const double phib = phi[fvSideSlot(g, s, c, i, j, k)];
                    double ic0, ic1, ic2, bc;
                    fvBCoefU(g, s, phib, nu, 0, ic0, bc);
                    fvBCoefU(g, s, phib, nu, 1, ic1, bc);
                    fvBCoefU(g, s, phib, nu, 2, ic2, bc);
                    bd += (ic0 + ic1 + ic2) / 3.0;
                }
            }
            double h = bd * U[3 * (size_t)c + m];
            // lduMatrix::H
            double hl = 0.0;
            if (k > 0) hl -= loU[2 * N + c - g.sz] * U[3 * (size_t)(c - g.sz) + m];
            if (j > 0) hl -= loU[N + c - g.sy] * U[3 * (size_t)(c - g.sy) + m];
            if (i > 0) hl -= loU[c - 1] * U[3 * (size_t)(c - 1) + m];
            if (i < g.nx - 1) hl -= upU[c] * U[3 * (size_t)(c + 1) + m];
            if (j < g.ny - 1) hl -= upU[N + c] * U[3 * (size_t)(c + g.sy) + m];
            if (k < g.nz - 1) hl -= upU[2 * N + c] * U[3 * (size_t)(c + g.sz) + m];
            h += hl + srcU[3 * (size_t)c + m];
            if (!inner) {
                for (int q = 0; q < 6; ++q) {
                    const int s = g.seq[q];
                    if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                    double ic, bc;
                    fvBCoefU(g, s, phi[fvSideSlot(g, s, c, i, j, k)], nu, m, ic, bc);
                    h += bc;                                          // addBoundarySource
                }
            }
            h /= g.V;
            H[m] = g.valid[m] ? h : 0.0;
        }
        const double r = rAU[c];
        HbyA[3 * (size_t)c] = r * H[0];
        HbyA[3 * (size_t)c + 1] = r * H[1];
        HbyA[3 * (size_t)c + 2] = r * H[2];
    }
}

// boundary face of phiHbyA (constrainHbyA + ddtCorr on the patch)
__device__ __forceinline__ double fvPhiHbyAB(const BoxGeom& g, int s, int d, double rDeltaT, double hbyaC, double u0C,
                                             double rAUc, double phi0b)
{
    if (g.kindU[s] == FV_EMPTY) return 0.0;
    const bool fixed = g.kindU[s] == FV_FIXED_VALUE;
    const double hb = fixed ? g.valU[s][d] : hbyaC;
    const double u0b = fixed ? g.valU[s][d] : u0C;
    const double flux = g.bSf[s] * hb;
    const double phiCorr = phi0b - g.bSf[s] * u0b;
    double coeff = 1.0 - fmin(fabs(phiCorr) / (fabs(phi0b) + FV_SMALL), 1.0);
    if (fixed) coeff = 0.0;
    return flux + rAUc * ((coeff * rDeltaT) * phiCorr);
}

// phiHbyA = fvc::flux(HbyA) + fvc::interpolate(rAU)*fvc::ddtCorr(U, phi)              icoFoamYade.C:101-106
// and the pEqn off-diagonals  upper = deltaCoeffs*(interpolate(rAU)*magSf)            icoFoamYade.C:118-121
__global__ void __launch_bounds__(BLK)
k_phiHbyA(BoxGeom g, double rDeltaT, const double* __restrict__ HbyA, const double* __restrict__ U0,
          const double* __restrict__ phi0, const double* __restrict__ rAU, double* __restrict__ phiHbyA,
          double* __restrict__ upP)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const int N = g.N;
        const double rc = rAU[c];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
            const double hc = HbyA[3 * (size_t)c + d], u0c = U0[3 * (size_t)c + d];
            if (v < nd - 1) {
                const int n = c + sd;
                const double flux = g.Sf[d] * fvLerp(g.w[d], hc, HbyA[3 * (size_t)n + d]);
                const double u0f = g.Sf[d] * fvLerp(g.w[d], u0c, U0[3 * (size_t)n + d]);
                const double ph0 = phi0[d * N + c];
                const double phiCorr = ph0 - u0f;
                const double coeff = 1.0 - fmin(fabs(phiCorr) / (fabs(ph0) + FV_SMALL), 1.0);
                const double rf = fvLerp(g.w[d], rc, rAU[n]);
                phiHbyA[d * N + c] = flux + rf * ((coeff * rDeltaT) * phiCorr);
                upP[d * N + c] = g.dc[d] * (rf * g.magSf[d]);
            } else {
                phiHbyA[d * N + c] = fvPhiHbyAB(g, 2 * d + 1, d, rDeltaT, hc, u0c, rc, phi0[d * N + c]);
                upP[d * N + c] = 0.0;
            }
            if (v == 0) {
                const int sl = fvSideSlot(g, 2 * d, c, i, j, k);
                phiHbyA[sl] = fvPhiHbyAB(g, 2 * d, d, rDeltaT, hc, u0c, rc, phi0[sl]);
            }
        }
    }
}

// adjustPhi(phiHbyA, U, p): sums over the boundary faces, then the outflow scaling          icoFoamYade.C:108
__global__ void __launch_bounds__(BLK) k_adjust_sum(BoxGeom g, const double* __restrict__ phiHbyA, FvRed red, FvStepDev* out)
{
    double v[4] = {0.0, 0.0, 0.0, 0.0};       // massIn fixedMassOut adjustableMassOut sum|phi_b|
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        if (fvInterior(g, i, j, k)) continue;
        for (int q = 0; q < 6; ++q) {
            const int s = g.seq[q];
            if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
            const double ph = phiHbyA[fvSideSlot(g, s, c, i, j, k)];
            if (ph < 0) v[0] -= ph;
            else if (g.kindU[s] == FV_FIXED_VALUE) v[1] += ph;
            else v[2] += ph;
            v[3] += fabs(ph);
        }
    }
    fvGridReduce<4, false, BLK>(v, red, [=](const double* t) {
        const double totalFlux = FV_VSMALL + t[3];
        double massCorr = 1.0;
        int fail = 0;
        const double magAdj = fabs(t[2]);
        if (magAdj > FV_VSMALL && magAdj / totalFlux > FV_SMALL) massCorr = (t[0] - t[1]) / t[2];
        else if (fabs(t[1] - t[0]) / totalFlux > 1e-8) fail = 1;
        out->massIn = t[0]; out->fixedMassOut = t[1]; out->adjustableMassOut = t[2]; out->totalFlux = totalFlux;
        out->massCorr = massCorr;
        out->adjustFail = fail;
    });
}
__global__ void __launch_bounds__(BLK) k_adjust_scale(BoxGeom g, double* __restrict__ phiHbyA, const FvStepDev* st)
{
    const double massCorr = st->massCorr;
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        if (fvInterior(g, i, j, k)) continue;
        for (int q = 0; q < 6; ++q) {
            const int s = g.seq[q];
            if (g.kindU[s] != FV_ZERO_GRADIENT || !fvOnSide(g, s, i, j, k)) continue;
            const int sl = fvSideSlot(g, s, c, i, j, k);
            if (phiHbyA[sl] > 0.0) phiHbyA[sl] *= massCorr;
        }
    }
}

// pEqn: fvm::laplacian(rAU, p) == fvc::div(phiHbyA); setReference; solve's diag/source with the boundary
// coefficients                                                                       icoFoamYade.C:118-125
__global__ void __launch_bounds__(BLK)
k_pEqn(BoxGeom g, int pRefCell, double pRefValue, const double* __restrict__ upP, const double* __restrict__ phiHbyA,
       const double* __restrict__ rAU, double* __restrict__ dgP, double* __restrict__ bP)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const int N = g.N;
        double diag = 0.0;
        if (k > 0) diag -= upP[2 * N + c - g.sz];
        if (j > 0) diag -= upP[N + c - g.sy];
        if (i > 0) diag -= upP[c - 1];
        if (i < g.nx - 1) diag -= upP[c];
        if (j < g.ny - 1) diag -= upP[N + c];
        if (k < g.nz - 1) diag -= upP[2 * N + c];
        double src = 0.0;
        src += g.V * fvDivCell(g, phiHbyA, c, i, j, k);
        if (g.pNeedRef && c == pRefCell) {
            src += diag * pRefValue;
            diag += diag;
        }
        if (!fvInterior(g, i, j, k)) {
            const double rc = rAU[c];
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindP[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                double ic, bc;
                fvBCoefP(g, s, rc, ic, bc);
                diag += ic;
                src += bc;
            }
        }
        dgP[c] = diag;
        bP[c] = src;
    }
}

// phi = phiHbyA - pEqn.flux()                                                         icoFoamYade.C:127-130
__global__ void __launch_bounds__(BLK)
k_flux_update(BoxGeom g, const double* __restrict__ upP, const double* __restrict__ phiHbyA, const double* __restrict__ p,
              const double* __restrict__ rAU, double* __restrict__ phi)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const int N = g.N;
        const double pc = p[c], rc = rAU[c];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
            if (v < nd - 1) {
                const double u = upP[d * N + c];
                phi[d * N + c] = phiHbyA[d * N + c] - (u * p[c + sd] - u * pc);
            } else {
                double ic, bc;
                fvBCoefP(g, 2 * d + 1, rc, ic, bc);
                phi[d * N + c] = phiHbyA[d * N + c] - (ic * pc - bc);
            }
            if (v == 0) {
                double ic, bc;
                fvBCoefP(g, 2 * d, rc, ic, bc);
                const int sl = fvSideSlot(g, 2 * d, c, i, j, k);
                phi[sl] = phiHbyA[sl] - (ic * pc - bc);
            }
        }
    }
}

// continuityErrs.H + U = HbyA - rAU*fvc::grad(p)                                      icoFoamYade.C:134-137
__global__ void __launch_bounds__(BLK)
k_correct_U(BoxGeom g, double dt, int corr, const double* __restrict__ phi, const double* __restrict__ p,
            const double* __restrict__ HbyA, const double* __restrict__ rAU, double* __restrict__ U,
            double* __restrict__ gradPOut, FvRed red, FvStepDev* out)
{
    double v[2] = {0.0, 0.0};
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double dv = fvDivCell(g, phi, c, i, j, k);
        v[0] += fabs(dv) * g.V;
        v[1] += dv * g.V;
        double gp[3];
        fvGradP(g, p, c, i, j, k, gp);
        const double r = rAU[c];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            gradPOut[3 * (size_t)c + m] = gp[m];
            U[3 * (size_t)c + m] = HbyA[3 * (size_t)c + m] - r * gp[m];
        }
    }
    const double sumV = g.sumV;
    fvGridReduce<2, false, BLK>(v, red, [=](const double* t) {
        out->sumLocal = dt * (t[0] / sumV);
        out->global = dt * (t[1] / sumV);
        if (corr < 8) { out->corrSumLocal[corr] = out->sumLocal; out->corrGlobal[corr] = out->global; }
    });
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define FV_LAUNCH(kernel, grid, ...)                                                        \
    do {                                                                                    \
        kernel<<<(grid), BLK, 0, h->stream>>>(__VA_ARGS__);                                 \
        FY_CHECK_LAUNCH();                                                                  \
    } while (0)

template <class Op, bool REV, bool DOT>
int launchWave(fy_ctx* h, FvState* s, Op op, FvSolveDev* st)
{
    BoxGeom g = s->g;
    FvRed red = s->red;
    void* args[] = {&g, &op, &red, &st};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_wave<Op, REV, DOT>, dim3(s->waveGrid), dim3(BLK), args, 0,
                                                h->stream);
    h->launches++;
    if (e != cudaSuccess) {
        h->err = std::string("cooperative launch: ") + cudaGetErrorString(e);
        return FY_ERR_CUDA;
    }
    return FY_OK;
}

int readSolve(fy_ctx* h, FvState* s)
{
    FY_CUDA(cudaMemcpyAsync(s->hSolve, s->dSolve, sizeof(FvSolveDev), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}

}  // namespace

// PCG on owner-slot coefficients (device pointers).  Iteration kernels are queued in batches and test the
// device-side `done` flag themselves; the host looks at the state once per batch.
int fvPcgSolve(fy_ctx* h, FvState* s, const double* dg, const double* up, const double* b, double* psi, double tol,
               double relTol, int maxIter, int precond, fy_solver_perf* perf)
{
    const BoxGeom& g = s->g;
    const int N = g.N, G = s->cellGrid;
    int rc;
    k_solve_begin<<<1, 1, 0, h->stream>>>(s->dSolve, tol, relTol, maxIter, precond);
    FY_CHECK_LAUNCH();
    FV_LAUNCH(k_avg, G, N, psi, s->red, s->dSolve);
    FV_LAUNCH(k_solve_init, G, g, dg, up, up, b, psi, s->rA, s->red, s->dSolve);
    if (precond == FV_PRECOND_DIC) {
        if ((rc = launchWave<OpDicD, false, false>(h, s, OpDicD{dg, up, s->rD}, s->dSolve))) return rc;
        FV_LAUNCH(k_recip, G, N, s->rD, s->rD, s->dSolve);
    } else if (precond == FV_PRECOND_DIAGONAL) {
        FV_LAUNCH(k_recip, G, N, dg, s->rD, s->dSolve);
    }
    int queued = 0;
    bool sampled = false;
    const bool prof = h->profiling && s->pev[0];
    for (;;) {
        if ((rc = readSolve(h, s))) return rc;
        if (sampled) {
            for (int q = 0; q < 5; ++q) {
                float ms = 0;
                cudaEventElapsedTime(&ms, s->pev[q], s->pev[q + 1]);
                s->kernelMs[q] += ms;
            }
            s->kernelSamples++;
            sampled = false;
        }
        if (s->hSolve->done) break;
        int batch = s->pcgBatch;
        if (queued >= 4 * batch) batch *= 2;
        for (int it = 0; it < batch; ++it) {
            const bool ev = prof && it == 0;
            if (ev) cudaEventRecord(s->pev[0], h->stream);
            if (precond == FV_PRECOND_DIC) {
                if ((rc = launchWave<OpDicFwd, false, false>(h, s, OpDicFwd{s->rD, up, s->rA, s->wA}, s->dSolve))) return rc;
                if (ev) cudaEventRecord(s->pev[1], h->stream);
                if ((rc = launchWave<OpDicBwd, true, true>(h, s, OpDicBwd{s->rD, up, s->rA, s->wA}, s->dSolve))) return rc;
            } else {
                if (ev) cudaEventRecord(s->pev[1], h->stream);
                FV_LAUNCH(k_precond_diag, G, N, precond == FV_PRECOND_DIAGONAL ? s->rD : (const double*)nullptr, s->rA, s->wA,
                          s->red, s->dSolve);
            }
            if (ev) cudaEventRecord(s->pev[2], h->stream);
            FV_LAUNCH(k_pcg_dir, G, N, s->wA, s->pA, s->dSolve);
            if (ev) cudaEventRecord(s->pev[3], h->stream);
            FV_LAUNCH(k_pcg_amul, G, g, dg, up, s->pA, s->wA, s->red, s->dSolve);
            if (ev) cudaEventRecord(s->pev[4], h->stream);
            FV_LAUNCH(k_pcg_update, G, N, s->pA, s->wA, psi, s->rA, s->red, s->dSolve);
            if (ev) { cudaEventRecord(s->pev[5], h->stream); sampled = true; }
        }
        queued += batch;
    }
    s->pcgIterations += s->hSolve->nIter;
    if (perf) {
        perf->initialResidual = s->hSolve->initRes;
        perf->finalResidual = s->hSolve->finalRes;
        perf->nIterations = s->hSolve->nIter;
    }
    return FY_OK;
}

// smoothSolver + symGaussSeidel (nSweeps 1) on owner-slot coefficients (device pointers)
int fvSmoothSolve(fy_ctx* h, FvState* s, const double* dg, const double* lo, const double* up, const double* b,
                  double* psi, double tol, double relTol, int maxIter, fy_solver_perf* perf)
{
    const BoxGeom& g = s->g;
    const int N = g.N, G = s->cellGrid;
    int rc;
    k_solve_begin<<<1, 1, 0, h->stream>>>(s->dSolve, tol, relTol, maxIter, 0);
    FY_CHECK_LAUNCH();
    FV_LAUNCH(k_avg, G, N, psi, s->red, s->dSolve);
    FV_LAUNCH(k_solve_init, G, g, dg, lo, up, b, psi, (double*)nullptr, s->red, s->dSolve);
    for (;;) {
        if ((rc = readSolve(h, s))) return rc;
        if (s->hSolve->done) break;
        if ((rc = launchWave<OpGsFwd, false, false>(h, s, OpGsFwd{dg, lo, up, b, psi, s->bPrime}, s->dSolve))) return rc;
        if ((rc = launchWave<OpGsBwd, true, false>(h, s, OpGsBwd{dg, up, s->bPrime, psi}, s->dSolve))) return rc;
        FV_LAUNCH(k_residual, G, g, dg, lo, up, b, psi, s->red, s->dSolve);
    }
    if (perf) {
        perf->initialResidual = s->hSolve->initRes;
        perf->finalResidual = s->hSolve->finalRes;
        perf->nIterations = s->hSolve->nIter;
    }
    return FY_OK;
}

int fvDicPrecondition(fy_ctx* h, FvState* s, const double* dg, const double* up, const double* rA, double* wA)
{
    int rc;
    if ((rc = launchWave<OpDicD, false, false>(h, s, OpDicD{dg, up, s->rD}, nullptr))) return rc;
    FV_LAUNCH(k_recip, s->cellGrid, s->g.N, s->rD, s->rD, (const FvSolveDev*)nullptr);
    if ((rc = launchWave<OpDicFwd, false, false>(h, s, OpDicFwd{s->rD, up, rA, wA}, nullptr))) return rc;
    if ((rc = launchWave<OpDicBwd, true, false>(h, s, OpDicBwd{s->rD, up, rA, wA}, nullptr))) return rc;
    return FY_OK;
}

int fvCreatePhi(fy_ctx* h, FvState* s)
{
    FV_LAUNCH(k_create_phi, s->cellGrid, s->g, h->dField[FY_F_U], s->phi);
    return FY_OK;
}

int fvGradVector(fy_ctx* h, FvState* s, const double* dU, double* dOut)
{
    FV_LAUNCH(k_grad_vector, s->cellGrid, s->g, dU, dOut);
    return FY_OK;
}
int fvGradScalar(fy_ctx* h, FvState* s, const double* dP, double* dOut)
{
    FV_LAUNCH(k_grad_scalar, s->cellGrid, s->g, dP, dOut);
    return FY_OK;
}
int fvDivFlux(fy_ctx* h, FvState* s, const double* dPhiSlots, double* dOut)
{
    FV_LAUNCH(k_div_flux, s->cellGrid, s->g, dPhiSlots, dOut);
    return FY_OK;
}
int fvFacesToSlots(fy_ctx* h, FvState* s, int n, const double* dFaces, double* dSlots)
{
    FV_LAUNCH(k_faces_to_slots, s->cellGrid, n, s->dSlotOfFace, dFaces, dSlots);
    return FY_OK;
}
int fvSlotsToFaces(fy_ctx* h, FvState* s, int n, const double* dSlots, double* dFaces)
{
    FV_LAUNCH(k_slots_to_faces, s->cellGrid, n, s->dSlotOfFace, dSlots, dFaces);
    return FY_OK;
}

// CourantNo.H + vGrad = fvc::grad(U)                                                   icoFoamYade.C:68-71
int fvIcoPre(fy_ctx* h, FvState* s, double dt)
{
    FV_LAUNCH(k_courant, s->cellGrid, s->g, s->phi, dt, s->red, s->dStep);
    FV_LAUNCH(k_grad_vector, s->cellGrid, s->g, h->dField[FY_F_U], h->dField[FY_F_VGRAD]);
    return FY_OK;
}

// icoFoamYade.C:79-140
int fvIcoSolve(fy_ctx* h, FvState* s, double dt)
{
    const BoxGeom& g = s->g;
    const int N = g.N, G = s->cellGrid;
    const double rDeltaT = 1.0 / dt;
    double* U = h->dField[FY_F_U];
    double* p = h->dField[FY_F_P];
    const fy_piso_controls& ctl = s->ctl;
    int rc;
    cudaEvent_t* ev = h->ev;
    float msMom = 0, msP = 0, msOther = 0, tmp = 0;
    s->stats.nPSolves = 0;
    for (auto& u : s->stats.U) u = fy_solver_perf{0, 0, 0, 0};

    cudaEventRecord(ev[0], h->stream);
    // oldTime fields
    FY_CUDA(cudaMemcpyAsync(s->U0, U, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    FY_CUDA(cudaMemcpyAsync(s->phi0, s->phi, (size_t)g.nSlots * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    FV_LAUNCH(k_assemble_U, G, g, s->nu, rDeltaT, s->phi, s->U0, h->dField[FY_F_USOURCE], s->diagU, s->loU, s->upU, s->srcU,
              s->dgU, s->rAU);
    if (ctl.momentumPredictor) {
        FV_LAUNCH(k_grad_scalar, G, g, p, s->gradP);
        FV_LAUNCH(k_usolve_setup, G, g, s->nu, s->phi, s->srcU, s->gradP, U, s->bU, s->psiU);
        for (int m = 0; m < 3; ++m) {
            if (!g.valid[m]) continue;
            if ((rc = fvSmoothSolve(h, s, s->dgU + (size_t)m * N, s->loU, s->upU, s->bU + (size_t)m * N,
                                    s->psiU + (size_t)m * N, ctl.UTol, ctl.URelTol, ctl.maxIter, &s->stats.U[m])))
                return rc;
            FV_LAUNCH(k_store_component, G, N, s->psiU + (size_t)m * N, m, U);
        }
    }
    cudaEventRecord(ev[1], h->stream);
    for (int corr = 1; corr <= ctl.nCorrectors; ++corr) {
        cudaEventRecord(ev[2], h->stream);
        FV_LAUNCH(k_HbyA, G, g, s->nu, s->phi0, s->loU, s->upU, s->srcU, U, s->rAU, s->HbyA);
        FV_LAUNCH(k_phiHbyA, G, g, rDeltaT, s->HbyA, s->U0, s->phi0, s->rAU, s->phiHbyA, s->upP);
        if (g.pNeedRef) {
            FV_LAUNCH(k_adjust_sum, G, g, s->phiHbyA, s->red, s->dStep);
            FV_LAUNCH(k_adjust_scale, G, g, s->phiHbyA, s->dStep);
        }
        cudaEventRecord(ev[3], h->stream);
        for (int nonOrth = 0; nonOrth <= ctl.nNonOrthogonalCorrectors; ++nonOrth) {
            FV_LAUNCH(k_pEqn, G, g, ctl.pRefCell, ctl.pRefValue, s->upP, s->phiHbyA, s->rAU, s->dgP, s->bP);
            const bool fin = corr == ctl.nCorrectors && nonOrth == ctl.nNonOrthogonalCorrectors;
            fy_solver_perf perf{0, 0, 0, 0};
            if ((rc = fvPcgSolve(h, s, s->dgP, s->upP, s->bP, p, fin ? ctl.pFinalTol : ctl.pTol,
                                 fin ? ctl.pFinalRelTol : ctl.pRelTol, ctl.maxIter, ctl.preconditioner, &perf)))
                return rc;
            if (s->stats.nPSolves < 8) s->stats.p[s->stats.nPSolves] = perf;
            s->stats.nPSolves++;
            if (nonOrth == ctl.nNonOrthogonalCorrectors) FV_LAUNCH(k_flux_update, G, g, s->upP, s->phiHbyA, p, s->rAU, s->phi);
        }
        cudaEventRecord(ev[4], h->stream);
        FV_LAUNCH(k_correct_U, G, g, dt, corr - 1, s->phi, p, s->HbyA, s->rAU, U, s->gradP, s->red, s->dStep);
        cudaEventRecord(ev[5], h->stream);
        FY_CUDA(cudaEventSynchronize(ev[5]));
        cudaEventElapsedTime(&tmp, ev[2], ev[3]); msOther += tmp;
        cudaEventElapsedTime(&tmp, ev[3], ev[4]); msP += tmp;
        cudaEventElapsedTime(&tmp, ev[4], ev[5]); msOther += tmp;
    }
    cudaEventElapsedTime(&msMom, ev[0], ev[1]);
    h->phaseMs[6] = msMom;
    h->phaseMs[7] = msP;
    s->stats.pad_ = 0;
    FY_CUDA(cudaMemcpyAsync(s->hStep, s->dStep, sizeof(FvStepDev), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    if (g.pNeedRef && s->hStep->adjustFail) {
        h->err = "adjustPhi: continuity error cannot be removed by adjusting the outflow";
        return FY_ERR_NOT_CONVERGED;
    }
    s->stats.CoNum = s->hStep->CoNum;
    s->stats.meanCoNum = s->hStep->meanCoNum;
    s->stats.sumLocalContErr = s->hStep->sumLocal;
    s->stats.globalContErr = s->hStep->global;
    for (int q = 0; q < 8; ++q) {
        s->stats.corrSumLocal[q] = s->hStep->corrSumLocal[q];
        s->stats.corrGlobal[q] = s->hStep->corrGlobal[q];
    }
    for (int q = 0; q < ctl.nCorrectors && q < 8; ++q) s->cumulativeContErr += s->hStep->corrGlobal[q];
    s->stats.cumulativeContErr = s->cumulativeContErr;
    s->fluidMs[0] = msMom; s->fluidMs[1] = msP; s->fluidMs[2] = msOther;
    return FY_OK;
}

// ---------------------------------------------------------------------------------------------
// creation: checks that the LDU mesh IS the uniform hex box (and in blockMesh order), derives the
// per-direction constants from the caller's own face arrays (so they carry the caller's bits), builds
// the OpenFOAM-face-order -> owner-slot map.
// ---------------------------------------------------------------------------------------------
namespace {
template <class T>
int devAlloc(fy_ctx* h, T** p, size_t n)
{
    FY_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    FY_CUDA(cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    return FY_OK;
}
}

int fvCreate(fy_ctx* h, const fy_mesh_desc* m)
{
    FvState* s = new FvState();
    h->fv = s;
    fy_piso_default_controls(&s->ctl);
    std::memset(&s->stats, 0, sizeof(s->stats));
    auto no = [&](const std::string& why) { s->supported = false; s->why = why; return FY_OK; };
    const int nx = m->boxN[0], ny = m->boxN[1], nz = m->boxN[2];
    if (nx <= 0 || ny <= 0 || nz <= 0 || (long long)nx * ny * nz != m->nCells) return no("mesh has no hex-box descriptor (boxN)");
    const int N = m->nCells, Fi = m->nInternalFaces;
    if (Fi != 3LL * N - (long long)nx * ny - (long long)ny * nz - (long long)nx * nz) return no("internal face count is not that of a hex box");
    if (m->nPatches <= 0) return no("no boundary patches");
    BoxGeom& g = s->g;
    std::memset(&g, 0, sizeof(g));
    g.nx = nx; g.ny = ny; g.nz = nz; g.N = N; g.sy = nx; g.sz = nx * ny;
    g.off[0] = 3 * N; g.off[1] = g.off[0] + ny * nz; g.off[2] = g.off[1] + nx * nz;
    g.nSlots = g.off[2] + nx * ny;
    g.V = m->V[0];
    double sumV = 0;
    for (int c = 0; c < N; ++c) {
        if (m->V[c] != g.V) return no("cell volumes are not uniform");
        sumV += m->V[c];
    }
    g.sumV = sumV;
    s->hSlotOfFace.assign((size_t)Fi, 0);
    bool seen[3] = {false, false, false};
    const int stride[3] = {1, nx, nx * ny};
    for (int f = 0; f < Fi; ++f) {
        const int o = m->owner[f], dn = m->neighbour[f] - o;
        const double* sf = m->Sf + 3 * (size_t)f;
        int d = 0;
        if (std::fabs(sf[1]) > std::fabs(sf[d])) d = 1;
        if (std::fabs(sf[2]) > std::fabs(sf[d])) d = 2;
        if (o < 0 || o >= N || dn != stride[d]) return no("face list is not a hex box in blockMesh order");
        for (int q = 0; q < 3; ++q)
            if (q != d && sf[q] != 0.0) return no("face area vectors are not axis aligned");
        if (!seen[d]) {
            seen[d] = true;
            g.Sf[d] = sf[d]; g.magSf[d] = m->magSf[f]; g.dc[d] = m->deltaCoeffs[f]; g.w[d] = m->weights[f];
        } else if (g.Sf[d] != sf[d] || g.magSf[d] != m->magSf[f] || g.dc[d] != m->deltaCoeffs[f] || g.w[d] != m->weights[f]) {
            return no("internal face geometry is not uniform per direction");
        }
        s->hSlotOfFace[f] = d * N + o;
    }
    // boundary faces -> sides
    int firstFaceOfSide[6], lastFaceOfSide[6] = {0, 0, 0, 0, 0, 0}, facesOfSide[6] = {0, 0, 0, 0, 0, 0};
    bool sideSeen[6] = {false, false, false, false, false, false};
    for (int q = 0; q < 6; ++q) { firstFaceOfSide[q] = 1 << 30; g.kindU[q] = g.kindP[q] = FV_EMPTY; }
    int b = 0;
    std::vector<int> bslot;
    for (int pI = 0; pI < m->nPatches; ++pI) {
        const fy_patch_desc& pd = m->patches[pI];
        for (int q = 0; q < pd.nFaces; ++q, ++b) {
            const double* sf = pd.Sf + 3 * (size_t)q;
            int d = 0;
            if (std::fabs(sf[1]) > std::fabs(sf[d])) d = 1;
            if (std::fabs(sf[2]) > std::fabs(sf[d])) d = 2;
            for (int r = 0; r < 3; ++r)
                if (r != d && sf[r] != 0.0) return no("boundary face area vectors are not axis aligned");
            const int side = 2 * d + (sf[d] > 0 ? 1 : 0);
            const int c = pd.faceCells[q];
            const int i = c % nx, j = (c / nx) % ny, k = c / (nx * ny);
            const int v = d == 0 ? i : (d == 1 ? j : k), nd = d == 0 ? nx : (d == 1 ? ny : nz);
            if (c < 0 || c >= N || v != ((side & 1) ? nd - 1 : 0)) return no("boundary face is not on the box surface");
            if (!sideSeen[side]) {
                sideSeen[side] = true;
                firstFaceOfSide[side] = b;
                g.bSf[side] = sf[d]; g.bMagSf[side] = pd.magSf[q]; g.bDc[side] = pd.deltaCoeffs[q];
                g.kindU[side] = pd.bcU; g.kindP[side] = pd.bcP;
                for (int r = 0; r < 3; ++r) g.valU[side][r] = pd.valueU[r];
                g.valP[side] = pd.valueP;
            } else {
                if (g.bSf[side] != sf[d] || g.bMagSf[side] != pd.magSf[q] || g.bDc[side] != pd.deltaCoeffs[q])
                    return no("boundary face geometry is not uniform per side");
                if (g.kindU[side] != pd.bcU || g.kindP[side] != pd.bcP || g.valP[side] != pd.valueP ||
                    g.valU[side][0] != pd.valueU[0] || g.valU[side][1] != pd.valueU[1] || g.valU[side][2] != pd.valueU[2])
                    return no("a box side carries more than one boundary condition");
            }
            lastFaceOfSide[side] = b;
            facesOfSide[side]++;
            int slot;
            if (side & 1) slot = d * N + c;
            else slot = g.off[d] + (d == 0 ? j + ny * k : (d == 1 ? i + nx * k : i + nx * j));
            bslot.push_back(slot);
        }
    }
    s->nFi = Fi;
    s->nB = b;
    if (b != 2 * (nx * ny + ny * nz + nx * nz)) return no("boundary does not cover the box surface exactly once");
    for (int q = 0; q < 6; ++q) {
        if (!sideSeen[q]) return no("a box side has no boundary faces");
        if (lastFaceOfSide[q] - firstFaceOfSide[q] + 1 != facesOfSide[q]) return no("the faces of a box side are not contiguous in the boundary list");
    }
    {
        // a side's faces must be contiguous in the boundary list (each side belongs to one patch, whole)
        int order[6] = {0, 1, 2, 3, 4, 5};
        std::sort(order, order + 6, [&](int a, int c2) { return firstFaceOfSide[a] < firstFaceOfSide[c2]; });
        for (int q = 0; q < 6; ++q) g.seq[q] = order[q];
    }
    for (int d = 0; d < 3; ++d) {
        const bool e0 = g.kindU[2 * d] == FV_EMPTY, e1 = g.kindU[2 * d + 1] == FV_EMPTY;
        if (e0 != e1) return no("empty patches must come in opposite pairs");
        if ((g.kindP[2 * d] == FV_EMPTY) != e0 || (g.kindP[2 * d + 1] == FV_EMPTY) != e1) return no("a patch must be empty for U and p alike");
        g.valid[d] = e0 ? 0 : 1;
        if (e0 && (d == 0 ? nx : (d == 1 ? ny : nz)) != 1) return no("an empty direction must be one cell thick");
    }
    g.pNeedRef = 1;
    for (int q = 0; q < 6; ++q)
        if (g.kindP[q] == FV_FIXED_VALUE) g.pNeedRef = 0;
    s->hSlotOfFace.insert(s->hSlotOfFace.end(), bslot.begin(), bslot.end());

    // device buffers
    int rc;
    const size_t NS = (size_t)g.nSlots, N3 = 3 * (size_t)N;
    if ((rc = devAlloc(h, &s->dSlotOfFace, s->hSlotOfFace.size()))) return rc;
    FY_CUDA(cudaMemcpyAsync(s->dSlotOfFace, s->hSlotOfFace.data(), s->hSlotOfFace.size() * sizeof(int),
                            cudaMemcpyHostToDevice, h->stream));
    double** bufs[] = {&s->phi, &s->phi0, &s->phiHbyA};
    for (auto pp : bufs) if ((rc = devAlloc(h, pp, NS))) return rc;
    double** b3[] = {&s->U0, &s->HbyA, &s->gradP, &s->loU, &s->upU, &s->srcU, &s->dgU, &s->bU, &s->psiU, &s->upP};
    for (auto pp : b3) if ((rc = devAlloc(h, pp, N3))) return rc;
    double** b1[] = {&s->rAU, &s->diagU, &s->bPrime, &s->dgP, &s->bP, &s->rD, &s->pA, &s->wA, &s->rA};
    for (auto pp : b1) if ((rc = devAlloc(h, pp, (size_t)N))) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    s->cellGrid = std::min((N + BLK - 1) / BLK, sms * 8);
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, (const void*)k_wave<OpGsFwd, false, false>, BLK, 0);
    perSm = std::max(1, std::min(perSm, 4));
    s->waveGrid = std::max(1, std::min((ny * nz + BLK - 1) / BLK, sms * perSm));
    if ((rc = devAlloc(h, &s->red.partial, 4 * (size_t)std::max(s->cellGrid, s->waveGrid)))) return rc;
    if ((rc = devAlloc(h, &s->red.ticket, 1))) return rc;
    if ((rc = devAlloc(h, &s->dSolve, 1))) return rc;
    if ((rc = devAlloc(h, &s->dStep, 1))) return rc;
    FY_CUDA(cudaHostAlloc((void**)&s->hSolve, sizeof(FvSolveDev), cudaHostAllocDefault));
    FY_CUDA(cudaHostAlloc((void**)&s->hStep, sizeof(FvStepDev), cudaHostAllocDefault));
    for (auto& e : s->pev) cudaEventCreate(&e);
    // the face field phi lives in owner slots; the ABI's FY_F_PHI is served through the slot map
    FY_CUDA(cudaStreamSynchronize(h->stream));
    s->supported = true;
    return FY_OK;
}

void fvDestroy(fy_ctx* h)
{
    FvState* s = h->fv;
    if (!s) return;
    void* ptrs[] = {s->dSlotOfFace, s->phi, s->phi0, s->phiHbyA, s->U0, s->HbyA, s->rAU, s->gradP, s->diagU, s->loU, s->upU,
                    s->srcU, s->dgU, s->bU, s->psiU, s->bPrime, s->upP, s->dgP, s->bP, s->rD, s->pA, s->wA, s->rA, s->stage,
                    s->red.partial, s->red.ticket, s->dSolve, s->dStep};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& e : s->pev) if (e) cudaEventDestroy(e);
    if (s->hSolve) cudaFreeHost(s->hSolve);
    if (s->hStep) cudaFreeHost(s->hStep);
    delete s;
    h->fv = nullptr;
}
