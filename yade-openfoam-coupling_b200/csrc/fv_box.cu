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
#include <algorithm>
#include <cmath>
#include <cstring>

#include "fv_solver.h"

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
__device__ __forceinline__ double fvPatchP(const BoxGeom& g, int s, double pc, double gradP = 0.0)
{
    if (g.kindP[s] == FV_FIXED_FLUX_PRESSURE) return pc + gradP / g.bDc[s];       // fixedGradient: p_P + gradient/deltaCoeffs
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
                a += g.bSf[s] * fvPatchP(g, s, pc, fvFluxGradP(g, s, c, i, j, k));
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
// pimpleFoamYade.C:73-76: the explicit operators whose results FoamYade reads in the Gaussian branch.
//   fvc::div(phi, U)        = surfaceIntegrate(phi_f * interpolate(U))                (gaussConvectionScheme, linear)
//   fvc::laplacian(gamma,U) = surfaceIntegrate((interpolate(gamma)*magSf) * snGrad(U)) (gaussLaplacianScheme,
//                             snGrad = deltaCoeffs*(U_N - U_P); gammaB = value of gamma on the patches -- alphac's are
//                             `calculated` and hold the 1.0 FoamYade::initFields assigns field-wide, F.C:67)
// Face order per cell as everywhere: z-, y-, x- (cell is the neighbour: -=), x+, y+, z+ (owner: +=), boundary.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK)
    k_div_phi_vector(BoxGeom g, const double* __restrict__ phi, const double* __restrict__ U, double* __restrict__ out)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double uc[3] = {U[3 * (size_t)c], U[3 * (size_t)c + 1], U[3 * (size_t)c + 2]};
        double a[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 2; d >= 0; --d) {
            if (fvIdx(d, i, j, k) == 0) continue;
            const int cn = c - fvStride(g, d);
            const double ph = phi[d * (size_t)g.N + cn];
            const double* un = U + 3 * (size_t)cn;
#pragma unroll
            for (int m = 0; m < 3; ++m) a[m] -= ph * fvLerp(g.w[d], un[m], uc[m]);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (fvIdx(d, i, j, k) == fvN(g, d) - 1) continue;
            const double ph = phi[d * (size_t)g.N + c];
            const double* un = U + 3 * (size_t)(c + fvStride(g, d));
#pragma unroll
            for (int m = 0; m < 3; ++m) a[m] += ph * fvLerp(g.w[d], uc[m], un[m]);
        }
        if (!fvInterior(g, i, j, k)) {
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                const double ph = phi[fvSideSlot(g, s, c, i, j, k)];
#pragma unroll
                for (int m = 0; m < 3; ++m) a[m] += ph * (g.kindU[s] == FV_FIXED_VALUE ? g.valU[s][m] : uc[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) out[3 * (size_t)c + m] = a[m] / g.V;
    }
}

// out = scale * fvc::laplacian(gamma, U)   (scale = 2 nu gives pimpleFoamYade's divT; 1.0 is exact)
__global__ void __launch_bounds__(BLK)
k_laplacian_gamma_vector(BoxGeom g, double scale, const double* __restrict__ gamma, double gammaB, const double* __restrict__ U,
                         double* __restrict__ out)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double uc[3] = {U[3 * (size_t)c], U[3 * (size_t)c + 1], U[3 * (size_t)c + 2]};
        const double gc = gamma[c];
        double a[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 2; d >= 0; --d) {
            if (fvIdx(d, i, j, k) == 0) continue;
            const int cn = c - fvStride(g, d);
            const double gm = fvLerp(g.w[d], gamma[cn], gc) * g.magSf[d];
            const double* un = U + 3 * (size_t)cn;
#pragma unroll
            for (int m = 0; m < 3; ++m) a[m] -= gm * (g.dc[d] * (uc[m] - un[m]));
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (fvIdx(d, i, j, k) == fvN(g, d) - 1) continue;
            const int cp = c + fvStride(g, d);
            const double gm = fvLerp(g.w[d], gc, gamma[cp]) * g.magSf[d];
            const double* un = U + 3 * (size_t)cp;
#pragma unroll
            for (int m = 0; m < 3; ++m) a[m] += gm * (g.dc[d] * (un[m] - uc[m]));
        }
        if (!fvInterior(g, i, j, k)) {
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                const double gm = gammaB * g.bMagSf[s];
#pragma unroll
                for (int m = 0; m < 3; ++m)
                    a[m] += gm * (g.kindU[s] == FV_FIXED_VALUE ? g.bDc[s] * (g.valU[s][m] - uc[m]) : 0.0);
            }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) out[3 * (size_t)c + m] = scale * (a[m] / g.V);
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
    double v[4] = {0.0, 0.0, 0.0, 0.0};       // massIn fixedMassOut adjustableMassOut sum|phi| (internal faces: OpenFOAM's
                                              // sum(mag(phi)) of a surface field reduces the internal field only)
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        if (i < g.nx - 1) v[3] += fabs(phiHbyA[c]);
        if (j < g.ny - 1) v[3] += fabs(phiHbyA[(size_t)g.N + c]);
        if (k < g.nz - 1) v[3] += fabs(phiHbyA[2 * (size_t)g.N + c]);
        if (fvInterior(g, i, j, k)) continue;
        for (int q = 0; q < 6; ++q) {
            const int s = g.seq[q];
            if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
            const double ph = phiHbyA[fvSideSlot(g, s, c, i, j, k)];
            if (ph < 0) v[0] -= ph;
            else if (g.kindU[s] == FV_FIXED_VALUE) v[1] += ph;
            else v[2] += ph;
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
// pimpleFoamYade: UcEqn.H + pEqn.H + continuityErrs.H on the box  (pimpleFoamYade.C:82-104, one outer corrector,
// laminar).  Operator definitions and the boundary / old-time conventions: oracle/fv_oracle.cc (pimpleSolve).
// alphacf = fvc::interpolate(alphac) is recomputed from alphac wherever a face needs it (its patches hold 1.0, so
// alphacf_b*x == x bit for bit and the icoFoam boundary-coefficient helpers serve unchanged).
// ---------------------------------------------------------------------------------------------
constexpr double FV_ALPHA_B = 1.0;

// surfaceIntegrate(alphacf*phi) of one cell -- fvc::div(alphaPhic), fvc::div(alphacf*phiHbyA), fvc::div(alphacf*phic)
__device__ __forceinline__ double fvDivAlphaCell(const BoxGeom& g, const double* __restrict__ alpha,
                                                 const double* __restrict__ phi, int c, int i, int j, int k)
{
    const double ac = alpha[c];
    double d = 0.0;
    if (k > 0) d -= fvLerp(g.w[2], alpha[c - g.sz], ac) * phi[2 * g.N + c - g.sz];
    if (j > 0) d -= fvLerp(g.w[1], alpha[c - g.sy], ac) * phi[g.N + c - g.sy];
    if (i > 0) d -= fvLerp(g.w[0], alpha[c - 1], ac) * phi[c - 1];
    if (i < g.nx - 1) d += fvLerp(g.w[0], ac, alpha[c + 1]) * phi[c];
    if (j < g.ny - 1) d += fvLerp(g.w[1], ac, alpha[c + g.sy]) * phi[g.N + c];
    if (k < g.nz - 1) d += fvLerp(g.w[2], ac, alpha[c + g.sz]) * phi[2 * g.N + c];
    if (!fvInterior(g, i, j, k)) {
        for (int q = 0; q < 6; ++q) {
            const int s = g.seq[q];
            if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
            d += FV_ALPHA_B * phi[fvSideSlot(g, s, c, i, j, k)];
        }
    }
    return d / g.V;
}

// fvc::div((alpha*nuEff)*dev2(T(fvc::grad(U)))): the explicit half of the laminar divDevRhoReff.  Only row d of the
// stress tensor meets a face normal to d; patch values of grad(U) carry gaussGrad's normal-gradient correction.
__device__ __forceinline__ void fvDevRow(const double* __restrict__ gr, int d, double a, double* x)
{
    const double sph = (2.0 / 3.0) * (gr[0] + gr[4] + gr[8]);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        double t = gr[3 * m + d];                     // T(grad)_dm
        if (m == d) t -= sph;
        x[m] = a * t;
    }
}
__global__ void __launch_bounds__(BLK)
k_pim_div_dev(BoxGeom g, double nu, const double* __restrict__ alpha, const double* __restrict__ U,
              const double* __restrict__ vGrad, double* __restrict__ out)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double* gc = vGrad + 9 * (size_t)c;
        const double ac = alpha[c] * nu;
        double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 2; d >= 0; --d) {
            if (fvIdx(d, i, j, k) == 0) continue;
            const int cn = c - fvStride(g, d);
            double xp[3], xn[3];
            fvDevRow(vGrad + 9 * (size_t)cn, d, alpha[cn] * nu, xp);
            fvDevRow(gc, d, ac, xn);
#pragma unroll
            for (int m = 0; m < 3; ++m) acc[m] -= g.Sf[d] * fvLerp(g.w[d], xp[m], xn[m]);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (fvIdx(d, i, j, k) == fvN(g, d) - 1) continue;
            const int cp = c + fvStride(g, d);
            double xp[3], xn[3];
            fvDevRow(gc, d, ac, xp);
            fvDevRow(vGrad + 9 * (size_t)cp, d, alpha[cp] * nu, xn);
#pragma unroll
            for (int m = 0; m < 3; ++m) acc[m] += g.Sf[d] * fvLerp(g.w[d], xp[m], xn[m]);
        }
        if (!fvInterior(g, i, j, k)) {
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                const int d = s >> 1;
                const double n = g.bSf[s] / g.bMagSf[s];
                double gb[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) gb[t] = gc[t];
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    const double sn = g.kindU[s] == FV_FIXED_VALUE ? g.bDc[s] * (g.valU[s][m] - U[3 * (size_t)c + m]) : 0.0;
                    const double nG = n * gc[3 * d + m];
                    gb[3 * d + m] = gc[3 * d + m] + n * (sn - nG);
                }
                double xb[3];
                fvDevRow(gb, d, FV_ALPHA_B * nu, xb);
#pragma unroll
                for (int m = 0; m < 3; ++m) acc[m] += g.bSf[s] * xb[m];
            }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) out[3 * (size_t)c + m] = acc[m] / g.V;
    }
}

// UcEqn = fvm::ddt(alphac,Uc) + fvm::div(alphaPhic,Uc) - fvm::Sp(fvc::ddt(alphac) + fvc::div(alphaPhic),Uc)
//         + divDevRhoReff(Uc) == fvm::Sp(uSourceDrag,Uc)                                        pim/UcEqn.H:3-11
__global__ void __launch_bounds__(BLK)
k_pim_assemble_U(BoxGeom g, double nu, double rDeltaT, const double* __restrict__ phi, const double* __restrict__ alpha,
                 const double* __restrict__ alpha0, const double* __restrict__ U0, const double* __restrict__ uSourceDrag,
                 const double* __restrict__ divDev, double relax, const double* __restrict__ Ucur, double* __restrict__ diagU,
                 double* __restrict__ loU, double* __restrict__ upU, double* __restrict__ srcU, double* __restrict__ dgU,
                 double* __restrict__ rAU)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const int N = g.N;
        const double ac = alpha[c], a0 = alpha0[c];
        const double diagD = (rDeltaT * ac) * g.V;
        double diagC = 0.0, diagL = 0.0, dv = 0.0;
        double sumOff = 0.0;                          // lduMatrix::sumMagOffDiag of this row, in face order (UcEqn.relax())
#pragma unroll
        for (int d = 2; d >= 0; --d) {
            const int v = fvIdx(d, i, j, k), sd = fvStride(g, d);
            if (v > 0) {
                const double an = alpha[c - sd];
                const double aph = fvLerp(g.w[d], an, ac) * phi[d * N + c - sd];
                const double lowerC = -g.w[d] * aph;
                const double upperC = lowerC + aph;
                const double upperL = g.dc[d] * (fvLerp(g.w[d], an * nu, ac * nu) * g.magSf[d]);
                diagC -= upperC;
                diagL -= upperL;
                dv -= aph;
                sumOff += fabs(lowerC + (-upperL));   // this cell is the face's neighbour: |Lower[face]|
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
            double lo = 0.0, up = 0.0;
            if (v < nd - 1) {
                const double an = alpha[c + sd];
                const double aph = fvLerp(g.w[d], ac, an) * phi[d * N + c];
                const double lowerC = -g.w[d] * aph;
                const double upperC = lowerC + aph;
                const double upperL = g.dc[d] * (fvLerp(g.w[d], ac * nu, an * nu) * g.magSf[d]);
                diagC -= lowerC;
                diagL -= upperL;
                dv += aph;
                lo = lowerC + (-upperL);
                up = upperC + (-upperL);
                sumOff += fabs(up);                   // this cell is the face's owner: |Upper[face]|
            }
            loU[d * N + c] = lo;
            upU[d * N + c] = up;
        }
        double D = 0.0, dgB[3] = {0.0, 0.0, 0.0};
        if (!fvInterior(g, i, j, k)) {
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                dv += FV_ALPHA_B * phi[fvSideSlot(g, s, c, i, j, k)];
            }
        }
        const double spDiv = rDeltaT * (ac - a0) + dv / g.V;
        double diag = (((diagD + diagC) - g.V * spDiv) + (-diagL)) - g.V * uSourceDrag[c];
        double relaxSrc = 0.0;                        // D - D0 of fvMatrix::relax: the source gets (D - D0)*psi
        if (relax > 0.0) {
            // UcEqn.relax()   pim/UcEqn.H:13  [OF-6 fvMatrix.C relax(alpha)]: boundary internal coefficients count with
            // their largest-magnitude component while dominance is enforced, and leave with their smallest component
            const double D0 = diag;
            const bool bnd = !fvInterior(g, i, j, k);
            if (bnd) {
                for (int q = 0; q < 6; ++q) {
                    const int s = g.seq[q];
                    if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                    const double phib = FV_ALPHA_B * phi[fvSideSlot(g, s, c, i, j, k)];
                    double ic[3], bc;
#pragma unroll
                    for (int m = 0; m < 3; ++m) fvBCoefU(g, s, phib, FV_ALPHA_B * nu, m, ic[m], bc);
                    diag += fmax(fmax(fabs(ic[0]), fabs(ic[1])), fabs(ic[2]));
                }
            }
            diag = fmax(fabs(diag), sumOff);
            diag /= relax;
            if (bnd) {
                for (int q = 0; q < 6; ++q) {
                    const int s = g.seq[q];
                    if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                    const double phib = FV_ALPHA_B * phi[fvSideSlot(g, s, c, i, j, k)];
                    double ic[3], bc;
#pragma unroll
                    for (int m = 0; m < 3; ++m) fvBCoefU(g, s, phib, FV_ALPHA_B * nu, m, ic[m], bc);
                    diag -= fmin(fmin(ic[0], ic[1]), ic[2]);
                }
            }
            relaxSrc = diag - D0;
        }
        diagU[c] = diag;
        D = diag;
        dgB[0] = dgB[1] = dgB[2] = diag;
        if (!fvInterior(g, i, j, k)) {
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindU[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                const double phib = FV_ALPHA_B * phi[fvSideSlot(g, s, c, i, j, k)];
                double ic[3], bc;
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    fvBCoefU(g, s, phib, FV_ALPHA_B * nu, m, ic[m], bc);
                    dgB[m] += ic[m];
                }
                D += (ic[0] + ic[1] + ic[2]) / 3.0;
            }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            dgU[m * (size_t)N + c] = dgB[m];
            double sU = ((rDeltaT * a0) * U0[3 * (size_t)c + m]) * g.V;
            sU -= g.V * (-divDev[3 * (size_t)c + m]);
            if (relax > 0.0) sU += relaxSrc * Ucur[3 * (size_t)c + m];
            srcU[3 * (size_t)c + m] = sU;
        }
        rAU[c] = 1.0 / (D / g.V);
    }
}

// phicForces = fvc::flux(rAUc*uSource) + rAUcf*(g & Sf)                                        pim/UcEqn.H:17-20
__global__ void __launch_bounds__(BLK)
k_pim_forces(BoxGeom g, double g0, double g1, double g2, const double* __restrict__ rAU, const double* __restrict__ uSource,
             double* __restrict__ phicForces)
{
    const double gv[3] = {g0, g1, g2};
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double rc = rAU[c];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d);
            if (v < nd - 1) {
                const int n = c + fvStride(g, d);
                const double rn = rAU[n];
                const double flux = g.Sf[d] * fvLerp(g.w[d], rc * uSource[3 * (size_t)c + d], rn * uSource[3 * (size_t)n + d]);
                phicForces[d * g.N + c] = flux + fvLerp(g.w[d], rc, rn) * (gv[d] * g.Sf[d]);
            } else {
                const int s = 2 * d + 1;
                phicForces[d * g.N + c] = g.kindU[s] == FV_EMPTY ? 0.0 : g.bSf[s] * (rc * 0.0) + rc * (gv[d] * g.bSf[s]);
            }
            if (v == 0) {
                const int s = 2 * d;
                phicForces[fvSideSlot(g, s, c, i, j, k)] =
                    g.kindU[s] == FV_EMPTY ? 0.0 : g.bSf[s] * (rc * 0.0) + rc * (gv[d] * g.bSf[s]);
            }
        }
    }
}

// fvc::reconstruct(ssf) of one cell: inv(surfaceSum(SfHat*Sf)) & surfaceSum(SfHat*ssf).  On the box the tensor is
// diagonal; `face(d, hi)` returns the (SfHat component, ssf) pair of the cell's lower / upper face in direction d
// (ok=false: an empty face, no contribution).  g.reconRm = directions tensorField inv() removes (no faces at all).
template <class Face>
__device__ __forceinline__ void fvReconstruct(const BoxGeom& g, Face face, double* out)
{
    double T[3], v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        T[d] = 0.0;
        v[d] = 0.0;
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
            double nHat, area, ssf;
            if (!face(d, hi, nHat, area, ssf)) continue;
            T[d] += nHat * area;
            v[d] += nHat * ssf;
        }
        if (g.reconRm[d]) T[d] += 1.0;
    }
    const double det = T[0] * T[1] * T[2];
    double inv[3] = {(T[1] * T[2]) / det, (T[0] * T[2]) / det, (T[0] * T[1]) / det};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (g.reconRm[d]) inv[d] -= 1.0;
        out[d] = inv[d] * v[d];
    }
}

// -fvc::reconstruct(phicForces/rAUcf - fvc::snGrad(p)*magSf): the momentum predictor's source, stored negated so the
// icoFoam set-up kernel (source + V*(-x)) adds it                                                pim/UcEqn.H:24-32
__global__ void __launch_bounds__(BLK)
k_pim_predictor_source(BoxGeom g, const double* __restrict__ phicForces, const double* __restrict__ rAU,
                       const double* __restrict__ p, double* __restrict__ negRecon)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double rc = rAU[c], pc = p[c];
        auto face = [&](int d, int hi, double& nHat, double& area, double& ssf) -> bool {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
            const bool bnd = hi ? (v == nd - 1) : (v == 0);
            if (!bnd) {
                const int P = hi ? c : c - sd, Nb = hi ? c + sd : c;
                const double rf = fvLerp(g.w[d], rAU[P], rAU[Nb]);
                nHat = g.Sf[d] / g.magSf[d];
                area = g.Sf[d];
                ssf = phicForces[d * g.N + P] / rf - (g.dc[d] * (p[Nb] - p[P])) * g.magSf[d];
                return true;
            }
            const int s = 2 * d + hi;
            if (g.kindU[s] == FV_EMPTY) return false;
            const double sn = g.kindP[s] == FV_FIXED_VALUE ? g.bDc[s] * (g.valP[s] - pc) : fvFluxGradP(g, s, c, i, j, k);
            nHat = g.bSf[s] / g.bMagSf[s];
            area = g.bSf[s];
            ssf = phicForces[fvSideSlot(g, s, c, i, j, k)] / rc - sn * g.bMagSf[s];
            return true;
        };
        double r[3];
        fvReconstruct(g, face, r);
#pragma unroll
        for (int m = 0; m < 3; ++m) negRecon[3 * (size_t)c + m] = -r[m];
    }
}

// phiHbyA = fvc::flux(HbyA) + alphacf*rAUcf*fvc::ddtCorr(Uc, phic)  and  upper(pEqn) = deltaCoeffs*((alphacf*rAUcf)*magSf)
//                                                                                       pim/pEqn.H:4-11, 26-31
__global__ void __launch_bounds__(BLK)
k_pim_phiHbyA(BoxGeom g, double rDeltaT, const double* __restrict__ HbyA, const double* __restrict__ U0,
              const double* __restrict__ phi0, const double* __restrict__ rAU, const double* __restrict__ alpha,
              double* __restrict__ phiHbyA, double* __restrict__ upP)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const int N = g.N;
        const double rc = rAU[c], ac = alpha[c];
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
                const double arf = fvLerp(g.w[d], ac, alpha[n]) * fvLerp(g.w[d], rc, rAU[n]);
                phiHbyA[d * N + c] = flux + arf * ((coeff * rDeltaT) * phiCorr);
                upP[d * N + c] = g.dc[d] * (arf * g.magSf[d]);
            } else {
                phiHbyA[d * N + c] = fvPhiHbyAB(g, 2 * d + 1, d, rDeltaT, hc, u0c, FV_ALPHA_B * rc, phi0[d * N + c]);
                upP[d * N + c] = 0.0;
            }
            if (v == 0) {
                const int sl = fvSideSlot(g, 2 * d, c, i, j, k);
                phiHbyA[sl] = fvPhiHbyAB(g, 2 * d, d, rDeltaT, hc, u0c, FV_ALPHA_B * rc, phi0[sl]);
            }
        }
    }
}

// phiHbyA += phicForces                                                                          pim/pEqn.H:18
__global__ void __launch_bounds__(BLK) k_pim_add_forces(int n, const double* __restrict__ phicForces, double* __restrict__ phiHbyA)
{
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) phiHbyA[q] += phicForces[q];
}

// constrainPressure(p, Uc, phiHbyA, rAUcf)  pim/pEqn.H:21  [OF-6 constrainPressure.C]: on fixedFluxPressure patches
// snGrad(p) = (phiHbyA_b - Sf_b & U_b)/(magSf_b rAUcf_b), so that the corrected flux through the face equals the wall's
__global__ void __launch_bounds__(BLK)
k_pim_constrain_pressure(BoxGeom g, const double* __restrict__ phiHbyA, const double* __restrict__ U,
                         const double* __restrict__ rAU, double* __restrict__ bGradP)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        if (fvInterior(g, i, j, k)) continue;
        for (int s = 0; s < 6; ++s) {
            if (g.kindP[s] != FV_FIXED_FLUX_PRESSURE || !fvOnSide(g, s, i, j, k)) continue;
            const int d = s >> 1, sl = fvSideSlot(g, s, c, i, j, k);
            const double ub = g.kindU[s] == FV_FIXED_VALUE ? g.valU[s][d] : U[3 * (size_t)c + d];
            double su = 0.0;                     // Sf_b & U_b, summed over the three components like the CPU loop (two are 0*u)
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const double um = g.kindU[s] == FV_FIXED_VALUE ? g.valU[s][m] : U[3 * (size_t)c + m];
                su += (m == d ? g.bSf[s] : 0.0) * um;
            }
            (void)ub;
            bGradP[sl] = (phiHbyA[sl] - su) / (g.bMagSf[s] * rAU[c]);
        }
    }
}

// fvm::laplacian(alphacf*rAUcf, p) == fvc::ddt(alphac) + fvc::div(alphacf*phiHbyA); setReference    pim/pEqn.H:26-33
__global__ void __launch_bounds__(BLK)
k_pim_pEqn(BoxGeom g, int pRefCell, double pRefValue, double rDeltaT, const double* __restrict__ upP,
           const double* __restrict__ phiHbyA, const double* __restrict__ rAU, const double* __restrict__ alpha,
           const double* __restrict__ alpha0, double* __restrict__ dgP, double* __restrict__ bP)
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
        src += g.V * (rDeltaT * (alpha[c] - alpha0[c]) + fvDivAlphaCell(g, alpha, phiHbyA, c, i, j, k));
        if (g.pNeedRef && c == pRefCell) {
            src += diag * pRefValue;
            diag += diag;
        }
        if (!fvInterior(g, i, j, k)) {
            const double rc = FV_ALPHA_B * rAU[c];
            for (int q = 0; q < 6; ++q) {
                const int s = g.seq[q];
                if (g.kindP[s] == FV_EMPTY || !fvOnSide(g, s, i, j, k)) continue;
                double ic, bc;
                fvBCoefP(g, s, rc, ic, bc, fvFluxGradP(g, s, c, i, j, k));
                diag += ic;
                src += bc;
            }
        }
        dgP[c] = diag;
        bP[c] = src;
    }
}

// phic = phiHbyA - pEqn.flux()/alphacf                                                           pim/pEqn.H:39
__global__ void __launch_bounds__(BLK)
k_pim_flux_update(BoxGeom g, const double* __restrict__ upP, const double* __restrict__ phiHbyA, const double* __restrict__ p,
                  const double* __restrict__ rAU, const double* __restrict__ alpha, double* __restrict__ phi)
{
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const int N = g.N;
        const double pc = p[c], rc = FV_ALPHA_B * rAU[c], ac = alpha[c];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
            if (v < nd - 1) {
                const double u = upP[d * N + c];
                phi[d * N + c] = phiHbyA[d * N + c] - (u * p[c + sd] - u * pc) / fvLerp(g.w[d], ac, alpha[c + sd]);
            } else {
                double ic, bc;
                fvBCoefP(g, 2 * d + 1, rc, ic, bc, fvFluxGradP(g, 2 * d + 1, c, i, j, k));
                phi[d * N + c] = phiHbyA[d * N + c] - (ic * pc - bc) / FV_ALPHA_B;
            }
            if (v == 0) {
                double ic, bc;
                fvBCoefP(g, 2 * d, rc, ic, bc, fvFluxGradP(g, 2 * d, c, i, j, k));
                const int sl = fvSideSlot(g, 2 * d, c, i, j, k);
                phi[sl] = phiHbyA[sl] - (ic * pc - bc) / FV_ALPHA_B;
            }
        }
    }
}

// continuityErrs.H (contErr = fvc::ddt(alphac) + fvc::div(alphacf*phic)) and
// p.relax()   pim/pEqn.H:41  [OF-6 GeometricField::relax(alpha)]: p = prevIter + alpha*(p - prevIter)
__global__ void __launch_bounds__(BLK)
k_relax_field(int n, double a, const double* __restrict__ prev, double* __restrict__ x)
{
    for (int c = blockIdx.x * BLK + threadIdx.x; c < n; c += gridDim.x * BLK) x[c] = prev[c] + a * (x[c] - prev[c]);
}

// Uc = HbyA + rAUc*fvc::reconstruct((phicForces - pEqn.flux()/alphacf)/rAUcf)      pim/continuityErrs.H, pEqn.H:43-45
__global__ void __launch_bounds__(BLK)
k_pim_correct_U(BoxGeom g, double dt, double rDeltaT, int corr, const double* __restrict__ phi, const double* __restrict__ p,
                const double* __restrict__ upP, const double* __restrict__ phicForces, const double* __restrict__ HbyA,
                const double* __restrict__ rAU, const double* __restrict__ alpha, const double* __restrict__ alpha0,
                double* __restrict__ U, FvRed red, FvStepDev* out)
{
    double sums[2] = {0.0, 0.0};
    FV_CELL_LOOP(g, c) {
        int i, j, k;
        fvIJK(g, c, i, j, k);
        const double e = rDeltaT * (alpha[c] - alpha0[c]) + fvDivAlphaCell(g, alpha, phi, c, i, j, k);
        sums[0] += fabs(e) * g.V;
        sums[1] += e * g.V;
        const double rc = rAU[c], pc = p[c];
        auto face = [&](int d, int hi, double& nHat, double& area, double& ssf) -> bool {
            const int v = fvIdx(d, i, j, k), nd = fvN(g, d), sd = fvStride(g, d);
            const bool bnd = hi ? (v == nd - 1) : (v == 0);
            if (!bnd) {
                const int P = hi ? c : c - sd, Nb = hi ? c + sd : c;
                const double u = upP[d * g.N + P];
                const double fluxByA = (u * p[Nb] - u * p[P]) / fvLerp(g.w[d], alpha[P], alpha[Nb]);
                nHat = g.Sf[d] / g.magSf[d];
                area = g.Sf[d];
                ssf = (phicForces[d * g.N + P] - fluxByA) / fvLerp(g.w[d], rAU[P], rAU[Nb]);
                return true;
            }
            const int s = 2 * d + hi;
            if (g.kindU[s] == FV_EMPTY) return false;
            double ic, bc;
            fvBCoefP(g, s, FV_ALPHA_B * rc, ic, bc, fvFluxGradP(g, s, c, i, j, k));
            nHat = g.bSf[s] / g.bMagSf[s];
            area = g.bSf[s];
            ssf = (phicForces[fvSideSlot(g, s, c, i, j, k)] - (ic * pc - bc) / FV_ALPHA_B) / rc;
            return true;
        };
        double r[3];
        fvReconstruct(g, face, r);
#pragma unroll
        for (int m = 0; m < 3; ++m) U[3 * (size_t)c + m] = HbyA[3 * (size_t)c + m] + rc * r[m];
    }
    const double sumV = g.sumV;
    fvGridReduce<2, false, BLK>(sums, red, [=](const double* t) {
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

}  // namespace

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

int fvDivPhiVector(fy_ctx* h, FvState* s, const double* dPhiSlots, const double* dU, double* dOut)
{
    FV_LAUNCH(k_div_phi_vector, s->cellGrid, s->g, dPhiSlots, dU, dOut);
    return FY_OK;
}
int fvLaplacianGammaVector(fy_ctx* h, FvState* s, double scale, const double* dGamma, double gammaB, const double* dU,
                           double* dOut)
{
    FV_LAUNCH(k_laplacian_gamma_vector, s->cellGrid, s->g, scale, dGamma, gammaB, dU, dOut);
    return FY_OK;
}

// pimpleFoamYade.C:71-76: CourantNo, then the four fields FoamYade's Gaussian branch reads --
//   ddtU_f = fvc::ddt(Uc) + fvc::div(phic, Uc)   gradP = fvc::grad(p)
//   divT   = 2 nu fvc::laplacian(alphac, Uc)     vGrad = fvc::grad(Uc)
// fvc::ddt(Uc) is taken before Uc changes in the new time step, when GeometricField::oldTime() has just stored
// Uc.oldTime() := Uc, so the Euler term is rDeltaT*(Uc - Uc) = +0 and ddtU_f == fvc::div(phic, Uc) bit for bit.
int fvPimplePre(fy_ctx* h, FvState* s, double dt)
{
    const double* U = h->dField[FY_F_U];
    FV_LAUNCH(k_courant, s->cellGrid, s->g, s->phi, dt, s->red, s->dStep);
    FV_LAUNCH(k_div_phi_vector, s->cellGrid, s->g, s->phi, U, h->dField[FY_F_DDTU]);
    FV_LAUNCH(k_grad_scalar, s->cellGrid, s->g, h->dField[FY_F_P], h->dField[FY_F_GRADP]);
    FV_LAUNCH(k_laplacian_gamma_vector, s->cellGrid, s->g, 2 * s->nu, h->dField[FY_F_ALPHA], 1.0, U, h->dField[FY_F_DIVT]);
    FV_LAUNCH(k_grad_vector, s->cellGrid, s->g, U, h->dField[FY_F_VGRAD]);
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
    if (s->hasFluxP) { h->err = "fy_ico_solve: fixedFluxPressure patches are served by fy_pimple_solve only"; return FY_ERR_UNSUPPORTED; }
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
        if ((rc = fvSmoothSetMatrix(h, s, s->loU, s->upU))) return rc;
        // decomposed run: the momentum components are independent solves with one matrix -- each goes to one rank
        // (component m to rank m mod min(ranks, 3)) and travels to the others afterwards; same arithmetic as one rank
        // solving all three
        const PenState& P = s->pen;
        const int nOwners = P.dist ? std::min(P.nranks, 3) : 1;
        for (int m = 0; m < 3; ++m) {
            if (!g.valid[m] || (P.dist && m % nOwners != P.rank)) continue;
            if ((rc = fvSmoothSolve(h, s, s->dgU + (size_t)m * N, s->bU + (size_t)m * N, s->psiU + (size_t)m * N, ctl.UTol,
                                    ctl.URelTol, ctl.maxIter, &s->stats.U[m])))
                return rc;
        }
        if (P.dist) {
            double* ptr[6];
            size_t cnt[6];
            int root[6], nb = 0;
            double hp[12];
            for (int m = 0; m < 3; ++m) {
                const fy_solver_perf& u = s->stats.U[m];
                hp[4 * m] = u.initialResidual; hp[4 * m + 1] = u.finalResidual; hp[4 * m + 2] = u.nIterations; hp[4 * m + 3] = 0;
                if (!g.valid[m]) continue;
                if (m % nOwners == P.rank)
                    FY_CUDA(cudaMemcpyAsync(P.perfBuf + 4 * m, hp + 4 * m, 4 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
                ptr[nb] = s->psiU + (size_t)m * N; cnt[nb] = (size_t)N; root[nb++] = m % nOwners;
                ptr[nb] = P.perfBuf + 4 * m; cnt[nb] = 4; root[nb++] = m % nOwners;
            }
            if ((rc = fvDistBroadcastMany(h, s, nb, ptr, cnt, root))) return rc;
            FY_CUDA(cudaMemcpyAsync(hp, P.perfBuf, sizeof(hp), cudaMemcpyDeviceToHost, h->stream));
            FY_CUDA(cudaStreamSynchronize(h->stream));
            for (int m = 0; m < 3; ++m) {
                if (!g.valid[m]) continue;
                s->stats.U[m].initialResidual = hp[4 * m]; s->stats.U[m].finalResidual = hp[4 * m + 1]; s->stats.U[m].nIterations = (int)hp[4 * m + 2];
            }
        }
        for (int m = 0; m < 3; ++m)
            if (g.valid[m]) FV_LAUNCH(k_store_component, G, N, s->psiU + (size_t)m * N, m, U);
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
                                 fin ? ctl.pFinalRelTol : ctl.pRelTol, ctl.maxIter, ctl.preconditioner, &perf,
                                 s->stats.nPSolves > 0)))     // 1/A() is fixed for the step: every pEqn has the same matrix
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

// pimpleFoamYade.C:82-104: alphacf / alphaPhic, then nOuterCorrectors x { UcEqn.H (+ relax), the PISO loop over pEqn.H (+ p.relax()),
// continuityErrs.H }.
// Reads the coupling operator's device fields alpha, uSource, uSourceDrag; alphac.oldTime() == alphac (see the oracle).
int fvPimpleSolve(fy_ctx* h, FvState* s, double dt, const double gvec[3])
{
    const BoxGeom& g = s->g;
    const int N = g.N, G = s->cellGrid;
    const double rDeltaT = 1.0 / dt;
    double* U = h->dField[FY_F_U];
    double* p = h->dField[FY_F_P];
    const double* alpha = h->dField[FY_F_ALPHA];
    const double* alpha0 = alpha;
    const fy_piso_controls& ctl = s->ctl;
    int rc;
    cudaEvent_t* ev = h->ev;
    float msMom = 0, msP = 0, msOther = 0, tmp = 0;
    s->stats.nPSolves = 0;
    for (auto& u : s->stats.U) u = fy_solver_perf{0, 0, 0, 0};
    if (!s->phicForces) {
        FY_CUDA(cudaMalloc((void**)&s->phicForces, (size_t)g.nSlots * sizeof(double)));
        FY_CUDA(cudaMalloc((void**)&s->divDev, 3 * (size_t)N * sizeof(double)));
    }

    const int nOuter = s->nOuter < 1 ? 1 : s->nOuter;
    if (nOuter * ctl.nCorrectors > 8) { h->err = "fy_pimple_solve: nOuterCorrectors x nCorrectors > 8 (fy_ico_stats slots)"; return FY_ERR_INVALID; }
    if (nOuter > 1 && !s->pPrev) FY_CUDA(cudaMalloc((void**)&s->pPrev, (size_t)N * sizeof(double)));
    int corrTotal = 0;
    bool pMatrixValid = false;          // the pencil copy of the pEqn matrix + its DIC diagonal belong to the current 1/A()
    double lastFU = 0.0;

    cudaEventRecord(ev[0], h->stream);
    FY_CUDA(cudaMemcpyAsync(s->U0, U, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    FY_CUDA(cudaMemcpyAsync(s->phi0, s->phi, (size_t)g.nSlots * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    // --- Pressure-velocity PIMPLE corrector loop   pim.C:91-105  ([OF-6 pimpleControl::loop()]: the last outer corrector is
    // the "final iteration": ...Final relaxation factors, and the pFinal solver in its last PISO corrector)
    for (int outer = 1; outer <= nOuter; ++outer) {
    const bool finalOuter = outer == nOuter;
    const double fU = (finalOuter && s->relaxUFinal > 0) ? s->relaxUFinal : s->relaxU;
    const double fP = (finalOuter && s->relaxPFinal > 0) ? s->relaxPFinal : s->relaxP;
    if (nOuter != 1) {                                   // storePrevIterFields()
        FY_CUDA(cudaMemcpyAsync(s->pPrev, p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    } else if (fP > 0 && fP < 1) {
        h->err = "fy_pimple_solve: p.relax() needs the previous-iteration field, which PIMPLE stores only when nOuterCorrectors > 1";
        return FY_ERR_INVALID;
    }
    if (outer == 1 || fU != lastFU) pMatrixValid = false; // UcEqn.relax() changes A(), hence the pEqn matrix
    lastFU = fU;
    FV_LAUNCH(k_grad_vector, G, g, U, h->dField[FY_F_VGRAD]);
    FV_LAUNCH(k_pim_div_dev, G, g, s->nu, alpha, U, h->dField[FY_F_VGRAD], s->divDev);
    // alphaPhic = alphacf*phic is built once per time step, before the loop (pim.C:85): phi0 in every outer corrector
    FV_LAUNCH(k_pim_assemble_U, G, g, s->nu, rDeltaT, s->phi0, alpha, alpha0, s->U0, h->dField[FY_F_USOURCEDRAG], s->divDev,
              fU, U, s->diagU, s->loU, s->upU, s->srcU, s->dgU, s->rAU);
    FV_LAUNCH(k_pim_forces, G, g, gvec[0], gvec[1], gvec[2], s->rAU, h->dField[FY_F_USOURCE], s->phicForces);
    if (ctl.momentumPredictor) {
        FV_LAUNCH(k_pim_predictor_source, G, g, s->phicForces, s->rAU, p, s->gradP);
        FV_LAUNCH(k_usolve_setup, G, g, FV_ALPHA_B * s->nu, s->phi0, s->srcU, s->gradP, U, s->bU, s->psiU);
        if ((rc = fvSmoothSetMatrix(h, s, s->loU, s->upU))) return rc;
        for (int m = 0; m < 3; ++m) {
            if (!g.valid[m]) continue;
            if ((rc = fvSmoothSolve(h, s, s->dgU + (size_t)m * N, s->bU + (size_t)m * N, s->psiU + (size_t)m * N, ctl.UTol,
                                    ctl.URelTol, ctl.maxIter, &s->stats.U[m])))
                return rc;
            FV_LAUNCH(k_store_component, G, N, s->psiU + (size_t)m * N, m, U);
        }
    }
    if (outer == 1) cudaEventRecord(ev[1], h->stream);
    for (int corr = 1; corr <= ctl.nCorrectors; ++corr) {
        cudaEventRecord(ev[2], h->stream);
        FV_LAUNCH(k_HbyA, G, g, FV_ALPHA_B * s->nu, s->phi0, s->loU, s->upU, s->srcU, U, s->rAU, s->HbyA);
        FV_LAUNCH(k_pim_phiHbyA, G, g, rDeltaT, s->HbyA, s->U0, s->phi0, s->rAU, alpha, s->phiHbyA, s->upP);
        if (g.pNeedRef) {
            FV_LAUNCH(k_adjust_sum, G, g, s->phiHbyA, s->red, s->dStep);
            FV_LAUNCH(k_adjust_scale, G, g, s->phiHbyA, s->dStep);
        }
        FV_LAUNCH(k_pim_add_forces, G, g.nSlots, s->phicForces, s->phiHbyA);
        if (s->bGradP) FV_LAUNCH(k_pim_constrain_pressure, G, g, s->phiHbyA, U, s->rAU, s->bGradP);
        cudaEventRecord(ev[3], h->stream);
        for (int nonOrth = 0; nonOrth <= ctl.nNonOrthogonalCorrectors; ++nonOrth) {
            FV_LAUNCH(k_pim_pEqn, G, g, ctl.pRefCell, ctl.pRefValue, rDeltaT, s->upP, s->phiHbyA, s->rAU, alpha, alpha0, s->dgP,
                      s->bP);
            const bool fin = finalOuter && corr == ctl.nCorrectors && nonOrth == ctl.nNonOrthogonalCorrectors;   // pimple.finalInnerIter()
            fy_solver_perf perf{0, 0, 0, 0};
            if ((rc = fvPcgSolve(h, s, s->dgP, s->upP, s->bP, p, fin ? ctl.pFinalTol : ctl.pTol,
                                 fin ? ctl.pFinalRelTol : ctl.pRelTol, ctl.maxIter, ctl.preconditioner, &perf,
                                 pMatrixValid)))              // 1/A() is fixed while the relaxation factor is: same pEqn matrix
                return rc;
            pMatrixValid = true;
            if (s->stats.nPSolves < 8) s->stats.p[s->stats.nPSolves] = perf;
            s->stats.nPSolves++;
            if (nonOrth == ctl.nNonOrthogonalCorrectors) {
                FV_LAUNCH(k_pim_flux_update, G, g, s->upP, s->phiHbyA, p, s->rAU, alpha, s->phi);
                // p.relax()   pim/pEqn.H:41 -- after phic, before the Uc correction (whose pEqn.flux() sees the relaxed p)
                if (fP > 0 && fP < 1) FV_LAUNCH(k_relax_field, G, N, fP, s->pPrev, p);
            }
        }
        cudaEventRecord(ev[4], h->stream);
        FV_LAUNCH(k_pim_correct_U, G, g, dt, rDeltaT, corrTotal++, s->phi, p, s->upP, s->phicForces, s->HbyA, s->rAU, alpha, alpha0,
                  U, s->red, s->dStep);
        cudaEventRecord(ev[5], h->stream);
        FY_CUDA(cudaEventSynchronize(ev[5]));
        cudaEventElapsedTime(&tmp, ev[2], ev[3]); msOther += tmp;
        cudaEventElapsedTime(&tmp, ev[3], ev[4]); msP += tmp;
        cudaEventElapsedTime(&tmp, ev[4], ev[5]); msOther += tmp;
    }
    }   // outer corrector
    cudaEventElapsedTime(&msMom, ev[0], ev[1]);
    h->phaseMs[6] = msMom;
    h->phaseMs[7] = msP;
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
    for (int q = 0; q < corrTotal && q < 8; ++q) s->cumulativeContErr += s->hStep->corrGlobal[q];
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
        if (pd.bcU < FV_FIXED_VALUE || pd.bcU > FV_EMPTY || pd.bcP < FV_FIXED_VALUE || pd.bcP > FV_FIXED_FLUX_PRESSURE)
            return no("patch type not supported by the device FV path (U: fixedValue / zeroGradient / empty; p: those or fixedFluxPressure)");
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
    {
        // tensorField inv() decides from cell 0's surfaceSum(SfHat*Sf) which directions to remove [OF-6 tensorField.C]
        double T[3], scale = 0;
        for (int d = 0; d < 3; ++d) {
            const int nd = d == 0 ? nx : (d == 1 ? ny : nz);
            T[d] = 0;
            if (nd > 1) T[d] += (g.Sf[d] / g.magSf[d]) * g.Sf[d];
            for (int hi = 0; hi < 2; ++hi) {
                const int sd = 2 * d + hi;
                if ((hi == 0 || nd == 1) && g.kindU[sd] != FV_EMPTY) T[d] += (g.bSf[sd] / g.bMagSf[sd]) * g.bSf[sd];
            }
            scale += T[d] * T[d];
        }
        for (int d = 0; d < 3; ++d) g.reconRm[d] = (T[d] * T[d]) / scale < FV_SMALL ? 1 : 0;
    }
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
    double** b1[] = {&s->rAU, &s->diagU, &s->dgP, &s->bP};
    for (auto pp : b1) if ((rc = devAlloc(h, pp, (size_t)N))) return rc;
    g.bGradP = nullptr;
    for (int q = 0; q < 6; ++q) s->hasFluxP = s->hasFluxP || g.kindP[q] == FV_FIXED_FLUX_PRESSURE;
    if (s->hasFluxP) {                                     // fixedFluxPressure: the patch's gradient, one value per boundary face
        if ((rc = devAlloc(h, &s->bGradP, NS))) return rc;
        FY_CUDA(cudaMemsetAsync(s->bGradP, 0, NS * sizeof(double), h->stream));
        g.bGradP = s->bGradP;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    s->cellGrid = std::min((N + BLK - 1) / BLK, sms * 8);
    if ((rc = penCreate(h, s))) return rc;
    if ((rc = devAlloc(h, &s->red.partial, 4 * (size_t)std::max(s->cellGrid, s->pen.rowGrid)))) return rc;
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
                    s->srcU, s->dgU, s->bU, s->psiU, s->upP, s->dgP, s->bP, s->stage, s->phicForces, s->divDev, s->bGradP, s->pPrev,
                    s->red.partial, s->red.ticket, s->dSolve, s->dStep};
    for (void* p : ptrs) if (p) cudaFree(p);
    penDestroy(s);
    for (auto& e : s->pev) if (e) cudaEventDestroy(e);
    if (s->hSolve) cudaFreeHost(s->hSolve);
    if (s->hStep) cudaFreeHost(s->hStep);
    delete s;
    h->fv = nullptr;
}
