// oracle/fv_oracle.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// CPU restatement ("port") of the fluid half of the reference's hot path:
//   icoFoamYade/icoFoamYade.C:65-149     (PISO time step)
//   pimpleFoamYade/pimpleFoamYade.C:60-114, UcEqn.H, pEqn.H, CourantNo.H, continuityErrs.H
// The arithmetic of every fvm:: / fvc:: / solve call in those files belongs to OpenFOAM-6
// (README.md:17), a third-party dependency that is NOT under /root/reference and is not installed
// here, so each operator below restates OpenFOAM-6's published algorithm -- plain face loops over
// LDU addressing, in OpenFOAM's own loop order -- and cites the reference call site it serves.
// Case settings (schemes / solvers / tolerances) are not in the reference either; this file pins
// the stock OpenFOAM-6 cavity set: ddt Euler; grad, div, laplacian Gauss linear (orthogonal
// correction = none on a hex box); interpolate linear; p: PCG + DIC, U: smoothSolver symGaussSeidel.
//
// PARITY STATUS: the reference holds no tests or golden vectors for this half ("parity unpinned" at
// the OpenFOAM boundary).  What pins this file instead is the solver log of the
// stock OpenFOAM cavity tutorial (20x20x1 cells, first time step), see tests/test_fv_oracle.py.
//
// Build: make -C oracle oracle  ->  oracle/_build/liboracle.so  (-O2 -ffp-contract=off: the x86-64
// wmake build of OpenFOAM has no FMA contraction).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace
{
typedef std::vector<double> dvec;
typedef std::vector<int> ivec;

enum { BC_FIXED_VALUE = 0, BC_ZERO_GRADIENT = 1, BC_EMPTY = 2, BC_FIXED_FLUX_PRESSURE = 3 };   // the last one: p only, pimpleSolve only
enum { PRECOND_DIC = 0, PRECOND_DIAGONAL = 1, PRECOND_NONE = 2 };

const double SMALL = 1e-15;     // OpenFOAM `small` (double precision build)
const double VSMALL = 1e-300;   // OpenFOAM `vSmall`

struct Patch
{
    int start = 0, n = 0;       // range in the concatenated boundary-face arrays
    int bcU = 0, bcP = 0;
    double valueU[3] = {0, 0, 0};
    double valueP = 0;
};

struct Mesh
{
    int nCells = 0, nFaces = 0, nB = 0;
    dvec V;
    ivec l, u;                  // owner ("lower") / neighbour ("upper") of internal faces
    dvec Sf, magSf, w, dc;      // [Fi][3], [Fi], linear weights, deltaCoeffs
    std::vector<Patch> patches;
    ivec bCell;                 // [nB] faceCells
    dvec bSf, bMagSf, bDc;      // [nB][3], [nB], [nB]
    ivec ownStart;              // [N+1]
    bool validCmpt[3] = {true, true, true};   // false for the direction "empty" patches remove
    mutable dvec bGradP;        // [nB] gradient of the fixedFluxPressure patches (set by constrainPressure), else 0
    // decomposed run (decomposePar): processor of every cell, empty = one domain.  The internal faces between two
    // processors are processor-patch faces there: Amul / residual / sums still see them (Pstream halo + global sums),
    // the DIC preconditioner does not -- it factorises each processor's own lduMatrix [OF-6 DICPreconditioner.C]
    ivec procOf;
};

struct SolverPerf
{
    double initialResidual = 0, finalResidual = 0;
    int nIterations = 0;
};

// ---------------------------------------------------------------------------------------------
// lduMatrix primitives [OF-6 lduMatrixATmul.C / lduMatrixOperations.C]
// ---------------------------------------------------------------------------------------------
// Amul: Apsi = diag*psi; per face: Apsi[u] += lower*psi[l]; Apsi[l] += upper*psi[u]
void Amul(const Mesh& m, const dvec& diag, const dvec& lower, const dvec& upper, const double* psi, double* Apsi)
{
    for (int c = 0; c < m.nCells; ++c) Apsi[c] = diag[c]*psi[c];
    for (int f = 0; f < m.nFaces; ++f) {
        Apsi[m.u[f]] += lower[f]*psi[m.l[f]];
        Apsi[m.l[f]] += upper[f]*psi[m.u[f]];
    }
}

// sumA: diag + sum of off-diagonals of the row
void sumA(const Mesh& m, const dvec& diag, const dvec& lower, const dvec& upper, double* s)
{
    for (int c = 0; c < m.nCells; ++c) s[c] = diag[c];
    for (int f = 0; f < m.nFaces; ++f) {
        s[m.u[f]] += lower[f];
        s[m.l[f]] += upper[f];
    }
}

// residual: rA = source - diag*psi; per face: rA[u] -= lower*psi[l]; rA[l] -= upper*psi[u]
void residual(const Mesh& m, const dvec& diag, const dvec& lower, const dvec& upper, const double* psi,
              const double* source, double* rA)
{
    for (int c = 0; c < m.nCells; ++c) rA[c] = source[c] - diag[c]*psi[c];
    for (int f = 0; f < m.nFaces; ++f) {
        rA[m.u[f]] -= lower[f]*psi[m.l[f]];
        rA[m.l[f]] -= upper[f]*psi[m.u[f]];
    }
}

// negSumDiag: Diag[l] -= Lower; Diag[u] -= Upper, face order
void negSumDiag(const Mesh& m, const dvec& lower, const dvec& upper, dvec& diag)
{
    diag.assign(m.nCells, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        diag[m.l[f]] -= lower[f];
        diag[m.u[f]] -= upper[f];
    }
}

double sumMag(const double* a, int n)
{
    double s = 0;
    for (int i = 0; i < n; ++i) s += std::fabs(a[i]);
    return s;
}
double sumProd(const double* a, const double* b, int n)
{
    double s = 0;
    for (int i = 0; i < n; ++i) s += a[i]*b[i];
    return s;
}

// lduMatrix::solver::normFactor [OF-6 lduMatrixSolver.C]:
//   tmp = sumA * gAverage(psi);  return gSum(mag(Apsi - tmp) + mag(source - tmp)) + 1e-20
double normFactor(const Mesh& m, const dvec& diag, const dvec& lower, const dvec& upper, const double* psi,
                  const double* source, const double* Apsi)
{
    dvec tmp(m.nCells);
    sumA(m, diag, lower, upper, tmp.data());
    double avg = 0;
    for (int c = 0; c < m.nCells; ++c) avg += psi[c];
    avg /= m.nCells;
    double s = 0;
    for (int c = 0; c < m.nCells; ++c) {
        const double t = tmp[c]*avg;
        s += std::fabs(Apsi[c] - t) + std::fabs(source[c] - t);
    }
    return s + 1e-20;
}

bool checkConvergence(const SolverPerf& sp, double tol, double relTol)
{
    return sp.finalResidual < tol || (relTol > 1e-20 && sp.finalResidual < relTol*sp.initialResidual);
}

// ---------------------------------------------------------------------------------------------
// PCG [OF-6 PCG.C] with DIC [OF-6 DICPreconditioner.C] / diagonal / no preconditioner.
// Serves pEqn.solve (icoFoamYade.C:125, pimpleFoamYade/pEqn.H:35).
// ---------------------------------------------------------------------------------------------
void dicReciprocalD(const Mesh& m, const dvec& diag, const dvec& upper, dvec& rD)
{
    rD = diag;
    for (int f = 0; f < m.nFaces; ++f) rD[m.u[f]] -= upper[f]*upper[f]/rD[m.l[f]];
    for (int c = 0; c < m.nCells; ++c) rD[c] = 1.0/rD[c];
}

void dicPrecondition(const Mesh& m, const dvec& upper, const dvec& rD, const double* rA, double* wA)
{
    for (int c = 0; c < m.nCells; ++c) wA[c] = rD[c]*rA[c];
    for (int f = 0; f < m.nFaces; ++f) wA[m.u[f]] -= rD[m.u[f]]*upper[f]*wA[m.l[f]];
    for (int f = m.nFaces - 1; f >= 0; --f) wA[m.l[f]] -= rD[m.l[f]]*upper[f]*wA[m.u[f]];
}

// ---------------------------------------------------------------------------------------------
// The same PCG with the work of a DECOMPOSED run spread over host threads (cpu baseline on all cores, bench.py): one
// OpenMP thread per processor (slab) does what an MPI rank of `icoFoamYade -parallel` does -- its rows of Amul, its
// part of every sum, the DIC substitutions of its own matrix -- and the faces between two slabs (the processor-patch
// faces) are applied in a short serial pass.  Needs a partition (Mesh::procOf, non-decreasing in the cell index).  The
// sums are associated per slab, so the last digits (not the algorithm) differ from pcgSolve's; it is a TIMING path and
// is never the parity checker.
// ---------------------------------------------------------------------------------------------
int g_threads = 1;

struct Slabs {
    std::vector<int> c0, c1, f0, f1;          // cell and owned-face ranges of every slab
    std::vector<int> cross;                   // faces between two slabs
};

Slabs slabsOf(const Mesh& m)
{
    Slabs S;
    int n = 0;
    for (int c = 0; c < m.nCells; ++c) n = std::max(n, m.procOf[c] + 1);
    S.c0.assign(n, m.nCells); S.c1.assign(n, 0);
    for (int c = 0; c < m.nCells; ++c) { const int r = m.procOf[c]; S.c0[r] = std::min(S.c0[r], c); S.c1[r] = std::max(S.c1[r], c + 1); }
    S.f0.resize(n); S.f1.resize(n);
    for (int r = 0; r < n; ++r) { S.f0[r] = m.ownStart[S.c0[r]]; S.f1[r] = m.ownStart[S.c1[r]]; }
    for (int f = 0; f < m.nFaces; ++f) if (m.procOf[m.l[f]] != m.procOf[m.u[f]]) S.cross.push_back(f);
    return S;
}

SolverPerf pcgSolvePar(const Mesh& m, const dvec& diag, const dvec& upper, const dvec& source, double* psi, double tol,
                       double relTol, int maxIter, int precond)
{
    const int n = m.nCells;
    const Slabs S = slabsOf(m);
    const int R = (int)S.c0.size();
    const int T = std::max(1, std::min(g_threads, R));
    SolverPerf sp;
    dvec pA(n), wA(n), rA(n), rD, part(R), part2(R);
    auto amul = [&](const double* x, double* y) {
#pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int r = 0; r < R; ++r) {
            for (int c = S.c0[r]; c < S.c1[r]; ++c) y[c] = diag[c]*x[c];
            for (int f = S.f0[r]; f < S.f1[r]; ++f) {
                if (m.procOf[m.u[f]] != r) continue;
                y[m.u[f]] += upper[f]*x[m.l[f]];
                y[m.l[f]] += upper[f]*x[m.u[f]];
            }
        }
        for (int f : S.cross) { y[m.u[f]] += upper[f]*x[m.l[f]]; y[m.l[f]] += upper[f]*x[m.u[f]]; }
    };
    auto reduce = [&](const dvec& p) { double s = 0; for (int r = 0; r < R; ++r) s += p[r]; return s; };
    amul(psi, wA.data());
    // normFactor
    dvec sA(n);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int r = 0; r < R; ++r) {
        double a = 0;
        for (int c = S.c0[r]; c < S.c1[r]; ++c) { rA[c] = source[c] - wA[c]; sA[c] = diag[c]; a += psi[c]; }
        for (int f = S.f0[r]; f < S.f1[r]; ++f) { if (m.procOf[m.u[f]] != r) continue; sA[m.u[f]] += upper[f]; sA[m.l[f]] += upper[f]; }
        part[r] = a;
    }
    for (int f : S.cross) { sA[m.u[f]] += upper[f]; sA[m.l[f]] += upper[f]; }
    const double avg = reduce(part)/n;
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int r = 0; r < R; ++r) {
        double a = 0, b = 0;
        for (int c = S.c0[r]; c < S.c1[r]; ++c) { const double t = sA[c]*avg; a += std::fabs(wA[c] - t) + std::fabs(source[c] - t); b += std::fabs(rA[c]); }
        part[r] = a; part2[r] = b;
    }
    const double nf = reduce(part) + 1e-20;
    sp.initialResidual = reduce(part2)/nf;
    sp.finalResidual = sp.initialResidual;
    double wArA = 1e20, wArAold = wArA;
    if (!checkConvergence(sp, tol, relTol)) {
        if (precond == PRECOND_DIC) {
            rD = diag;
#pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int r = 0; r < R; ++r) {
                for (int f = S.f0[r]; f < S.f1[r]; ++f) if (m.procOf[m.u[f]] == r) rD[m.u[f]] -= upper[f]*upper[f]/rD[m.l[f]];
                for (int c = S.c0[r]; c < S.c1[r]; ++c) rD[c] = 1.0/rD[c];
            }
        } else if (precond == PRECOND_DIAGONAL) { rD.resize(n); for (int c = 0; c < n; ++c) rD[c] = 1.0/diag[c]; }
        do {
            wArAold = wArA;
#pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int r = 0; r < R; ++r) {
                if (precond == PRECOND_DIC) {
                    for (int c = S.c0[r]; c < S.c1[r]; ++c) wA[c] = rD[c]*rA[c];
                    for (int f = S.f0[r]; f < S.f1[r]; ++f) if (m.procOf[m.u[f]] == r) wA[m.u[f]] -= rD[m.u[f]]*upper[f]*wA[m.l[f]];
                    for (int f = S.f1[r] - 1; f >= S.f0[r]; --f) if (m.procOf[m.u[f]] == r) wA[m.l[f]] -= rD[m.l[f]]*upper[f]*wA[m.u[f]];
                } else if (precond == PRECOND_DIAGONAL) { for (int c = S.c0[r]; c < S.c1[r]; ++c) wA[c] = rD[c]*rA[c]; }
                else { for (int c = S.c0[r]; c < S.c1[r]; ++c) wA[c] = rA[c]; }
                double a = 0;
                for (int c = S.c0[r]; c < S.c1[r]; ++c) a += wA[c]*rA[c];
                part[r] = a;
            }
            wArA = reduce(part);
            const double beta = wArA/wArAold;
            const bool first = sp.nIterations == 0;
#pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int r = 0; r < R; ++r)
                for (int c = S.c0[r]; c < S.c1[r]; ++c) pA[c] = first ? wA[c] : wA[c] + beta*pA[c];
            amul(pA.data(), wA.data());
#pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int r = 0; r < R; ++r) { double a = 0; for (int c = S.c0[r]; c < S.c1[r]; ++c) a += wA[c]*pA[c]; part[r] = a; }
            const double wApA = reduce(part);
            if (std::fabs(wApA)/nf < VSMALL) break;
            const double alpha = wArA/wApA;
#pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int r = 0; r < R; ++r) {
                double a = 0;
                for (int c = S.c0[r]; c < S.c1[r]; ++c) { psi[c] += alpha*pA[c]; rA[c] -= alpha*wA[c]; a += std::fabs(rA[c]); }
                part[r] = a;
            }
            sp.finalResidual = reduce(part)/nf;
        } while (sp.nIterations++ < maxIter && !checkConvergence(sp, tol, relTol));
    }
    return sp;
}

SolverPerf pcgSolve(const Mesh& m, const dvec& diag, const dvec& upper, const dvec& source, double* psi, double tol,
                    double relTol, int maxIter, int precond)
{
    if (g_threads > 1 && !m.procOf.empty()) return pcgSolvePar(m, diag, upper, source, psi, tol, relTol, maxIter, precond);
    const int n = m.nCells;
    SolverPerf sp;
    dvec pA(n), wA(n), rA(n);
    double wArA = 1e20, wArAold = wArA;     // solverPerformance::great_
    Amul(m, diag, upper, upper, psi, wA.data());
    for (int c = 0; c < n; ++c) rA[c] = source[c] - wA[c];
    const double nf = normFactor(m, diag, upper, upper, psi, source.data(), wA.data());
    sp.initialResidual = sumMag(rA.data(), n)/nf;
    sp.finalResidual = sp.initialResidual;
    if (!checkConvergence(sp, tol, relTol)) {
        dvec rD;
        dvec upperPre;                                   // the coefficients the preconditioner sees
        const dvec* upP = &upper;
        if (precond == PRECOND_DIC && !m.procOf.empty()) {
            upperPre = upper;
            for (int f = 0; f < m.nFaces; ++f)
                if (m.procOf[m.l[f]] != m.procOf[m.u[f]]) upperPre[f] = 0.0;
            upP = &upperPre;
        }
        if (precond == PRECOND_DIC) dicReciprocalD(m, diag, *upP, rD);
        else if (precond == PRECOND_DIAGONAL) { rD.resize(n); for (int c = 0; c < n; ++c) rD[c] = 1.0/diag[c]; }
        do {
            wArAold = wArA;
            if (precond == PRECOND_DIC) dicPrecondition(m, *upP, rD, rA.data(), wA.data());
            else if (precond == PRECOND_DIAGONAL) { for (int c = 0; c < n; ++c) wA[c] = rD[c]*rA[c]; }
            else { for (int c = 0; c < n; ++c) wA[c] = rA[c]; }
            wArA = sumProd(wA.data(), rA.data(), n);
            if (sp.nIterations == 0) {
                for (int c = 0; c < n; ++c) pA[c] = wA[c];
            } else {
                const double beta = wArA/wArAold;
                for (int c = 0; c < n; ++c) pA[c] = wA[c] + beta*pA[c];
            }
            Amul(m, diag, upper, upper, pA.data(), wA.data());
            const double wApA = sumProd(wA.data(), pA.data(), n);
            if (std::fabs(wApA)/nf < VSMALL) break;       // checkSingularity
            const double alpha = wArA/wApA;
            for (int c = 0; c < n; ++c) {
                psi[c] += alpha*pA[c];
                rA[c] -= alpha*wA[c];
            }
            sp.finalResidual = sumMag(rA.data(), n)/nf;
        } while (sp.nIterations++ < maxIter && !checkConvergence(sp, tol, relTol));
    }
    return sp;
}

// ---------------------------------------------------------------------------------------------
// smoothSolver + symGaussSeidel [OF-6 smoothSolver.C, symGaussSeidelSmoother.C], nSweeps = 1.
// Serves solve(UEqn == -grad p) (icoFoamYade.C:93, pimpleFoamYade/UcEqn.H:24), one component at a time.
// ---------------------------------------------------------------------------------------------
void symGaussSeidelSweep(const Mesh& m, const dvec& diag, const dvec& lower, const dvec& upper, const double* source,
                         double* psi)
{
    const int n = m.nCells;
    dvec bPrime(source, source + n);
    for (int c = 0; c < n; ++c) {
        double psii = bPrime[c];
        for (int f = m.ownStart[c]; f < m.ownStart[c + 1]; ++f) psii -= upper[f]*psi[m.u[f]];
        psii /= diag[c];
        for (int f = m.ownStart[c]; f < m.ownStart[c + 1]; ++f) bPrime[m.u[f]] -= lower[f]*psii;
        psi[c] = psii;
    }
    for (int c = n - 1; c >= 0; --c) {
        double psii = bPrime[c];
        for (int f = m.ownStart[c]; f < m.ownStart[c + 1]; ++f) psii -= upper[f]*psi[m.u[f]];
        psii /= diag[c];
        for (int f = m.ownStart[c]; f < m.ownStart[c + 1]; ++f) bPrime[m.u[f]] -= lower[f]*psii;
        psi[c] = psii;
    }
}

SolverPerf smoothSolve(const Mesh& m, const dvec& diag, const dvec& lower, const dvec& upper, const dvec& source,
                       double* psi, double tol, double relTol, int maxIter)
{
    const int n = m.nCells;
    SolverPerf sp;
    dvec Apsi(n), r(n);
    Amul(m, diag, lower, upper, psi, Apsi.data());
    const double nf = normFactor(m, diag, lower, upper, psi, source.data(), Apsi.data());
    for (int c = 0; c < n; ++c) r[c] = source[c] - Apsi[c];
    sp.initialResidual = sumMag(r.data(), n)/nf;
    sp.finalResidual = sp.initialResidual;
    if (!checkConvergence(sp, tol, relTol)) {
        do {
            symGaussSeidelSweep(m, diag, lower, upper, source.data(), psi);
            residual(m, diag, lower, upper, psi, source.data(), r.data());
            sp.finalResidual = sumMag(r.data(), n)/nf;
        } while ((sp.nIterations += 1) < maxIter && !checkConvergence(sp, tol, relTol));
    }
    return sp;
}

// ---------------------------------------------------------------------------------------------
// boundary values
// ---------------------------------------------------------------------------------------------
inline void patchU(const Mesh& m, const Patch& p, int b, const double* U, double* out)
{
    if (p.bcU == BC_FIXED_VALUE) { out[0] = p.valueU[0]; out[1] = p.valueU[1]; out[2] = p.valueU[2]; }
    else { const int c = m.bCell[b]; out[0] = U[3*(size_t)c]; out[1] = U[3*(size_t)c + 1]; out[2] = U[3*(size_t)c + 2]; }
}
inline double patchP(const Mesh& m, const Patch& p, int b, const double* P)
{
    if (p.bcP == BC_FIXED_FLUX_PRESSURE) return P[m.bCell[b]] + m.bGradP[b]/m.bDc[b];      // fixedGradient: p_P + gradient/deltaCoeffs
    return p.bcP == BC_FIXED_VALUE ? p.valueP : P[m.bCell[b]];
}

// ---------------------------------------------------------------------------------------------
// fvc operators [OF-6 gaussGrad.C, surfaceInterpolationScheme.C, fvcSurfaceIntegrate.C]
// ---------------------------------------------------------------------------------------------
// linear interpolate: lambda*(phiP - phiN) + phiN
inline double lerp(double w, double P, double N) { return w*(P - N) + N; }

// fvc::grad(U): (1/V) sum_f Sf (x) U_f  -> [N][9] row-major T_ij = Sf_i U_j.   icoFoamYade.C:71, pimpleFoamYade.C:76
void gradVector(const Mesh& m, const double* U, double* g)
{
    std::fill(g, g + 9*(size_t)m.nCells, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        const int P = m.l[f], N = m.u[f];
        double uf[3];
        for (int j = 0; j < 3; ++j) uf[j] = lerp(m.w[f], U[3*(size_t)P + j], U[3*(size_t)N + j]);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double t = m.Sf[3*(size_t)f + i]*uf[j];
                g[9*(size_t)P + 3*i + j] += t;
                g[9*(size_t)N + 3*i + j] -= t;
            }
    }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            double ub[3];
            patchU(m, p, b, U, ub);
            const int c = m.bCell[b];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) g[9*(size_t)c + 3*i + j] += m.bSf[3*(size_t)b + i]*ub[j];
        }
    }
    for (int c = 0; c < m.nCells; ++c)
        for (int k = 0; k < 9; ++k) g[9*(size_t)c + k] /= m.V[c];
}

// fvc::grad(p) -> [N][3].   icoFoamYade.C:93,136; pimpleFoamYade.C:74
void gradScalar(const Mesh& m, const double* p, double* g)
{
    std::fill(g, g + 3*(size_t)m.nCells, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        const int P = m.l[f], N = m.u[f];
        const double pf = lerp(m.w[f], p[P], p[N]);
        for (int i = 0; i < 3; ++i) {
            const double t = m.Sf[3*(size_t)f + i]*pf;
            g[3*(size_t)P + i] += t;
            g[3*(size_t)N + i] -= t;
        }
    }
    for (const Patch& pt : m.patches) {
        if (pt.bcP == BC_EMPTY) continue;
        for (int b = pt.start; b < pt.start + pt.n; ++b) {
            const double pb = patchP(m, pt, b, p);
            const int c = m.bCell[b];
            for (int i = 0; i < 3; ++i) g[3*(size_t)c + i] += m.bSf[3*(size_t)b + i]*pb;
        }
    }
    for (int c = 0; c < m.nCells; ++c)
        for (int k = 0; k < 3; ++k) g[3*(size_t)c + k] /= m.V[c];
}

// fvc::div(phi) = surfaceIntegrate: (1/V) sum_f +-phi_f.   icoFoamYade.C:120 and continuityErrs.H
void divFlux(const Mesh& m, const double* phi, double* d)
{
    std::fill(d, d + m.nCells, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        d[m.l[f]] += phi[f];
        d[m.u[f]] -= phi[f];
    }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) d[m.bCell[b]] += phi[m.nFaces + b];
    }
    for (int c = 0; c < m.nCells; ++c) d[c] /= m.V[c];
}

// fvc::div(phi, U) = surfaceIntegrate(phi_f * interpolate(U))  -> [N][3]
// [OF-6 gaussConvectionScheme.C fvcDiv, linear].   pimpleFoamYade.C:73 (second term of ddtU_f)
void divPhiU(const Mesh& m, const double* phi, const double* U, double* d)
{
    std::fill(d, d + 3*(size_t)m.nCells, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        const int P = m.l[f], N = m.u[f];
        for (int j = 0; j < 3; ++j) {
            const double t = phi[f]*lerp(m.w[f], U[3*(size_t)P + j], U[3*(size_t)N + j]);
            d[3*(size_t)P + j] += t;
            d[3*(size_t)N + j] -= t;
        }
    }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            double ub[3];
            patchU(m, p, b, U, ub);
            const int c = m.bCell[b];
            for (int j = 0; j < 3; ++j) d[3*(size_t)c + j] += phi[m.nFaces + b]*ub[j];
        }
    }
    for (int c = 0; c < m.nCells; ++c)
        for (int j = 0; j < 3; ++j) d[3*(size_t)c + j] /= m.V[c];
}

// fvc::laplacian(gamma, U) = surfaceIntegrate((interpolate(gamma)*magSf) * snGrad(U))  -> [N][3]
// [OF-6 gaussLaplacianScheme.C fvcLaplacian + uncorrectedSnGrad / snGradScheme::snGrad: deltaCoeffs*(U_N - U_P);
//  fixedValue patch snGrad = deltaCoeffs_b*(U_b - U_P), zeroGradient patch snGrad = 0; the non-orthogonal
//  correction of "corrected" is identically zero on the hex box].  gammaB = value of gamma on every non-empty patch:
//  pimpleFoamYade's alphac has `calculated` patches (pim/createFields.H: built from a dimensionedScalar) that hold the
//  1.0 FoamYade::initFields assigns field-wide (F.C:67); FoamYade only ever writes cell values afterwards.
//  pimpleFoamYade.C:75 (divT = 2 nu fvc::laplacian(alphac, Uc))
void laplacianGammaU(const Mesh& m, const double* gamma, double gammaB, const double* U, double* d)
{
    std::fill(d, d + 3*(size_t)m.nCells, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        const int P = m.l[f], N = m.u[f];
        const double gm = lerp(m.w[f], gamma[P], gamma[N])*m.magSf[f];
        for (int j = 0; j < 3; ++j) {
            const double t = gm*(m.dc[f]*(U[3*(size_t)N + j] - U[3*(size_t)P + j]));
            d[3*(size_t)P + j] += t;
            d[3*(size_t)N + j] -= t;
        }
    }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            const int c = m.bCell[b];
            const double gm = gammaB*m.bMagSf[b];
            for (int j = 0; j < 3; ++j) {
                const double sn = p.bcU == BC_FIXED_VALUE ? m.bDc[b]*(p.valueU[j] - U[3*(size_t)c + j]) : 0.0;
                d[3*(size_t)c + j] += gm*sn;
            }
        }
    }
    for (int c = 0; c < m.nCells; ++c)
        for (int j = 0; j < 3; ++j) d[3*(size_t)c + j] /= m.V[c];
}

// CourantNo.H (stock; icoFoamYade.C:68 / pimpleFoamYade/CourantNo.H:32-49): sumPhi = surfaceSum(mag(phi))
void courant(const Mesh& m, const double* phi, double dt, double* CoNum, double* meanCoNum)
{
    dvec s(m.nCells, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        const double a = std::fabs(phi[f]);
        s[m.l[f]] += a;
        s[m.u[f]] += a;
    }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) s[m.bCell[b]] += std::fabs(phi[m.nFaces + b]);
    }
    double mx = -1e300, sum = 0, sv = 0;
    for (int c = 0; c < m.nCells; ++c) {
        mx = std::max(mx, s[c]/m.V[c]);
        sum += s[c];
        sv += m.V[c];
    }
    *CoNum = 0.5*mx*dt;
    *meanCoNum = 0.5*(sum/sv)*dt;
}

// ---------------------------------------------------------------------------------------------
// the icoFoamYade state and time step
// ---------------------------------------------------------------------------------------------
struct Ctl
{
    int nCorrectors = 2;
    int nNonOrthCorrectors = 0;
    int momentumPredictor = 1;
    int pRefCell = 0;
    double pRefValue = 0;
    double pTol = 1e-6, pRelTol = 0.05, pFinalTol = 1e-6, pFinalRelTol = 0.0;
    double UTol = 1e-5, URelTol = 0.0;
    int maxIter = 1000;
    int precond = PRECOND_DIC;
    // PIMPLE sub-dictionary + relaxationFactors (pimpleFoamYade only).  A factor <= 0 means "no entry in fvSolution":
    // fvMatrix::relax() / GeometricField::relax() are then no-ops [OF-6 fvMatrix.C relax(), GeometricField.C relax()]
    int nOuterCorrectors = 1;
    double relaxU = 0, relaxUFinal = 0, relaxP = 0, relaxPFinal = 0;
};

struct Stats
{
    double CoNum = 0, meanCoNum = 0;
    SolverPerf U[3];
    SolverPerf p[8];            // one per pressure solve of the step, in order
    int nPSolves = 0;
    double sumLocalContErr = 0, globalContErr = 0, cumulativeContErr = 0;
    double corrSumLocal[8] = {0}, corrGlobal[8] = {0};   // continuityErrs.H after each PISO corrector
};

struct Ico
{
    Mesh m;
    double nu = 0.01;
    dvec U, p, phi;             // [N][3], [N], [Fi + nB]
    dvec uSource;               // [N][3]
    dvec vGrad;                 // [N][9]
    Ctl ctl;
    Stats st;
    double cumulativeContErr = 0;
    // intermediates of the last corrector, kept for stage-by-stage parity tests
    dvec rAU, HbyA, phiHbyA, gradP, diagU, upperU, lowerU, sourceU, diagP, upperP, sourceP;
    dvec icU, bcU;              // UEqn internalCoeffs / boundaryCoeffs [nB][3]
    struct Pim* pim = nullptr;  // pimpleFoamYade intermediates (allocated on first use)
    double tMomentum = 0, tPressure = 0, tOther = 0;
};

bool pNeedsReference(const Mesh& m)
{
    for (const Patch& p : m.patches)
        if (p.bcP == BC_FIXED_VALUE && p.n > 0) return false;
    return true;
}

double nowSec()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// UEqn = fvm::ddt(U) + fvm::div(phi,U) - fvm::laplacian(nu,U) == uSource          icoFoamYade.C:79-85
// [OF-6 EulerDdtScheme::fvmDdt, gaussConvectionScheme::fvmDiv, gaussLaplacianScheme::fvmLaplacianUncorrected,
//  fvMatrix operator+, operator-, operator==(fvMatrix, volField)]
void assembleUEqn(Ico& s, double dt, const dvec& U0)
{
    const Mesh& m = s.m;
    const int N = m.nCells, Fi = m.nFaces, nB = m.nB;
    const double rDeltaT = 1.0/dt;
    dvec diagD(N), lowerC(Fi), upperC(Fi), diagC, upperL(Fi), diagL;
    s.sourceU.assign(3*(size_t)N, 0.0);
    for (int c = 0; c < N; ++c) {
        diagD[c] = rDeltaT*m.V[c];
        for (int j = 0; j < 3; ++j) s.sourceU[3*(size_t)c + j] = rDeltaT*U0[3*(size_t)c + j]*m.V[c];
    }
    for (int f = 0; f < Fi; ++f) {
        lowerC[f] = -m.w[f]*s.phi[f];
        upperC[f] = lowerC[f] + s.phi[f];
        upperL[f] = m.dc[f]*(s.nu*m.magSf[f]);
    }
    negSumDiag(m, lowerC, upperC, diagC);
    negSumDiag(m, upperL, upperL, diagL);
    s.diagU.resize(N);
    s.upperU.resize(Fi);
    s.lowerU.resize(Fi);
    for (int c = 0; c < N; ++c) s.diagU[c] = (diagD[c] + diagC[c]) - diagL[c];
    for (int f = 0; f < Fi; ++f) {
        s.upperU[f] = upperC[f] - upperL[f];
        s.lowerU[f] = lowerC[f] - upperL[f];
    }
    // boundary coefficients: convection  ic = phi_b*valueInternalCoeffs, bc = -phi_b*valueBoundaryCoeffs;
    //                        laplacian   ic = gammaMagSf_b*gradientInternalCoeffs, bc = -gammaMagSf_b*gradientBoundaryCoeffs
    s.icU.assign(3*(size_t)nB, 0.0);
    s.bcU.assign(3*(size_t)nB, 0.0);
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            const double phib = s.phi[Fi + b];
            const double gMagSf = s.nu*m.bMagSf[b];
            for (int j = 0; j < 3; ++j) {
                double icC, bcC, icL, bcL;
                if (p.bcU == BC_FIXED_VALUE) {
                    icC = phib*0.0;
                    bcC = -phib*p.valueU[j];
                    icL = gMagSf*(-1.0*m.bDc[b]);
                    bcL = -gMagSf*(m.bDc[b]*p.valueU[j]);
                } else {
                    icC = phib*1.0;
                    bcC = -phib*0.0;
                    icL = gMagSf*0.0;
                    bcL = -gMagSf*0.0;
                }
                s.icU[3*(size_t)b + j] = icC - icL;
                s.bcU[3*(size_t)b + j] = bcC - bcL;
            }
        }
    }
    // == uSource : source += V*uSource
    for (int c = 0; c < N; ++c)
        for (int j = 0; j < 3; ++j) s.sourceU[3*(size_t)c + j] += m.V[c]*s.uSource[3*(size_t)c + j];
}

// solve(UEqn == -fvc::grad(p))  icoFoamYade.C:93  [OF-6 fvMatrix<Type>::solveSegregated]
void momentumPredictor(Ico& s)
{
    const Mesh& m = s.m;
    const int N = m.nCells;
    s.gradP.resize(3*(size_t)N);
    gradScalar(m, s.p.data(), s.gradP.data());
    dvec src(s.sourceU);
    for (int c = 0; c < N; ++c)
        for (int j = 0; j < 3; ++j) src[3*(size_t)c + j] += m.V[c]*(-s.gradP[3*(size_t)c + j]);
    for (int b = 0; b < m.nB; ++b)          // addBoundarySource (patch order, face order)
        for (int j = 0; j < 3; ++j) src[3*(size_t)m.bCell[b] + j] += s.bcU[3*(size_t)b + j];
    dvec psi(N), b1(N), dg(N);
    for (int j = 0; j < 3; ++j) {
        s.st.U[j] = SolverPerf();
        if (!m.validCmpt[j]) continue;
        dg = s.diagU;
        for (int b = 0; b < m.nB; ++b) dg[m.bCell[b]] += s.icU[3*(size_t)b + j];      // addBoundaryDiag
        for (int c = 0; c < N; ++c) { psi[c] = s.U[3*(size_t)c + j]; b1[c] = src[3*(size_t)c + j]; }
        s.st.U[j] = smoothSolve(m, dg, s.lowerU, s.upperU, b1, psi.data(), s.ctl.UTol, s.ctl.URelTol, s.ctl.maxIter);
        for (int c = 0; c < N; ++c) s.U[3*(size_t)c + j] = psi[c];
    }
}

// rAU = 1/UEqn.A();  H = UEqn.H()       icoFoamYade.C:99-100  [OF-6 fvMatrix::A, fvMatrix::H, lduMatrix::H]
void computeRAUandH(Ico& s, dvec& H)
{
    const Mesh& m = s.m;
    const int N = m.nCells;
    dvec D(s.diagU);
    for (int b = 0; b < m.nB; ++b) {
        const double* ic = &s.icU[3*(size_t)b];
        D[m.bCell[b]] += (ic[0] + ic[1] + ic[2])/3.0;                                  // addCmptAvBoundaryDiag
    }
    s.rAU.resize(N);
    for (int c = 0; c < N; ++c) s.rAU[c] = 1.0/(D[c]/m.V[c]);
    H.assign(3*(size_t)N, 0.0);
    dvec bd(N);
    for (int j = 0; j < 3; ++j) {
        if (!m.validCmpt[j]) continue;
        std::fill(bd.begin(), bd.end(), 0.0);
        for (int b = 0; b < m.nB; ++b) bd[m.bCell[b]] += s.icU[3*(size_t)b + j];
        for (int c = 0; c < N; ++c) bd[c] = -bd[c];
        for (int b = 0; b < m.nB; ++b) {
            const double* ic = &s.icU[3*(size_t)b];
            bd[m.bCell[b]] += (ic[0] + ic[1] + ic[2])/3.0;
        }
        for (int c = 0; c < N; ++c) H[3*(size_t)c + j] = bd[c]*s.U[3*(size_t)c + j];
    }
    dvec Hl(3*(size_t)N, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        const int P = m.l[f], Nb = m.u[f];
        for (int j = 0; j < 3; ++j) {
            Hl[3*(size_t)Nb + j] -= s.lowerU[f]*s.U[3*(size_t)P + j];
            Hl[3*(size_t)P + j] -= s.upperU[f]*s.U[3*(size_t)Nb + j];
        }
    }
    for (size_t i = 0; i < 3*(size_t)N; ++i) H[i] += Hl[i] + s.sourceU[i];
    for (int b = 0; b < m.nB; ++b)
        for (int j = 0; j < 3; ++j) H[3*(size_t)m.bCell[b] + j] += s.bcU[3*(size_t)b + j];
    for (int c = 0; c < N; ++c)
        for (int j = 0; j < 3; ++j) H[3*(size_t)c + j] /= m.V[c];
    for (int j = 0; j < 3; ++j)
        if (!m.validCmpt[j]) for (int c = 0; c < N; ++c) H[3*(size_t)c + j] = 0.0;
}

// adjustPhi(phiHbyA, U, p)  icoFoamYade.C:108  [OF-6 adjustPhi.C]
int adjustPhi(const Mesh& m, double* phi)
{
    if (!pNeedsReference(m)) return 0;
    double massIn = 0, fixedMassOut = 0, adjustableMassOut = 0;
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            const double ph = phi[m.nFaces + b];
            if (p.bcU == BC_FIXED_VALUE) { if (ph < 0) massIn -= ph; else fixedMassOut += ph; }
            else { if (ph < 0) massIn -= ph; else adjustableMassOut += ph; }
        }
    }
    double totalFlux = VSMALL;                                   // vSmall + sum(mag(phi)): gSum of the INTERNAL field
    for (int f = 0; f < m.nFaces; ++f) totalFlux += std::fabs(phi[f]);
    double massCorr = 1.0;
    const double magAdj = std::fabs(adjustableMassOut);
    if (magAdj > VSMALL && magAdj/totalFlux > SMALL) massCorr = (massIn - fixedMassOut)/adjustableMassOut;
    else if (std::fabs(fixedMassOut - massIn)/totalFlux > 1e-8) return -1;   // FatalError in OpenFOAM
    for (const Patch& p : m.patches) {
        if (p.bcU != BC_ZERO_GRADIENT) continue;
        for (int b = p.start; b < p.start + p.n; ++b)
            if (phi[m.nFaces + b] > 0.0) phi[m.nFaces + b] *= massCorr;
    }
    return 0;
}

int icoSolve(Ico& s, double dt)
{
    const Mesh& m = s.m;
    const int N = m.nCells, Fi = m.nFaces, nB = m.nB;
    const double rDeltaT = 1.0/dt;
    double t0 = nowSec();
    const dvec U0(s.U), phi0(s.phi);                 // oldTime fields
    s.st.nPSolves = 0;
    assembleUEqn(s, dt, U0);
    if (s.ctl.momentumPredictor) momentumPredictor(s);
    s.tMomentum += nowSec() - t0;

    dvec H, div(N), totalSource(N), dg(N), icP(nB), bcP(nB), rAUf(Fi);
    s.HbyA.resize(3*(size_t)N);
    s.phiHbyA.resize((size_t)Fi + nB);
    s.upperP.resize(Fi);
    for (int corr = 1; corr <= s.ctl.nCorrectors; ++corr) {
        t0 = nowSec();
        computeRAUandH(s, H);
        for (int c = 0; c < N; ++c)
            for (int j = 0; j < 3; ++j) s.HbyA[3*(size_t)c + j] = s.rAU[c]*H[3*(size_t)c + j];
        // phiHbyA = fvc::flux(HbyA) + fvc::interpolate(rAU)*fvc::ddtCorr(U, phi)      icoFoamYade.C:101-106
        // [OF-6 EulerDdtScheme::fvcDdtPhiCorr + ddtScheme::fvcDdtPhiCoeff]
        for (int f = 0; f < Fi; ++f) {
            const int P = m.l[f], Nb = m.u[f];
            double flux = 0, u0f = 0;
            for (int j = 0; j < 3; ++j) {
                flux += m.Sf[3*(size_t)f + j]*lerp(m.w[f], s.HbyA[3*(size_t)P + j], s.HbyA[3*(size_t)Nb + j]);
                u0f += m.Sf[3*(size_t)f + j]*lerp(m.w[f], U0[3*(size_t)P + j], U0[3*(size_t)Nb + j]);
            }
            const double phiCorr = phi0[f] - u0f;
            const double coeff = 1.0 - std::min(std::fabs(phiCorr)/(std::fabs(phi0[f]) + SMALL), 1.0);
            rAUf[f] = lerp(m.w[f], s.rAU[P], s.rAU[Nb]);
            s.phiHbyA[f] = flux + rAUf[f]*((coeff*rDeltaT)*phiCorr);
        }
        for (const Patch& p : m.patches) {
            for (int b = p.start; b < p.start + p.n; ++b) {
                if (p.bcU == BC_EMPTY) { s.phiHbyA[Fi + b] = 0.0; continue; }
                const int c = m.bCell[b];
                double hb[3], u0b[3];
                if (p.bcU == BC_FIXED_VALUE) { hb[0] = p.valueU[0]; hb[1] = p.valueU[1]; hb[2] = p.valueU[2]; }   // constrainHbyA
                else { hb[0] = s.HbyA[3*(size_t)c]; hb[1] = s.HbyA[3*(size_t)c + 1]; hb[2] = s.HbyA[3*(size_t)c + 2]; }
                patchU(m, p, b, U0.data(), u0b);
                double flux = 0, u0f = 0;
                for (int j = 0; j < 3; ++j) { flux += m.bSf[3*(size_t)b + j]*hb[j]; u0f += m.bSf[3*(size_t)b + j]*u0b[j]; }
                const double phiCorr = phi0[Fi + b] - u0f;
                double coeff = 1.0 - std::min(std::fabs(phiCorr)/(std::fabs(phi0[Fi + b]) + SMALL), 1.0);
                if (p.bcU == BC_FIXED_VALUE) coeff = 0.0;                     // fixesValue() patches
                s.phiHbyA[Fi + b] = flux + s.rAU[c]*((coeff*rDeltaT)*phiCorr);
            }
        }
        if (adjustPhi(m, s.phiHbyA.data()) != 0) return -1;
        // constrainPressure: fixedFluxPressure patches only -- none in the pinned BC set
        s.tOther += nowSec() - t0;
        for (int nonOrth = 0; nonOrth <= s.ctl.nNonOrthCorrectors; ++nonOrth) {
            t0 = nowSec();
            // pEqn: fvm::laplacian(rAU, p) == fvc::div(phiHbyA)                    icoFoamYade.C:118-121
            for (int f = 0; f < Fi; ++f) s.upperP[f] = m.dc[f]*(rAUf[f]*m.magSf[f]);
            negSumDiag(m, s.upperP, s.upperP, s.diagP);
            for (const Patch& p : m.patches) {
                for (int b = p.start; b < p.start + p.n; ++b) {
                    icP[b] = 0.0;
                    bcP[b] = 0.0;
                    if (p.bcP != BC_FIXED_VALUE) continue;
                    const double pGamma = s.rAU[m.bCell[b]]*m.bMagSf[b];
                    icP[b] = pGamma*(-1.0*m.bDc[b]);
                    bcP[b] = -pGamma*(m.bDc[b]*p.valueP);
                }
            }
            divFlux(m, s.phiHbyA.data(), div.data());
            s.sourceP.assign(N, 0.0);
            for (int c = 0; c < N; ++c) s.sourceP[c] += m.V[c]*div[c];
            if (pNeedsReference(m)) {                                               // icoFoamYade.C:123
                s.sourceP[s.ctl.pRefCell] += s.diagP[s.ctl.pRefCell]*s.ctl.pRefValue;
                s.diagP[s.ctl.pRefCell] += s.diagP[s.ctl.pRefCell];
            }
            // fvMatrix<scalar>::solveSegregated
            dg = s.diagP;
            totalSource = s.sourceP;
            for (int b = 0; b < nB; ++b) { dg[m.bCell[b]] += icP[b]; totalSource[m.bCell[b]] += bcP[b]; }
            const bool fin = (corr == s.ctl.nCorrectors) && (nonOrth == s.ctl.nNonOrthCorrectors);
            SolverPerf sp = pcgSolve(m, dg, s.upperP, totalSource, s.p.data(), fin ? s.ctl.pFinalTol : s.ctl.pTol,
                                     fin ? s.ctl.pFinalRelTol : s.ctl.pRelTol, s.ctl.maxIter, s.ctl.precond);
            if (s.st.nPSolves < 8) s.st.p[s.st.nPSolves] = sp;
            s.st.nPSolves++;
            s.tPressure += nowSec() - t0;
            if (nonOrth == s.ctl.nNonOrthCorrectors) {
                // phi = phiHbyA - pEqn.flux()                                        icoFoamYade.C:127-130
                for (int f = 0; f < Fi; ++f)
                    s.phi[f] = s.phiHbyA[f] - (s.upperP[f]*s.p[m.u[f]] - s.upperP[f]*s.p[m.l[f]]);
                for (int b = 0; b < nB; ++b) s.phi[Fi + b] = s.phiHbyA[Fi + b] - (icP[b]*s.p[m.bCell[b]] - bcP[b]);
            }
        }
        t0 = nowSec();
        // continuityErrs.H (stock)                                                 icoFoamYade.C:134
        divFlux(m, s.phi.data(), div.data());
        double sl = 0, sg = 0, sv = 0;
        for (int c = 0; c < N; ++c) { sl += std::fabs(div[c])*m.V[c]; sg += div[c]*m.V[c]; sv += m.V[c]; }
        s.st.sumLocalContErr = dt*(sl/sv);
        s.st.globalContErr = dt*(sg/sv);
        s.cumulativeContErr += s.st.globalContErr;
        s.st.cumulativeContErr = s.cumulativeContErr;
        if (corr <= 8) { s.st.corrSumLocal[corr - 1] = s.st.sumLocalContErr; s.st.corrGlobal[corr - 1] = s.st.globalContErr; }
        // U = HbyA - rAU*fvc::grad(p)                                              icoFoamYade.C:136-137
        s.gradP.resize(3*(size_t)N);
        gradScalar(m, s.p.data(), s.gradP.data());
        for (int c = 0; c < N; ++c)
            for (int j = 0; j < 3; ++j) s.U[3*(size_t)c + j] = s.HbyA[3*(size_t)c + j] - s.rAU[c]*s.gradP[3*(size_t)c + j];
        s.tOther += nowSec() - t0;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// pimpleFoamYade: UcEqn.H + pEqn.H + continuityErrs.H  (pimpleFoamYade.C:82-104 with nOuterCorrectors 1, laminar)
//
// PARITY UNPINNED: the reference ships no case, log or test for this solver and OpenFOAM is not installed, so this
// restates OpenFOAM-6's operators from their definitions (each cited) without a golden vector behind it.  Pinned
// case settings (the reference has none): Euler; Gauss linear everywhere; laplacian/snGrad uncorrected (orthogonal
// box); no relaxationFactors (UcEqn.relax() and p.relax() are then no-ops); simulationType laminar -> Stokes model,
// divDevRhoReff(U) = - fvc::div((alpha*nuEff)*dev2(T(fvc::grad(U)))) - fvm::laplacian(alpha*nuEff, U)
// [OF-6 linearViscousStress.C]; no fixedFluxPressure patch (constrainPressure is a no-op).
//
// Boundary values that matter: alphac / uSource / uSourceDrag have `calculated` patches (pim/createFields.H builds
// them from dimensioned constants).  alphac's hold 1.0 (F.C:67 assigns the whole field), uSource's hold 0; FoamYade
// writes cell values only and correctBoundaryConditions() of a calculated patch does nothing (pim.C:83-88).
//
// Old-time fields: Uc.oldTime() is stored by fvc::ddt(Uc) at pim.C:73 (before the step changes Uc) and phic.oldTime()
// by the first phic assignment, i.e. both are the previous step's, as in icoFoam.  alphac.oldTime() is stored at
// pim.C:83 (correctBoundaryConditions -> storeOldTimes), AFTER FoamYade wrote this step's void fraction into the
// cells through operator[] -- so alphac.oldTime() == alphac and fvc::ddt(alphac) is identically +0.  The arithmetic
// shape is kept (alpha0 is passed separately) so a driver that stores old times differently can say so.
// ---------------------------------------------------------------------------------------------
struct Pim
{
    dvec alphaf;                // alphacf [Fi + nB]
    dvec alphaPhi;              // alphaPhic
    dvec rAUf;                  // rAUcf
    dvec phicForces;            // [Fi + nB]
    dvec recon;                 // last fvc::reconstruct result [N][3]
    dvec divDev;                // fvc::div((alpha nu) dev2(T(grad U))) [N][3]
    dvec spDiv;                 // fvc::ddt(alphac) + fvc::div(alphaPhic) [N]
    dvec invT;                  // inv(surfaceSum(SfHat*Sf)) [N][9]
};

inline double det9(const double* t)
{
    return (t[0]*t[4]*t[8] + t[1]*t[5]*t[6] + t[2]*t[3]*t[7]) - (t[0]*t[5]*t[7] + t[1]*t[3]*t[8] + t[2]*t[4]*t[6]);
}
// [OF-6 TensorI.H inv(const Tensor&, const Cmpt dett)]
inline void inv9(const double* t, double* r)
{
    const double d = det9(t);
    r[0] = (t[4]*t[8] - t[7]*t[5])/d; r[1] = (t[2]*t[7] - t[1]*t[8])/d; r[2] = (t[1]*t[5] - t[2]*t[4])/d;
    r[3] = (t[6]*t[5] - t[3]*t[8])/d; r[4] = (t[0]*t[8] - t[2]*t[6])/d; r[5] = (t[3]*t[2] - t[0]*t[5])/d;
    r[6] = (t[3]*t[7] - t[4]*t[6])/d; r[7] = (t[1]*t[6] - t[0]*t[7])/d; r[8] = (t[0]*t[4] - t[3]*t[1])/d;
}

// inv(surfaceSum(SfHat*mesh.Sf())) with tensorField inv()'s removal of the directions an empty patch pair leaves
// without faces [OF-6 fvcReconstruct.C, tensorField.C inv(Field<tensor>&, const UList<tensor>&)]
void reconstructTensor(const Mesh& m, dvec& invT)
{
    const int N = m.nCells;
    dvec T(9*(size_t)N, 0.0);
    for (int f = 0; f < m.nFaces; ++f)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double t = (m.Sf[3*(size_t)f + i]/m.magSf[f])*m.Sf[3*(size_t)f + j];
                T[9*(size_t)m.l[f] + 3*i + j] += t;
                T[9*(size_t)m.u[f] + 3*i + j] += t;
            }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    T[9*(size_t)m.bCell[b] + 3*i + j] += (m.bSf[3*(size_t)b + i]/m.bMagSf[b])*m.bSf[3*(size_t)b + j];
    }
    invT.assign(9*(size_t)N, 0.0);
    if (N == 0) return;
    double scale = 0;
    for (int k = 0; k < 9; ++k) scale += T[k]*T[k];
    bool rm[3];
    for (int d = 0; d < 3; ++d) rm[d] = (T[4*d]*T[4*d])/scale < SMALL;
    for (int c = 0; c < N; ++c) {
        double t[9];
        for (int k = 0; k < 9; ++k) t[k] = T[9*(size_t)c + k];
        for (int d = 0; d < 3; ++d) if (rm[d]) t[4*d] += 1.0;
        inv9(t, &invT[9*(size_t)c]);
        for (int d = 0; d < 3; ++d) if (rm[d]) invT[9*(size_t)c + 4*d] -= 1.0;
    }
}

// fvc::reconstruct(ssf) = inv(surfaceSum(SfHat*Sf)) & surfaceSum(SfHat*ssf)     [OF-6 fvcReconstruct.C]
void reconstruct(const Mesh& m, const dvec& invT, const double* ssf, double* out)
{
    const int N = m.nCells;
    dvec v(3*(size_t)N, 0.0);
    for (int f = 0; f < m.nFaces; ++f)
        for (int i = 0; i < 3; ++i) {
            const double t = (m.Sf[3*(size_t)f + i]/m.magSf[f])*ssf[f];
            v[3*(size_t)m.l[f] + i] += t;
            v[3*(size_t)m.u[f] + i] += t;
        }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b)
            for (int i = 0; i < 3; ++i) v[3*(size_t)m.bCell[b] + i] += (m.bSf[3*(size_t)b + i]/m.bMagSf[b])*ssf[m.nFaces + b];
    }
    for (int c = 0; c < N; ++c) {
        const double* t = &invT[9*(size_t)c];
        const double* a = &v[3*(size_t)c];
        for (int i = 0; i < 3; ++i) out[3*(size_t)c + i] = t[3*i]*a[0] + t[3*i + 1]*a[1] + t[3*i + 2]*a[2];
    }
}

// fvc::snGrad(p)*mesh.magSf()  [OF-6 snGradScheme::snGrad; fixedValue patch: deltaCoeffs*(p_b - p_P), zeroGradient: 0]
void snGradPMagSf(const Mesh& m, const double* p, double* out)
{
    for (int f = 0; f < m.nFaces; ++f) out[f] = (m.dc[f]*(p[m.u[f]] - p[m.l[f]]))*m.magSf[f];
    for (const Patch& pt : m.patches)
        for (int b = pt.start; b < pt.start + pt.n; ++b) {
            double sn = 0.0;
            if (pt.bcP == BC_FIXED_VALUE) sn = m.bDc[b]*(pt.valueP - p[m.bCell[b]]);
            if (pt.bcP == BC_FIXED_FLUX_PRESSURE) sn = m.bGradP[b];
            out[m.nFaces + b] = pt.bcP == BC_EMPTY ? 0.0 : sn*m.bMagSf[b];
        }
}

// fvc::div((alpha*nuEff)*dev2(T(fvc::grad(U))))  [OF-6 linearViscousStress::divDevRhoReff, gaussDivScheme::fvcDiv,
// gaussGrad::correctBoundaryConditions for the patch values of grad(U), TensorI.H dev2 / T]
void divDevTerm(const Mesh& m, const double* alpha, double alphaB, double nu, const double* U, double* out)
{
    const int N = m.nCells;
    dvec g(9*(size_t)N), X(9*(size_t)N);
    gradVector(m, U, g.data());
    auto dev2T = [](const double* gr, double a, double* x) {
        double t[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) t[3*i + j] = gr[3*j + i];
        const double tr = t[0] + t[4] + t[8];
        const double sph = (2.0/3.0)*tr;
        t[0] -= sph; t[4] -= sph; t[8] -= sph;
        for (int k = 0; k < 9; ++k) x[k] = a*t[k];
    };
    for (int c = 0; c < N; ++c) dev2T(&g[9*(size_t)c], alpha[c]*nu, &X[9*(size_t)c]);
    std::fill(out, out + 3*(size_t)N, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        const int P = m.l[f], Nb = m.u[f];
        for (int j = 0; j < 3; ++j) {
            double t = 0;
            for (int i = 0; i < 3; ++i) t += m.Sf[3*(size_t)f + i]*lerp(m.w[f], X[9*(size_t)P + 3*i + j], X[9*(size_t)Nb + 3*i + j]);
            out[3*(size_t)P + j] += t;
            out[3*(size_t)Nb + j] -= t;
        }
    }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            const int c = m.bCell[b];
            double n[3], sn[3], gb[9], nG[3], xb[9];
            for (int i = 0; i < 3; ++i) n[i] = m.bSf[3*(size_t)b + i]/m.bMagSf[b];
            for (int j = 0; j < 3; ++j)
                sn[j] = p.bcU == BC_FIXED_VALUE ? m.bDc[b]*(p.valueU[j] - U[3*(size_t)c + j]) : 0.0;
            for (int k = 0; k < 9; ++k) gb[k] = g[9*(size_t)c + k];
            for (int j = 0; j < 3; ++j) nG[j] = n[0]*gb[j] + n[1]*gb[3 + j] + n[2]*gb[6 + j];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) gb[3*i + j] += n[i]*(sn[j] - nG[j]);
            dev2T(gb, alphaB*nu, xb);
            for (int j = 0; j < 3; ++j) {
                double t = 0;
                for (int i = 0; i < 3; ++i) t += m.bSf[3*(size_t)b + i]*xb[3*i + j];
                out[3*(size_t)c + j] += t;
            }
        }
    }
    for (int c = 0; c < N; ++c)
        for (int j = 0; j < 3; ++j) out[3*(size_t)c + j] /= m.V[c];
}

// UcEqn  (pim/UcEqn.H:3-11)
void assembleUcEqn(Ico& s, Pim& q, double dt, const dvec& U0, const dvec& phiA, const double* alpha, const double* alpha0,
                   double alphaB, const double* uSourceDrag)
{
    const Mesh& m = s.m;
    const int N = m.nCells, Fi = m.nFaces, nB = m.nB;
    const double rDeltaT = 1.0/dt;
    // alphacf = fvc::interpolate(alphac); alphaPhic = alphacf*phic            pim.C:84-85
    q.alphaf.resize((size_t)Fi + nB);
    q.alphaPhi.resize((size_t)Fi + nB);
    for (int f = 0; f < Fi; ++f) q.alphaf[f] = lerp(m.w[f], alpha[m.l[f]], alpha[m.u[f]]);
    for (int b = 0; b < nB; ++b) q.alphaf[Fi + b] = alphaB;
    for (int f = 0; f < Fi + nB; ++f) q.alphaPhi[f] = q.alphaf[f]*phiA[f];   // phiA: phic as it stood at pim.C:85
    // fvm::ddt(alphac, Uc)   [OF-6 EulerDdtScheme::fvmDdt(alpha, vf)]
    dvec diagD(N), lowerC(Fi), upperC(Fi), diagC, upperL(Fi), diagL;
    s.sourceU.assign(3*(size_t)N, 0.0);
    for (int c = 0; c < N; ++c) {
        diagD[c] = (rDeltaT*alpha[c])*m.V[c];
        for (int j = 0; j < 3; ++j) s.sourceU[3*(size_t)c + j] = ((rDeltaT*alpha0[c])*U0[3*(size_t)c + j])*m.V[c];
    }
    // fvm::div(alphaPhic, Uc), fvm::laplacian(alpha*nuEff, Uc)
    for (int f = 0; f < Fi; ++f) {
        lowerC[f] = -m.w[f]*q.alphaPhi[f];
        upperC[f] = lowerC[f] + q.alphaPhi[f];
        const double gf = lerp(m.w[f], alpha[m.l[f]]*s.nu, alpha[m.u[f]]*s.nu);
        upperL[f] = m.dc[f]*(gf*m.magSf[f]);
    }
    negSumDiag(m, lowerC, upperC, diagC);
    negSumDiag(m, upperL, upperL, diagL);
    // fvm::Sp(fvc::ddt(alphac) + fvc::div(alphaPhic), Uc)
    q.spDiv.resize(N);
    divFlux(m, q.alphaPhi.data(), q.spDiv.data());
    for (int c = 0; c < N; ++c) q.spDiv[c] = rDeltaT*(alpha[c] - alpha0[c]) + q.spDiv[c];
    // ((ddt + div) - Sp) + divDevRhoReff == Sp(uSourceDrag)
    s.diagU.resize(N);
    s.upperU.resize(Fi);
    s.lowerU.resize(Fi);
    for (int c = 0; c < N; ++c)
        s.diagU[c] = (((diagD[c] + diagC[c]) - m.V[c]*q.spDiv[c]) + (-diagL[c])) - m.V[c]*uSourceDrag[c];
    for (int f = 0; f < Fi; ++f) {
        s.upperU[f] = upperC[f] + (-upperL[f]);
        s.lowerU[f] = lowerC[f] + (-upperL[f]);
    }
    s.icU.assign(3*(size_t)nB, 0.0);
    s.bcU.assign(3*(size_t)nB, 0.0);
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            const double phib = q.alphaPhi[Fi + b];
            const double gMagSf = (alphaB*s.nu)*m.bMagSf[b];
            for (int j = 0; j < 3; ++j) {
                double icC, bcC, icL, bcL;
                if (p.bcU == BC_FIXED_VALUE) {
                    icC = phib*0.0;
                    bcC = -phib*p.valueU[j];
                    icL = gMagSf*(-1.0*m.bDc[b]);
                    bcL = -gMagSf*(m.bDc[b]*p.valueU[j]);
                } else {
                    icC = phib*1.0;
                    bcC = -phib*0.0;
                    icL = gMagSf*0.0;
                    bcL = -gMagSf*0.0;
                }
                s.icU[3*(size_t)b + j] = icC + (-icL);
                s.bcU[3*(size_t)b + j] = bcC + (-bcL);
            }
        }
    }
    // explicit part of divDevRhoReff: the matrix gets  source -= V*(-fvc::div(...))
    q.divDev.resize(3*(size_t)N);
    divDevTerm(m, alpha, alphaB, s.nu, s.U.data(), q.divDev.data());
    for (int c = 0; c < N; ++c)
        for (int j = 0; j < 3; ++j) s.sourceU[3*(size_t)c + j] -= m.V[c]*(-q.divDev[3*(size_t)c + j]);
}

// UcEqn.relax()   pim/UcEqn.H:13  [OF-6 fvMatrix.C relax(const scalar alpha)]: make the matrix diagonally dominant, divide the
// diagonal by alpha and move the difference, times the CURRENT field, to the source.  The non-coupled patches' internal
// coefficients count with their largest-magnitude component while dominance is enforced and are taken out again with
// their smallest component.  alpha <= 0: no relaxation factor in fvSolution -> not called at all.
void relaxUcEqn(Ico& s, double alpha)
{
    if (alpha <= 0) return;
    const Mesh& m = s.m;
    const int N = m.nCells, Fi = m.nFaces;
    dvec& D = s.diagU;
    const dvec D0(D);
    dvec sumOff(N, 0.0);
    for (int f = 0; f < Fi; ++f) {                                  // lduMatrix::sumMagOffDiag
        sumOff[m.u[f]] += std::fabs(s.lowerU[f]);
        sumOff[m.l[f]] += std::fabs(s.upperU[f]);
    }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;                            // (an empty patch field has size 0)
        for (int b = p.start; b < p.start + p.n; ++b) {
            const double* ic = &s.icU[3*(size_t)b];
            D[m.bCell[b]] += std::max(std::max(std::fabs(ic[0]), std::fabs(ic[1])), std::fabs(ic[2]));   // cmptMax(cmptMag(iCoeffs))
        }
    }
    for (int c = 0; c < N; ++c) D[c] = std::max(std::fabs(D[c]), sumOff[c]);
    for (int c = 0; c < N; ++c) D[c] /= alpha;
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            const double* ic = &s.icU[3*(size_t)b];
            D[m.bCell[b]] -= std::min(std::min(ic[0], ic[1]), ic[2]);                                     // cmptMin(iCoeffs)
        }
    }
    for (int c = 0; c < N; ++c)
        for (int j = 0; j < 3; ++j) s.sourceU[3*(size_t)c + j] += (D[c] - D0[c])*s.U[3*(size_t)c + j];
}

int pimpleSolve(Ico& s, Pim& q, double dt, const double* alpha, const double* alpha0, const double* uSourceDrag,
                const double* gvec)
{
    const Mesh& m = s.m;
    const int N = m.nCells, Fi = m.nFaces, nB = m.nB, nF = Fi + nB;
    const double rDeltaT = 1.0/dt, alphaB = 1.0;
    double t0 = nowSec();
    const dvec U0(s.U), phi0(s.phi);
    s.st.nPSolves = 0;
    if (q.invT.size() != 9*(size_t)N) reconstructTensor(m, q.invT);
    const int nOuter = std::max(1, s.ctl.nOuterCorrectors);
    dvec pPrev;
    int corrTotal = 0;
    // --- Pressure-velocity PIMPLE corrector loop   pim.C:91-105  [OF-6 pimpleControl::loop(): corr = 1..nOuterCorrectors;
    // on the last one data::finalIteration is set (-> the ...Final relaxation factors and, in the last PISO corrector, the
    // pFinal solver); prevIter fields are stored only when nOuterCorrectors != 1]
    for (int outer = 1; outer <= nOuter; ++outer) {
    const bool finalOuter = outer == nOuter;
    const double fU = (finalOuter && s.ctl.relaxUFinal > 0) ? s.ctl.relaxUFinal : s.ctl.relaxU;
    const double fP = (finalOuter && s.ctl.relaxPFinal > 0) ? s.ctl.relaxPFinal : s.ctl.relaxP;
    if (nOuter != 1) pPrev = s.p;                                   // storePrevIterFields()
    else if (fP > 0 && fP < 1) return -2;                           // p.relax() without a stored prevIter: FatalError in OpenFOAM
    t0 = nowSec();
    // alphaPhic is built ONCE per time step, before the loop (pim.C:85): every outer corrector convects with the old phic
    assembleUcEqn(s, q, dt, U0, phi0, alpha, alpha0, alphaB, uSourceDrag);
    relaxUcEqn(s, fU);                                              // UcEqn.relax()   pim/UcEqn.H:13

    // rAUc = 1/UcEqn.A(); rAUcf = fvc::interpolate(rAUc)       (A() does not depend on Uc: computed once)
    dvec H;
    computeRAUandH(s, H);
    q.rAUf.resize(nF);
    for (int f = 0; f < Fi; ++f) q.rAUf[f] = lerp(m.w[f], s.rAU[m.l[f]], s.rAU[m.u[f]]);
    for (int b = 0; b < nB; ++b) q.rAUf[Fi + b] = s.rAU[m.bCell[b]];
    // phicForces = fvc::flux(rAUc*uSource) + rAUcf*(g & Sf)                   pim/UcEqn.H:17-20
    q.phicForces.assign(nF, 0.0);
    for (int f = 0; f < Fi; ++f) {
        const int P = m.l[f], Nb = m.u[f];
        double flux = 0, gS = 0;
        for (int j = 0; j < 3; ++j) {
            flux += m.Sf[3*(size_t)f + j]*lerp(m.w[f], s.rAU[P]*s.uSource[3*(size_t)P + j], s.rAU[Nb]*s.uSource[3*(size_t)Nb + j]);
            gS += gvec[j]*m.Sf[3*(size_t)f + j];
        }
        q.phicForces[f] = flux + q.rAUf[f]*gS;
    }
    for (const Patch& p : m.patches) {
        if (p.bcU == BC_EMPTY) continue;
        for (int b = p.start; b < p.start + p.n; ++b) {
            double flux = 0, gS = 0;
            for (int j = 0; j < 3; ++j) {
                flux += m.bSf[3*(size_t)b + j]*(s.rAU[m.bCell[b]]*0.0);
                gS += gvec[j]*m.bSf[3*(size_t)b + j];
            }
            q.phicForces[Fi + b] = flux + q.rAUf[Fi + b]*gS;
        }
    }
    dvec ssf(nF), snG(nF);
    q.recon.resize(3*(size_t)N);
    if (s.ctl.momentumPredictor) {
        // solve(UcEqn == fvc::reconstruct(phicForces/rAUcf - fvc::snGrad(p)*mesh.magSf()))   pim/UcEqn.H:22-33
        snGradPMagSf(m, s.p.data(), snG.data());
        for (int f = 0; f < nF; ++f) ssf[f] = q.phicForces[f]/q.rAUf[f] - snG[f];
        reconstruct(m, q.invT, ssf.data(), q.recon.data());
        dvec src(s.sourceU);
        for (int c = 0; c < N; ++c)
            for (int j = 0; j < 3; ++j) src[3*(size_t)c + j] += m.V[c]*q.recon[3*(size_t)c + j];
        for (int b = 0; b < nB; ++b)
            for (int j = 0; j < 3; ++j) src[3*(size_t)m.bCell[b] + j] += s.bcU[3*(size_t)b + j];
        dvec psi(N), b1(N), dg(N);
        for (int j = 0; j < 3; ++j) {
            s.st.U[j] = SolverPerf();
            if (!m.validCmpt[j]) continue;
            dg = s.diagU;
            for (int b = 0; b < nB; ++b) dg[m.bCell[b]] += s.icU[3*(size_t)b + j];
            for (int c = 0; c < N; ++c) { psi[c] = s.U[3*(size_t)c + j]; b1[c] = src[3*(size_t)c + j]; }
            s.st.U[j] = smoothSolve(m, dg, s.lowerU, s.upperU, b1, psi.data(), s.ctl.UTol, s.ctl.URelTol, s.ctl.maxIter);
            for (int c = 0; c < N; ++c) s.U[3*(size_t)c + j] = psi[c];
        }
    }
    s.tMomentum += nowSec() - t0;

    dvec div(N), totalSource(N), dg(N), icP(nB), bcP(nB), aphi(nF), fluxByA(nF);
    s.HbyA.resize(3*(size_t)N);
    s.phiHbyA.resize(nF);
    s.upperP.resize(Fi);
    for (int corr = 1; corr <= s.ctl.nCorrectors; ++corr) {
        t0 = nowSec();
        // HbyA = constrainHbyA(rAUc*UcEqn.H(), Uc, p)                                       pim/pEqn.H:2
        computeRAUandH(s, H);
        for (int c = 0; c < N; ++c)
            for (int j = 0; j < 3; ++j) s.HbyA[3*(size_t)c + j] = s.rAU[c]*H[3*(size_t)c + j];
        // phiHbyA = fvc::flux(HbyA) + alphacf*rAUcf*fvc::ddtCorr(Uc, phic)                  pim/pEqn.H:4-11
        for (int f = 0; f < Fi; ++f) {
            const int P = m.l[f], Nb = m.u[f];
            double flux = 0, u0f = 0;
            for (int j = 0; j < 3; ++j) {
                flux += m.Sf[3*(size_t)f + j]*lerp(m.w[f], s.HbyA[3*(size_t)P + j], s.HbyA[3*(size_t)Nb + j]);
                u0f += m.Sf[3*(size_t)f + j]*lerp(m.w[f], U0[3*(size_t)P + j], U0[3*(size_t)Nb + j]);
            }
            const double phiCorr = phi0[f] - u0f;
            const double coeff = 1.0 - std::min(std::fabs(phiCorr)/(std::fabs(phi0[f]) + SMALL), 1.0);
            s.phiHbyA[f] = flux + (q.alphaf[f]*q.rAUf[f])*((coeff*rDeltaT)*phiCorr);
        }
        for (const Patch& p : m.patches) {
            for (int b = p.start; b < p.start + p.n; ++b) {
                if (p.bcU == BC_EMPTY) { s.phiHbyA[Fi + b] = 0.0; continue; }
                const int c = m.bCell[b];
                double hb[3], u0b[3];
                if (p.bcU == BC_FIXED_VALUE) { hb[0] = p.valueU[0]; hb[1] = p.valueU[1]; hb[2] = p.valueU[2]; }
                else { hb[0] = s.HbyA[3*(size_t)c]; hb[1] = s.HbyA[3*(size_t)c + 1]; hb[2] = s.HbyA[3*(size_t)c + 2]; }
                patchU(m, p, b, U0.data(), u0b);
                double flux = 0, u0f = 0;
                for (int j = 0; j < 3; ++j) { flux += m.bSf[3*(size_t)b + j]*hb[j]; u0f += m.bSf[3*(size_t)b + j]*u0b[j]; }
                const double phiCorr = phi0[Fi + b] - u0f;
                double coeff = 1.0 - std::min(std::fabs(phiCorr)/(std::fabs(phi0[Fi + b]) + SMALL), 1.0);
                if (p.bcU == BC_FIXED_VALUE) coeff = 0.0;
                s.phiHbyA[Fi + b] = flux + (q.alphaf[Fi + b]*q.rAUf[Fi + b])*((coeff*rDeltaT)*phiCorr);
            }
        }
        if (adjustPhi(m, s.phiHbyA.data()) != 0) return -1;                                  // pim/pEqn.H:13-16
        for (int f = 0; f < nF; ++f) s.phiHbyA[f] += q.phicForces[f];                        // pim/pEqn.H:18
        // constrainPressure(p, Uc, phiHbyA, rAUcf)   pim/pEqn.H:21  [OF-6 constrainPressure.C: on fixedFluxPressure
        // patches snGrad(p) = (phiHbyA_b - Sf_b & U_b)/(magSf_b rAUcf_b), so that the corrected flux equals the wall's]
        for (const Patch& p : m.patches) {
            if (p.bcP != BC_FIXED_FLUX_PRESSURE) continue;
            for (int b = p.start; b < p.start + p.n; ++b) {
                double ub[3], su = 0;
                patchU(m, p, b, s.U.data(), ub);
                for (int j = 0; j < 3; ++j) su += m.bSf[3*(size_t)b + j]*ub[j];
                m.bGradP[b] = (s.phiHbyA[Fi + b] - su)/(m.bMagSf[b]*q.rAUf[Fi + b]);
            }
        }
        s.tOther += nowSec() - t0;
        for (int nonOrth = 0; nonOrth <= s.ctl.nNonOrthCorrectors; ++nonOrth) {
            t0 = nowSec();
            // fvm::laplacian(alphacf*rAUcf, p) == fvc::ddt(alphac) + fvc::div(alphacf*phiHbyA)   pim/pEqn.H:26-31
            for (int f = 0; f < Fi; ++f) s.upperP[f] = m.dc[f]*((q.alphaf[f]*q.rAUf[f])*m.magSf[f]);
            negSumDiag(m, s.upperP, s.upperP, s.diagP);
            for (const Patch& p : m.patches) {
                for (int b = p.start; b < p.start + p.n; ++b) {
                    icP[b] = 0.0;
                    bcP[b] = 0.0;
                    const double pGamma = (q.alphaf[Fi + b]*q.rAUf[Fi + b])*m.bMagSf[b];
                    if (p.bcP == BC_FIXED_FLUX_PRESSURE) {                     // fixedGradient: gradientInternalCoeffs 0,
                        icP[b] = pGamma*0.0;                                   // gradientBoundaryCoeffs = gradient
                        bcP[b] = -pGamma*m.bGradP[b];
                    }
                    if (p.bcP != BC_FIXED_VALUE) continue;
                    icP[b] = pGamma*(-1.0*m.bDc[b]);
                    bcP[b] = -pGamma*(m.bDc[b]*p.valueP);
                }
            }
            for (int f = 0; f < nF; ++f) aphi[f] = q.alphaf[f]*s.phiHbyA[f];
            divFlux(m, aphi.data(), div.data());
            s.sourceP.assign(N, 0.0);
            for (int c = 0; c < N; ++c) s.sourceP[c] += m.V[c]*(rDeltaT*(alpha[c] - alpha0[c]) + div[c]);
            if (pNeedsReference(m)) {
                s.sourceP[s.ctl.pRefCell] += s.diagP[s.ctl.pRefCell]*s.ctl.pRefValue;
                s.diagP[s.ctl.pRefCell] += s.diagP[s.ctl.pRefCell];
            }
            dg = s.diagP;
            totalSource = s.sourceP;
            for (int b = 0; b < nB; ++b) { dg[m.bCell[b]] += icP[b]; totalSource[m.bCell[b]] += bcP[b]; }
            const bool fin = finalOuter && (corr == s.ctl.nCorrectors) && (nonOrth == s.ctl.nNonOrthCorrectors);   // pimple.finalInnerIter()
            SolverPerf sp = pcgSolve(m, dg, s.upperP, totalSource, s.p.data(), fin ? s.ctl.pFinalTol : s.ctl.pTol,
                                     fin ? s.ctl.pFinalRelTol : s.ctl.pRelTol, s.ctl.maxIter, s.ctl.precond);
            if (s.st.nPSolves < 8) s.st.p[s.st.nPSolves] = sp;
            s.st.nPSolves++;
            s.tPressure += nowSec() - t0;
            if (nonOrth == s.ctl.nNonOrthCorrectors) {
                t0 = nowSec();
                // phic = phiHbyA - pEqn.flux()/alphacf                                       pim/pEqn.H:39
                for (int f = 0; f < Fi; ++f) fluxByA[f] = (s.upperP[f]*s.p[m.u[f]] - s.upperP[f]*s.p[m.l[f]])/q.alphaf[f];
                for (int b = 0; b < nB; ++b) fluxByA[Fi + b] = (icP[b]*s.p[m.bCell[b]] - bcP[b])/q.alphaf[Fi + b];
                for (int f = 0; f < nF; ++f) s.phi[f] = s.phiHbyA[f] - fluxByA[f];
                // p.relax()   pim/pEqn.H:41  [OF-6 GeometricField::relax(alpha): if (alpha < 1) p == prevIter + alpha*(p - prevIter)];
                // the pEqn.flux() of the Uc correction below is evaluated with the RELAXED p (fvMatrix::flux() reads psi)
                if (fP > 0 && fP < 1) {
                    for (int c = 0; c < N; ++c) s.p[c] = pPrev[c] + fP*(s.p[c] - pPrev[c]);
                    for (int f = 0; f < Fi; ++f) fluxByA[f] = (s.upperP[f]*s.p[m.u[f]] - s.upperP[f]*s.p[m.l[f]])/q.alphaf[f];
                    for (int b = 0; b < nB; ++b) fluxByA[Fi + b] = (icP[b]*s.p[m.bCell[b]] - bcP[b])/q.alphaf[Fi + b];
                }
                // Uc = HbyA + rAUc*fvc::reconstruct((phicForces - pEqn.flux()/alphacf)/rAUcf)  pim/pEqn.H:43-45
                for (int f = 0; f < nF; ++f) ssf[f] = (q.phicForces[f] - fluxByA[f])/q.rAUf[f];
                reconstruct(m, q.invT, ssf.data(), q.recon.data());
                for (int c = 0; c < N; ++c)
                    for (int j = 0; j < 3; ++j)
                        s.U[3*(size_t)c + j] = s.HbyA[3*(size_t)c + j] + s.rAU[c]*q.recon[3*(size_t)c + j];
                s.tOther += nowSec() - t0;
            }
        }
        t0 = nowSec();
        // continuityErrs.H: contErr = fvc::ddt(alphac) + fvc::div(alphacf*phic)         pim/continuityErrs.H:32-46
        for (int f = 0; f < nF; ++f) aphi[f] = q.alphaf[f]*s.phi[f];
        divFlux(m, aphi.data(), div.data());
        double sl = 0, sg = 0, sv = 0;
        for (int c = 0; c < N; ++c) {
            const double e = rDeltaT*(alpha[c] - alpha0[c]) + div[c];
            sl += std::fabs(e)*m.V[c]; sg += e*m.V[c]; sv += m.V[c];
        }
        s.st.sumLocalContErr = dt*(sl/sv);
        s.st.globalContErr = dt*(sg/sv);
        s.cumulativeContErr += s.st.globalContErr;
        s.st.cumulativeContErr = s.cumulativeContErr;
        if (corrTotal < 8) { s.st.corrSumLocal[corrTotal] = s.st.sumLocalContErr; s.st.corrGlobal[corrTotal] = s.st.globalContErr; }
        corrTotal++;
        s.tOther += nowSec() - t0;
    }
    }   // outer corrector
    return 0;
}


Mesh* buildMesh(int nCells, const double* V, int nFaces, const int* owner, const int* neigh, const double* Sf,
                const double* magSf, const double* w, const double* dc, int nPatches, const int* patchSize,
                const int* bCell, const double* bSf, const double* bMagSf, const double* bDc, const int* bcU,
                const double* valueU, const int* bcP, const double* valueP, Mesh& m)
{
    m.nCells = nCells;
    m.nFaces = nFaces;
    m.V.assign(V, V + nCells);
    m.l.assign(owner, owner + nFaces);
    m.u.assign(neigh, neigh + nFaces);
    m.Sf.assign(Sf, Sf + 3*(size_t)nFaces);
    m.magSf.assign(magSf, magSf + nFaces);
    m.w.assign(w, w + nFaces);
    m.dc.assign(dc, dc + nFaces);
    int o = 0;
    for (int i = 0; i < nPatches; ++i) {
        Patch p;
        p.start = o;
        p.n = patchSize[i];
        p.bcU = bcU[i];
        p.bcP = bcP[i];
        for (int j = 0; j < 3; ++j) p.valueU[j] = valueU[3*i + j];
        p.valueP = valueP[i];
        m.patches.push_back(p);
        o += p.n;
    }
    m.nB = o;
    m.bCell.assign(bCell, bCell + o);
    m.bSf.assign(bSf, bSf + 3*(size_t)o);
    m.bMagSf.assign(bMagSf, bMagSf + o);
    m.bDc.assign(bDc, bDc + o);
    m.bGradP.assign(o, 0.0);
    m.ownStart.assign(nCells + 1, 0);
    for (int f = 0; f < nFaces; ++f) m.ownStart[owner[f] + 1]++;
    for (int c = 0; c < nCells; ++c) m.ownStart[c + 1] += m.ownStart[c];
    for (const Patch& p : m.patches) {
        if (p.bcU != BC_EMPTY || p.n == 0) continue;
        const double* s = &m.bSf[3*(size_t)p.start];
        int d = 0;
        if (std::fabs(s[1]) > std::fabs(s[d])) d = 1;
        if (std::fabs(s[2]) > std::fabs(s[d])) d = 2;
        m.validCmpt[d] = false;
    }
    return &m;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// C API (ctypes: oracle/port.py)
// ---------------------------------------------------------------------------------------------
extern "C" {

void* fvo_create(int nCells, const double* V, int nFaces, const int* owner, const int* neigh, const double* Sf,
                 const double* magSf, const double* w, const double* dc, int nPatches, const int* patchSize,
                 const int* bCell, const double* bSf, const double* bMagSf, const double* bDc, const int* bcU,
                 const double* valueU, const int* bcP, const double* valueP)
{
    Ico* s = new Ico();
    buildMesh(nCells, V, nFaces, owner, neigh, Sf, magSf, w, dc, nPatches, patchSize, bCell, bSf, bMagSf, bDc, bcU,
              valueU, bcP, valueP, s->m);
    const Mesh& m = s->m;
    s->U.assign(3*(size_t)nCells, 0.0);
    s->p.assign(nCells, 0.0);
    s->phi.assign((size_t)nFaces + m.nB, 0.0);
    s->uSource.assign(3*(size_t)nCells, 0.0);
    s->vGrad.assign(9*(size_t)nCells, 0.0);
    return s;
}
// decomposed-run semantics of the pressure preconditioner: procOf [nCells] (NULL = one domain)
void fvo_set_partition(void* h, const int* procOf)
{
    Ico* s = (Ico*)h;
    Mesh& m = s->m;
    if (procOf) m.procOf.assign(procOf, procOf + m.nCells);
    else m.procOf.clear();
}

// host threads of the timing path (pcgSolvePar); 1 = the sequential checker
void fvo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

void fvo_destroy(void* h)
{
    Ico* s = (Ico*)h;
    delete s->pim;
    delete s;
}

// ctl: nCorrectors nNonOrth momentumPredictor pRefCell maxIter precond | pRefValue pTol pRelTol pFinalTol pFinalRelTol UTol URelTol nu
void fvo_set_controls(void* h, const int* ic6, const double* dc8)
{
    Ico* s = (Ico*)h;
    s->ctl.nCorrectors = ic6[0]; s->ctl.nNonOrthCorrectors = ic6[1]; s->ctl.momentumPredictor = ic6[2];
    s->ctl.pRefCell = ic6[3]; s->ctl.maxIter = ic6[4]; s->ctl.precond = ic6[5];
    s->ctl.pRefValue = dc8[0]; s->ctl.pTol = dc8[1]; s->ctl.pRelTol = dc8[2]; s->ctl.pFinalTol = dc8[3];
    s->ctl.pFinalRelTol = dc8[4]; s->ctl.UTol = dc8[5]; s->ctl.URelTol = dc8[6]; s->nu = dc8[7];
}

// PIMPLE nOuterCorrectors + relaxationFactors { equations { U, UFinal } fields { p, pFinal } }; a factor <= 0 = no entry
void fvo_set_pimple_controls(void* h, int nOuter, const double* r4)
{
    Ico* s = (Ico*)h;
    s->ctl.nOuterCorrectors = nOuter;
    s->ctl.relaxU = r4[0]; s->ctl.relaxUFinal = r4[1]; s->ctl.relaxP = r4[2]; s->ctl.relaxPFinal = r4[3];
}

// name: U p phi uSource vGrad rAU HbyA phiHbyA gradP diagU upperU lowerU sourceU diagP upperP sourceP
double* fvo_field(void* h, const char* name, long* n)
{
    Ico* s = (Ico*)h;
    const std::string k(name);
    dvec* v = nullptr;
    if (k == "U") v = &s->U; else if (k == "p") v = &s->p; else if (k == "phi") v = &s->phi;
    else if (k == "uSource") v = &s->uSource; else if (k == "vGrad") v = &s->vGrad; else if (k == "rAU") v = &s->rAU;
    else if (k == "HbyA") v = &s->HbyA; else if (k == "phiHbyA") v = &s->phiHbyA; else if (k == "gradP") v = &s->gradP;
    else if (k == "diagU") v = &s->diagU; else if (k == "upperU") v = &s->upperU; else if (k == "lowerU") v = &s->lowerU;
    else if (k == "sourceU") v = &s->sourceU; else if (k == "diagP") v = &s->diagP; else if (k == "upperP") v = &s->upperP;
    else if (k == "sourceP") v = &s->sourceP; else if (k == "icU") v = &s->icU; else if (k == "bcU") v = &s->bcU;
    else if (k == "bGradP") v = &s->m.bGradP;            // [nB] fixedFluxPressure gradient (state between time steps)
    if (!v) { if (n) *n = 0; return nullptr; }
    if (n) *n = (long)v->size();
    return v->data();
}

// phi = linearInterpolate(U) & Sf  (createPhi.H, icoFoamYade/createFields.H:152)
void fvo_create_phi(void* h)
{
    Ico* s = (Ico*)h;
    const Mesh& m = s->m;
    for (int f = 0; f < m.nFaces; ++f) {
        double a = 0;
        for (int j = 0; j < 3; ++j) a += m.Sf[3*(size_t)f + j]*lerp(m.w[f], s->U[3*(size_t)m.l[f] + j], s->U[3*(size_t)m.u[f] + j]);
        s->phi[f] = a;
    }
    for (const Patch& p : m.patches)
        for (int b = p.start; b < p.start + p.n; ++b) {
            double ub[3] = {0, 0, 0}, a = 0;
            if (p.bcU != BC_EMPTY) patchU(m, p, b, s->U.data(), ub);
            for (int j = 0; j < 3; ++j) a += m.bSf[3*(size_t)b + j]*ub[j];
            s->phi[m.nFaces + b] = a;
        }
}

// CourantNo.H + vGrad = fvc::grad(U)   (icoFoamYade.C:68-71): what runs before setParticleAction
void fvo_ico_pre(void* h, double dt)
{
    Ico* s = (Ico*)h;
    courant(s->m, s->phi.data(), dt, &s->st.CoNum, &s->st.meanCoNum);
    gradVector(s->m, s->U.data(), s->vGrad.data());
}

// UEqn assembly, momentum predictor, PISO correctors (icoFoamYade.C:79-140); uSource is read from the state
int fvo_ico_solve(void* h, double dt) { return icoSolve(*(Ico*)h, dt); }

// out: CoNum meanCoNum sumLocal global cumulative nPSolves | U[3]x(init,final,iters) | p[8]x(init,final,iters) | corrSumLocal[8] | corrGlobal[8]
void fvo_get_stats(void* h, double* out)
{
    Ico* s = (Ico*)h;
    int o = 0;
    out[o++] = s->st.CoNum; out[o++] = s->st.meanCoNum; out[o++] = s->st.sumLocalContErr;
    out[o++] = s->st.globalContErr; out[o++] = s->st.cumulativeContErr; out[o++] = s->st.nPSolves;
    for (int j = 0; j < 3; ++j) { out[o++] = s->st.U[j].initialResidual; out[o++] = s->st.U[j].finalResidual; out[o++] = s->st.U[j].nIterations; }
    for (int k = 0; k < 8; ++k) { out[o++] = s->st.p[k].initialResidual; out[o++] = s->st.p[k].finalResidual; out[o++] = s->st.p[k].nIterations; }
    for (int k = 0; k < 8; ++k) out[o++] = s->st.corrSumLocal[k];
    for (int k = 0; k < 8; ++k) out[o++] = s->st.corrGlobal[k];
}
void fvo_get_times(void* h, double* out3)
{
    Ico* s = (Ico*)h;
    out3[0] = s->tMomentum; out3[1] = s->tPressure; out3[2] = s->tOther;
    s->tMomentum = s->tPressure = s->tOther = 0;
}

// stand-alone operators on the state's mesh (stage parity hooks)
void fvo_grad_vector(void* h, const double* U, double* out9) { gradVector(((Ico*)h)->m, U, out9); }
void fvo_grad_scalar(void* h, const double* p, double* out3) { gradScalar(((Ico*)h)->m, p, out3); }
void fvo_div_flux(void* h, const double* phi, double* out) { divFlux(((Ico*)h)->m, phi, out); }

void fvo_div_phi_vector(void* h, const double* phi, const double* U, double* out3) { divPhiU(((Ico*)h)->m, phi, U, out3); }
void fvo_laplacian_gamma_vector(void* h, const double* gamma, double gammaB, const double* U, double* out3)
{
    laplacianGammaU(((Ico*)h)->m, gamma, gammaB, U, out3);
}

// pimpleFoamYade.C:73-76 -- the four fields the solver hands to FoamYade before setParticleAction:
//   ddtU_f = fvc::ddt(Uc) + fvc::div(phic, Uc);  gradP = fvc::grad(p);  divT = 2*nu*fvc::laplacian(alphac, Uc);
//   vGrad = fvc::grad(Uc)
// fvc::ddt(Uc) is evaluated BEFORE Uc is touched in the new time step: GeometricField::oldTime() stores
// Uc.oldTime() := Uc at that first access [OF-6 GeometricField.C storeOldTimes], so the Euler term
// rDeltaT*(Uc - Uc.oldTime()) is rDeltaT*0 = +0 and ddtU_f = 0 + div(phic, Uc).
void fvo_pimple_pre(void* h, double dt, const double* alpha, double* ddtU3, double* gradP3, double* divT3, double* vGrad9)
{
    Ico* s = (Ico*)h;
    const Mesh& m = s->m;
    const double rDeltaT = 1.0/dt;
    courant(m, s->phi.data(), dt, &s->st.CoNum, &s->st.meanCoNum);      // pimpleFoamYade.C:71 #include "CourantNo.H"
    divPhiU(m, s->phi.data(), s->U.data(), ddtU3);
    for (size_t k = 0; k < 3*(size_t)m.nCells; ++k) ddtU3[k] = rDeltaT*(s->U[k] - s->U[k]) + ddtU3[k];
    gradScalar(m, s->p.data(), gradP3);
    laplacianGammaU(m, alpha, 1.0, s->U.data(), divT3);
    const double twoNu = 2*s->nu;
    for (size_t k = 0; k < 3*(size_t)m.nCells; ++k) divT3[k] = twoNu*divT3[k];
    gradVector(m, s->U.data(), vGrad9);
}

// UcEqn.H + pEqn.H + continuityErrs.H (pimpleFoamYade.C:82-104); alpha0 = alphac.oldTime() (== alpha in the reference,
// see the note above pimpleSolve), g = gravitational acceleration; uSource is read from the state
int fvo_pimple_solve(void* h, double dt, const double* alpha, const double* alpha0, const double* uSourceDrag, const double* g3)
{
    Ico* s = (Ico*)h;
    if (!s->pim) s->pim = new Pim();
    return pimpleSolve(*s, *s->pim, dt, alpha, alpha0, uSourceDrag, g3);
}
// pimple intermediates of the last step: phicForces alphaf rAUf recon divDev spDiv
double* fvo_pimple_field(void* h, const char* name, long* n)
{
    Ico* s = (Ico*)h;
    const std::string k(name);
    dvec* v = nullptr;
    if (s->pim) {
        if (k == "phicForces") v = &s->pim->phicForces; else if (k == "alphaf") v = &s->pim->alphaf;
        else if (k == "rAUf") v = &s->pim->rAUf; else if (k == "recon") v = &s->pim->recon;
        else if (k == "divDev") v = &s->pim->divDev; else if (k == "spDiv") v = &s->pim->spDiv;
    }
    if (!v) { if (n) *n = 0; return nullptr; }
    if (n) *n = (long)v->size();
    return v->data();
}
void fvo_reconstruct(void* h, const double* ssf, double* out3)
{
    Ico* s = (Ico*)h;
    if (!s->pim) s->pim = new Pim();
    if (s->pim->invT.size() != 9*(size_t)s->m.nCells) reconstructTensor(s->m, s->pim->invT);
    reconstruct(s->m, s->pim->invT, ssf, out3);
}
void fvo_div_dev(void* h, const double* alpha, double alphaB, double nu, const double* U, double* out3)
{
    divDevTerm(((Ico*)h)->m, alpha, alphaB, nu, U, out3);
}

// PCG on a caller-given symmetric LDU matrix over the state's addressing; out3 = init, final, iters
void fvo_pcg(void* h, const double* diag, const double* upper, const double* source, double* psi, double tol,
             double relTol, int maxIter, int precond, double* out3)
{
    const Mesh& m = ((Ico*)h)->m;
    const dvec d(diag, diag + m.nCells), u(upper, upper + m.nFaces), b(source, source + m.nCells);
    const SolverPerf sp = pcgSolve(m, d, u, b, psi, tol, relTol, maxIter, precond);
    out3[0] = sp.initialResidual; out3[1] = sp.finalResidual; out3[2] = sp.nIterations;
}
// smoothSolver/symGaussSeidel on a caller-given asymmetric LDU matrix
void fvo_smooth(void* h, const double* diag, const double* lower, const double* upper, const double* source,
                double* psi, double tol, double relTol, int maxIter, double* out3)
{
    const Mesh& m = ((Ico*)h)->m;
    const dvec d(diag, diag + m.nCells), lo(lower, lower + m.nFaces), u(upper, upper + m.nFaces), b(source, source + m.nCells);
    const SolverPerf sp = smoothSolve(m, d, lo, u, b, psi, tol, relTol, maxIter);
    out3[0] = sp.initialResidual; out3[1] = sp.finalResidual; out3[2] = sp.nIterations;
}
// one DIC application: wA = M^-1 rA
void fvo_dic(void* h, const double* diag, const double* upper, const double* rA, double* wA)
{
    const Mesh& m = ((Ico*)h)->m;
    const dvec d(diag, diag + m.nCells), u(upper, upper + m.nFaces);
    dvec rD;
    dicReciprocalD(m, d, u, rD);
    dicPrecondition(m, u, rD, rA, wA);
}
void fvo_amul(void* h, const double* diag, const double* lower, const double* upper, const double* psi, double* out)
{
    const Mesh& m = ((Ico*)h)->m;
    const dvec d(diag, diag + m.nCells), lo(lower, lower + m.nFaces), u(upper, upper + m.nFaces);
    Amul(m, d, lo, u, psi, out);
}

}  // extern "C"
