// coupling.cu -- the particle half of Foam::FoamYade::setParticleAction as sm_100a kernels.
//
//   k_locate_gauss   meshTree::nnearestCellsRange / nnearest (meshTree.C:148-238) fused with
//                    calcInterpWeightGaussian (FoamYade.C:293-316), the wire-record unpack of
//                    locateAllParticles (FoamYade.C:187-225) and the per-cell accumulate of
//                    buildCellPartList (FoamYade.C:261-290, here a dense atomic scatter instead of
//                    the reference's quadratic list scan)
//   k_void_fraction  setCellVolFraction (FoamYade.C:318-328)
//   k_force_gauss    hydroDragForce + archimedesForce (FoamYade.C:354-389, 415-435)
//   k_point_force    locatePt/findCell + stokesDragForce + stokesDragTorque (FoamYade.C:248-253, 437-453)
//   k_source_zero    setSourceZero (FoamYade.C:556-566)
//
// All arithmetic is fp64 and written in the reference's own operation order; the library is built
// with -fmad=false and the distance test of the tree descent additionally uses explicit
// round-to-nearest intrinsics, because the strict `<` tests on d^2 decide which cells are returned
// and the x86-64 reference build has no FMA contraction.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <mutex>
#include <set>
#include <utility>
#include <type_traits>
#include <cub/device/device_radix_sort.cuh>

#include "fy_ctx.h"

namespace {

__device__ __forceinline__ FyKdNode ldNode(const FyKdNode* __restrict__ t, int i)
{
    // one 256-bit load (sm_100+): x y z | id pad
    FyKdNode n;
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(t + i));
    n.x = __longlong_as_double((long long)a);
    n.y = __longlong_as_double((long long)b);
    n.z = __longlong_as_double((long long)c);
    n.id = (int)(unsigned)(d & 0xffffffffull);
    n.pad = 0;
    return n;
}

__device__ __forceinline__ double dist2(double px, double py, double pz, const FyKdNode& n)
{
    // meshTree.C:54-64: dist = 0; for i: ds = p1[i]-p2[i]; dist += ds*ds   (no contraction)
    const double dx = __dsub_rn(px, n.x), dy = __dsub_rn(py, n.y), dz = __dsub_rn(pz, n.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// The nearest-neighbour descent of meshTree.C:182-238 on the implicit tree, iteratively.
// Every node visit that STRICTLY improves the best squared distance found so far (initially that of
// the root, so the root itself never qualifies, MT.C:156,192) and lies within maxDist is an
// "improvement"; the reference's bounded container ends up holding the nearest (= latest) <= 12 of
// them in ascending distance.  ring/ringD keep the last 12; nImp counts all of them.
struct Trail {                                  // per-thread local memory (dynamically indexed)
    int id[FY_MAXLIST];
    double d2[FY_MAXLIST];
    int nImp;
    __device__ __forceinline__ void put(int s, int c, double d) { id[s] = c; d2[s] = d; }
    __device__ __forceinline__ int getId(int s) const { return id[s]; }
    __device__ __forceinline__ double getD(int s) const { return d2[s]; }
    __device__ __forceinline__ void setD(int s, double d) { d2[s] = d; }
};
// the same ring in SHARED memory, slot-major ([slot][thread]: conflict-free): 144 bytes per thread that no longer pass
// through the per-thread local-memory window (which ncu shows being written back to DRAM once per thread).  Measured
// slower than the local-memory ring (see gaussLocate); not the default.
struct TrailS {
    int* id;
    double* d2;
    int nImp;
    __device__ __forceinline__ void put(int s, int c, double d) { id[s * blockDim.x] = c; d2[s * blockDim.x] = d; }
    __device__ __forceinline__ int getId(int s) const { return id[s * blockDim.x]; }
    __device__ __forceinline__ double getD(int s) const { return d2[s * blockDim.x]; }
    __device__ __forceinline__ void setD(int s, double d) { d2[s * blockDim.x] = d; }
};

// The stack of pending "other" subtrees: 16 bytes per entry (the range [lo, hi) with the depth in the top 5 bits of hi,
// and the squared plane distance), per-thread local memory.  A subtree whose plane distance is not below the best
// distance AT PUSH TIME is never pushed: the reference tests it after the near side returned (MT.C:225), when `best`
// can only be smaller, so it would be rejected then anyway.  (Measured on B200: keeping the stack in shared memory
// instead -- 61 KB per 128-thread block, 12 warps per SM -- costs more in lost latency hiding than the local-memory
// traffic it removes: 2.39 ms against 1.39 ms at C2.)
constexpr int KD_STACK = 30;                  // >= tree depth (2^26 nodes: 27 levels)
constexpr int KD_BLOCK = 128;
constexpr int KD_SMEM = 0;

template <class TR>
__device__ __forceinline__ void kdDescend(const FyKdNode* __restrict__ tree, int nTree, double px, double py,
                                          double pz, double maxDist, TR& tr)
{
    tr.nImp = 0;
    if (nTree <= 0) return;
    constexpr int T = 1, me = 0;
    int2 sRange[KD_STACK];
    double sDf2[KD_STACK];
    int sp = 0;

    int lo = 0, hi = nTree, depth = 0;
    double best = dist2(px, py, pz, ldNode(tree, nTree >> 1));     // MT.C:156
    for (;;) {
        if (lo < hi) {
            const int md = lo + ((hi - lo) >> 1);
            const FyKdNode nd = ldNode(tree, md);
            const double d = dist2(px, py, pz, nd);
            if (d < best) {                                         // MT.C:192 (and the re-tests at 217, 229)
                best = d;
                if (d < maxDist) {                                  // MT.C:195
                    tr.put(tr.nImp % FY_MAXLIST, nd.id, d);
                    tr.nImp++;
                }
            }
            const int axis = depth % 3;
            const double nc = axis == 0 ? nd.x : (axis == 1 ? nd.y : nd.z);
            const double pc = axis == 0 ? px : (axis == 1 ? py : pz);
            const double df = __dsub_rn(nc, pc);                    // MT.C:200
            const double df2 = __dmul_rn(df, df);
            int olo, ohi;
            if (df > 0.0) { olo = md + 1; ohi = hi; hi = md; }      // next = left, other = right (MT.C:206-212)
            else          { olo = lo; ohi = md; lo = md + 1; }
            depth++;
            if (olo < ohi && df2 < best && sp < KD_STACK) {
                sRange[sp * T + me] = make_int2(olo, ohi | (depth << 26));
                sDf2[sp * T + me] = df2;
                sp++;
            }
        } else {
            bool got = false;
            while (sp > 0) {
                --sp;
                if (sDf2[sp * T + me] < best) {                     // MT.C:225, tested after the near side returned
                    const int2 r = sRange[sp * T + me];
                    lo = r.x; hi = r.y & ((1 << 26) - 1); depth = (int)((unsigned)r.y >> 26);
                    got = true;
                    break;
                }
            }
            if (!got) break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// parity hook: cell lists only
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_locate(const FyKdNode* __restrict__ tree, int nTree,
                                                const double* __restrict__ xyz, int stride, int n, double maxDist,
                                                int* __restrict__ ids, int* __restrict__ cnt)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double px = xyz[(size_t)p * stride], py = xyz[(size_t)p * stride + 1], pz = xyz[(size_t)p * stride + 2];
    Trail tr;
    kdDescend(tree, nTree, px, py, pz, maxDist, tr);
    const int k = tr.nImp < FY_MAXLIST ? tr.nImp : FY_MAXLIST;
    for (int j = 0; j < FY_MAXLIST; ++j)
        ids[(size_t)p * FY_MAXLIST + j] = j < k ? tr.getId((tr.nImp - 1 - j) % FY_MAXLIST) : -1;
    cnt[p] = k;
}

__device__ __forceinline__ int boxCell(double x, double y, double z, int nx, int ny, int nz, double x0, double y0,
                                       double z0, double hx, double hy, double hz)
{
    const double fi = floor((x - x0) / hx), fj = floor((y - y0) / hy), fk = floor((z - z0) / hz);
    if (!(fi >= 0 && fi < nx && fj >= 0 && fj < ny && fk >= 0 && fk < nz)) return -1;
    return (int)fi + nx * ((int)fj + ny * (int)fk);
}

__global__ void k_find_cell(const double* __restrict__ xyz, int stride, int n, int nx, int ny, int nz, double x0,
                            double y0, double z0, double hx, double hy, double hz, int* __restrict__ cell)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    cell[p] = boxCell(xyz[(size_t)p * stride], xyz[(size_t)p * stride + 1], xyz[(size_t)p * stride + 2], nx, ny, nz,
                      x0, y0, z0, hx, hy, hz);
}

// ---------------------------------------------------------------------------------------------
// Gaussian branch, pass 1: locate + weights + per-cell accumulate
// ---------------------------------------------------------------------------------------------
struct GaussConst {
    double maxDist;        // range^2 + 0.25 range^2                      (MT.C:155)
    double twoSigmaSq;     // 2*pow(sigmaInterp, 2)                       (F.C:308)
    double interpRangeCu;  // pow(interpRange, 3)                         (F.C:71)
    double sigmaPi;        // 1/pow(2 pi sigma^2, 1.5)                    (F.C:72)
};

template <bool SMEM_TRAIL>
__global__ void __launch_bounds__(128)
k_locate_gauss(const FyKdNode* __restrict__ tree, int nTree, const double* __restrict__ pdata, int n,
               const int* __restrict__ perm, GaussConst gc, int serial, int* __restrict__ ids, int* __restrict__ cnt,
               double* __restrict__ wts, int* __restrict__ found, double* __restrict__ pvolAcc,
               double* __restrict__ upAcc, int* __restrict__ stamp)
{
    __shared__ int sTrailId[SMEM_TRAIL ? FY_MAXLIST * KD_BLOCK : 1];
    __shared__ double sTrailD[SMEM_TRAIL ? FY_MAXLIST * KD_BLOCK : 1];
    // thread t works on particle perm[t] (particles sorted by position so that a warp shares tree paths and cells);
    // cell lists / weights are kept in sorted order, structure-of-arrays: ids[j*n + t]
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = perm[t];
    const double* rec = pdata + (size_t)p * 10;
    const double px = rec[0], py = rec[1], pz = rec[2];
    typename std::conditional<SMEM_TRAIL, TrailS, Trail>::type tr;
    if constexpr (SMEM_TRAIL) { tr.id = sTrailId + threadIdx.x; tr.d2 = sTrailD + threadIdx.x; }
    kdDescend(tree, nTree, px, py, pz, gc.maxDist, tr);
    const int k = tr.nImp < FY_MAXLIST ? tr.nImp : FY_MAXLIST;
    cnt[t] = k;
    found[p] = k > 0 ? 1 : -1;                                       // F.C:204,222 / 141
    if (k == 0) {
        for (int j = 0; j < FY_MAXLIST; ++j) ids[(size_t)j * n + t] = -1;
        return;
    }
    // weights (F.C:301-314): the squared distance is the same number the descent computed
    // ((C-p)^2 == (p-C)^2 term by term, same summation order), so no cell-centre gather is needed.
    double allwt = 0.0;
    for (int j = 0; j < k; ++j) {
        const int s = (tr.nImp - 1 - j) % FY_MAXLIST;
        const double wj = exp(-tr.getD(s) / gc.twoSigmaSq) * gc.interpRangeCu * gc.sigmaPi;
        tr.setD(s, wj);                                              // (the weight takes the distance's place: no second array)
        allwt += wj;
    }
    const double vx = rec[3], vy = rec[4], vz = rec[5];
    const double dia = 2 * rec[9];                                   // F.C:219
    const double vol = M_PI * pow(dia, 3.0) / 6.0;                   // F.H:36
    for (int j = 0; j < FY_MAXLIST; ++j) {
        if (j < k) {
            const int s = (tr.nImp - 1 - j) % FY_MAXLIST;
            const double wj = tr.getD(s) / allwt;                    // F.C:313
            const int c = tr.getId(s);
            ids[(size_t)j * n + t] = c;
            wts[(size_t)j * n + t] = wj;
            // F.C:271-272 / 278-279: pVol*weight ; (linearVelocity*weight)*pVol
            atomicAdd(&pvolAcc[c], vol * wj);
            atomicAdd(&upAcc[3 * (size_t)c], vx * wj * vol);
            atomicAdd(&upAcc[3 * (size_t)c + 1], vy * wj * vol);
            atomicAdd(&upAcc[3 * (size_t)c + 2], vz * wj * vol);
            stamp[c] = serial;
        } else {
            ids[(size_t)j * n + t] = -1;
            wts[(size_t)j * n + t] = 0.0;
        }
    }
}

// pass 2: setCellVolFraction on the cells this proc touched (F.C:318-328); consumes and clears the
// accumulators so that the next proc starts from zero.
__global__ void k_void_fraction(int nCells, int serial, const int* __restrict__ stamp, double* __restrict__ pvolAcc,
                                double* __restrict__ upAcc, const double* __restrict__ V, double* __restrict__ alpha,
                                double* __restrict__ uParticle)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    if (stamp[c] != serial) return;
    const double v = V[c];
    const double pvolC = 1.0 - (pvolAcc[c] / v);
    alpha[c] = (pvolC > 0.10) ? pvolC : 0.10;
    uParticle[3 * (size_t)c] = upAcc[3 * (size_t)c] / v;
    uParticle[3 * (size_t)c + 1] = upAcc[3 * (size_t)c + 1] / v;
    uParticle[3 * (size_t)c + 2] = upAcc[3 * (size_t)c + 2] / v;
    pvolAcc[c] = 0.0;
    upAcc[3 * (size_t)c] = 0.0;
    upAcc[3 * (size_t)c + 1] = 0.0;
    upAcc[3 * (size_t)c + 2] = 0.0;
}

// pass 3: hydroDragForce + archimedesForce per particle, reaction scattered to the cells.
struct ForceConst {
    double rhoF, nu, small;
    // the forces the reference defines but never calls (EXTRA instantiations only)
    double rhoP, deltaT;
    int addedMass, torque;          // addedMassForce F.C:392-413; calcHydroTorque's Gaussian branch F.C:467-478
};

// addedMassForce (F.C:392-413) from the gathered sums: pv = (sum vol*w)/listSize (the reference divides by the list length),
// f = pv*(ddtUf - linearVelocity/deltaT)*rhoP
__device__ __forceinline__ void addedMassOf(const ForceConst& fc, double pvSum, int listSize, const double* ddtUf, double vx,
                                            double vy, double vz, double* f)
{
    const double pv = pvSum / listSize;
    f[0] = pv * (ddtUf[0] - (vx / fc.deltaT)) * fc.rhoP;
    f[1] = pv * (ddtUf[1] - (vy / fc.deltaT)) * fc.rhoP;
    f[2] = pv * (ddtUf[2] - (vz / fc.deltaT)) * fc.rhoP;
}
// hydroTorque += M_PI*(pow(dia,3))*(wfluid - rotationalVelocity)*nu*rhoF   (F.C:478)
__device__ __forceinline__ void gaussTorqueOf(const ForceConst& fc, double dia, const double* wf, const double* rec, double* T)
{
    const double c = M_PI * (pow(dia, 3.0));
    T[0] = c * (wf[0] - rec[6]) * fc.nu * fc.rhoF;
    T[1] = c * (wf[1] - rec[7]) * fc.nu * fc.rhoF;
    T[2] = c * (wf[2] - rec[8]) * fc.nu * fc.rhoF;
}

template <bool EXTRA>
__global__ void __launch_bounds__(128)
k_force_gauss(const double* __restrict__ pdata, int n, const int* __restrict__ perm, const int* __restrict__ ids,
              const int* __restrict__ cnt, const double* __restrict__ wts, ForceConst fc, const double* __restrict__ U,
              const double* __restrict__ alpha, const double* __restrict__ uParticle,
              const double* __restrict__ gradP, const double* __restrict__ divT, const double* __restrict__ V,
              const double* __restrict__ ddtU, const double* __restrict__ vGrad,
              double* __restrict__ uSourceDrag, double* __restrict__ uSource, double* __restrict__ force)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = perm[t];
    double* F = force + (size_t)p * 6;
    const int k = cnt[t];
    if (k <= 0) {
        F[0] = F[1] = F[2] = F[3] = F[4] = F[5] = 0.0;                // F.C:142 zero-initialised buffer
        return;
    }
    const double* rec = pdata + (size_t)p * 10;
    const double vx = rec[3], vy = rec[4], vz = rec[5];
    const double dia = 2 * rec[9];
    const double vol = M_PI * pow(dia, 3.0) / 6.0;
    const double rhoF = fc.rhoF, nu = fc.nu;

    int id[FY_MAXLIST];
    double w[FY_MAXLIST];
    // ---- hydroDragForce gather (F.C:358-365) and archimedesForce gather (F.C:416-424)
    double ufx = 0, ufy = 0, ufz = 0, alpha_f = 0, pv = 0;
    double dtx = 0, dty = 0, dtz = 0, pgx = 0, pgy = 0, pgz = 0;
    double ddtUf[3] = {0, 0, 0}, wf[3] = {0, 0, 0};
    const double twoNu = 2.0 * nu;
    for (int j = 0; j < k; ++j) {
        const int c = ids[(size_t)j * n + t];
        const double wj = wts[(size_t)j * n + t];
        id[j] = c;
        w[j] = wj;
        if (EXTRA) {
            if (fc.addedMass)
                for (int m = 0; m < 3; ++m) ddtUf[m] = ddtUf[m] + (ddtU[3 * (size_t)c + m] * wj);              // F.C:400
            if (fc.torque) {
                const double* g = vGrad + 9 * (size_t)c;                                                       // F.C:471-473
                wf[0] += ((g[5] - g[7]) * wj);
                wf[1] += ((g[6] - g[2]) * wj);
                wf[2] += ((g[3] - g[1]) * wj);
            }
        }
        ufx += U[3 * (size_t)c] * wj;
        ufy += U[3 * (size_t)c + 1] * wj;
        ufz += U[3 * (size_t)c + 2] * wj;
        alpha_f += alpha[c] * wj;
        pv += vol * wj;
        dtx = dtx + (twoNu * divT[3 * (size_t)c] * wj * rhoF);
        dty = dty + (twoNu * divT[3 * (size_t)c + 1] * wj * rhoF);
        dtz = dtz + (twoNu * divT[3 * (size_t)c + 2] * wj * rhoF);
        pgx = pgx + (gradP[3 * (size_t)c] * wj);
        pgy = pgy + (gradP[3 * (size_t)c + 1] * wj);
        pgz = pgz + (gradP[3 * (size_t)c + 2] * wj);
    }
    const double alpha_p = 1 - alpha_f;
    const double urx = ufx - vx, ury = ufy - vy, urz = ufz - vz;
    const double magUR = sqrt(urx * urx + ury * ury + urz * urz);
    const double Re = fc.small + ((magUR * dia) / nu);                                     // F.C:370
    const double cd = Re < 1000 ? (24 / (Re)) * (1 + (0.15 * pow(Re, 0.687))) : 0.44;      // F.C:371
    double coeff;
    if (alpha_f > 0.8) {
        coeff = 0.75 * cd * alpha_f * alpha_p * rhoF * magUR * pow(alpha_f, -2.65);        // F.C:374
    } else {
        const double cf1 = 150 * ((alpha_p * alpha_p) / alpha_f) * ((nu * rhoF) / (dia * dia));
        const double cf2 = 1.75 * alpha_p * rhoF * (1 / dia) * magUR;
        coeff = cf1 + cf2;
    }
    const double pc = pv * coeff, ooap = 1 / (alpha_p);
    double Fx = pc * urx * ooap, Fy = pc * ury * ooap, Fz = pc * urz * ooap;               // F.C:381
    // archimedes (F.C:426): f = pv*(-pg + divt)
    const double ax = pv * (-pgx + dtx), ay = pv * (-pgy + dty), az = pv * (-pgz + dtz);
    Fx += ax; Fy += ay; Fz += az;
    double am[3] = {0, 0, 0}, T[3] = {0.0, 0.0, 0.0};
    if (EXTRA) {
        if (fc.addedMass) {
            addedMassOf(fc, pv, k, ddtUf, vx, vy, vz, am);
            Fx += am[0]; Fy += am[1]; Fz += am[2];
        }
        if (fc.torque) gaussTorqueOf(fc, dia, wf, rec, T);
    }
    F[0] = Fx; F[1] = Fy; F[2] = Fz;
    F[3] = T[0]; F[4] = T[1]; F[5] = T[2];                           // zero unless enabled: torque is disabled on this branch (F.C:618)

    // ---- scatter (F.C:384-387 and 429-434); the two uSource contributions of a pair are summed
    //      before the atomic so that each cell component sees one RED per pair
    const double oorho = 1 / rhoF;
    for (int j = 0; j < k; ++j) {
        const int c = id[j];
        const double wj = w[j];
        const double mcw = -coeff * wj;
        atomicAdd(&uSourceDrag[c], mcw * oorho);
        const double ooCellVol = 1. / (V[c] * rhoF);
        double sx = (mcw * uParticle[3 * (size_t)c]) / rhoF + (-ax * wj * ooCellVol);
        double sy = (mcw * uParticle[3 * (size_t)c + 1]) / rhoF + (-ay * wj * ooCellVol);
        double sz = (mcw * uParticle[3 * (size_t)c + 2]) / rhoF + (-az * wj * ooCellVol);
        if (EXTRA && fc.addedMass) {                                 // F.C:410
            sx += (-am[0] * wj * ooCellVol);
            sy += (-am[1] * wj * ooCellVol);
            sz += (-am[2] * wj * ooCellVol);
        }
        atomicAdd(&uSource[3 * (size_t)c], sx);
        atomicAdd(&uSource[3 * (size_t)c + 1], sy);
        atomicAdd(&uSource[3 * (size_t)c + 2], sz);
    }
}


// ---------------------------------------------------------------------------------------------
// Full-support Gaussian mode ("range based search", README.md:5): EVERY cell whose centre lies within the
// search bound of meshTree::nnearestCellsRange (d^2 < range^2 + 0.25 range^2, MT.C:155) carries a weight -- ~370 cells
// per particle on a uniform mesh (range = 4 h) instead of the <= 12 cells of the k-d descent's improvement trail.
// One WARP per particle: the lanes stride over the candidate cells of the particle's index box (the hex box's own
// tensor-product centre coordinates, verified against mesh.C() at fy_create), the normaliser and the eleven gathered
// sums of hydroDragForce / archimedesForce are warp-shuffle reductions, the reaction is scattered with one RED per
// cell and component.  The weights are never stored (a 370-entry list per particle would be 4.4 KB): both passes
// recompute exp(-d^2/2 sigma^2) from the same d^2, bit for bit.  Oracle: the unmodified reference's own
// calcInterpWeightGaussian / hydroDragForce / archimedesForce fed with the full cell lists (oracle/ref_harness.cpp).
// ---------------------------------------------------------------------------------------------
struct RangeBox {
    int nx, ny, nz;
    double x0, y0, z0, hx, hy, hz;
    const double* ax;      // centre coordinates: xs[nx] | ys[ny] | zs[nz]
    double R;              // sqrt(maxDist)
    int cap;               // hits kept in shared memory per particle (<= RANGE_CAP; environment FY_RANGE_CAP, for tests)
};

__device__ __forceinline__ void rangeSpan(double p, double R, double x0, double h, int n, int& lo, int& cnt)
{
    // indices whose centre x0 + (i + 1/2) h can lie within R of p, one index of slack each side (the exact test decides)
    double a = floor((p - R - x0) / h - 0.5), b = floor((p + R - x0) / h - 0.5) + 1.0;
    a = a < 0.0 ? 0.0 : a;
    b = b > (double)(n - 1) ? (double)(n - 1) : b;
    lo = a > (double)n ? n : (int)a;
    const int hi = b < -1.0 ? -1 : (int)b;
    cnt = hi >= lo ? hi - lo + 1 : 0;
}

__device__ __forceinline__ double warpSum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct RangeIter {
    int lo[3], ni, nj, nk;
    double px, py, pz;
    __device__ __forceinline__ void init(const RangeBox& b, double x, double y, double z)
    {
        px = x; py = y; pz = z;
        rangeSpan(x, b.R, b.x0, b.hx, b.nx, lo[0], ni);
        rangeSpan(y, b.R, b.y0, b.hy, b.ny, lo[1], nj);
        rangeSpan(z, b.R, b.z0, b.hz, b.nz, lo[2], nk);
        if (ni == 0 || nj == 0 || nk == 0) ni = nj = nk = 0;
    }
};

// One pass over the candidate cells of a particle, by the whole warp: lane l takes the candidate ROWS (j, k) l, l + 32, ...
// (one integer division per row, not per candidate; the y and z terms of the distance once per row) and all lanes walk
// i together.  f(hit, cell, d2) is called CONVERGENTLY for every step, so it may contain warp collectives; d2 is
// ((dx^2 + dy^2) + dz^2) in meshTree::distance's order (MT.C:54-64).  A row whose y/z terms alone reach the bound is
// skipped: rounding is monotone, so fl(fl(dx^2 + dy^2) + dz^2) >= fl(dy^2 + dz^2) for every dx.
template <class F>
__device__ __forceinline__ void rangeScan(const RangeBox& b, const RangeIter& it, double maxDist, F&& f)
{
    const int lane = threadIdx.x & 31;
    const int nrows = it.nj * it.nk;
    for (int base = 0; base < nrows; base += 32) {
        const int r = base + lane;
        const bool valid = r < nrows;
        const int dk = valid ? r / it.nj : 0, dj = valid ? r - dk * it.nj : 0;
        const int j = it.lo[1] + dj, k = it.lo[2] + dk;
        const double dy = __dsub_rn(b.ax[b.nx + j], it.py), dz = __dsub_rn(b.ax[b.nx + b.ny + k], it.pz);
        const double dysq = __dmul_rn(dy, dy), dzsq = __dmul_rn(dz, dz);
        const bool rowOk = valid && __dadd_rn(dysq, dzsq) < maxDist;
        if (!__any_sync(0xffffffffu, rowOk)) continue;
        const int crow = b.nx * (j + b.ny * k);
        for (int di = 0; di < it.ni; ++di) {
            const int i = it.lo[0] + di;
            const double dx = __dsub_rn(b.ax[i], it.px);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), dysq), dzsq);
            f(rowOk && d2 < maxDist, crow + i, d2);
        }
    }
}

constexpr int RANGE_WARPS = 8;
// The hits of a particle are compacted into shared memory (cell id + unnormalised weight: 12 bytes each) by the scan, so
// that exp() is evaluated once per hit and pass, and the gather / scatter loops run over a dense list.  A uniform mesh
// gives <= ~410 hits (range = 4 h); a particle with more (strongly anisotropic cells) takes the recomputing path.
constexpr int RANGE_CAP = 448;

struct RangeList {
    int* id;
    double* w;
    int count;
};

// scan + compaction: fills L (up to RANGE_CAP entries), returns the lane-partial sum of the weights of ALL hits
__device__ __forceinline__ double rangeCollect(const RangeBox& b, const RangeIter& it, const GaussConst& gc, RangeList& L)
{
    const int lane = threadIdx.x & 31;
    double sw = 0.0;
    int count = 0;
    rangeScan(b, it, gc.maxDist, [&](bool hit, int c, double d2) {
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const double w = exp(-d2 / gc.twoSigmaSq) * gc.interpRangeCu * gc.sigmaPi;      // F.C:308
            sw += w;
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (pos < b.cap) { L.id[pos] = c; L.w[pos] = w; }
        }
        count += __popc(m);
    });
    L.count = count;
    __syncwarp();
    return sw;
}

// pass 1: count + normaliser + per-cell accumulate (locateAllParticles + calcInterpWeightGaussian + buildCellPartList)
__global__ void __launch_bounds__(RANGE_WARPS * 32)
k_range_accumulate(RangeBox b, const double* __restrict__ pdata, int n, const int* __restrict__ perm, GaussConst gc,
                   int serial, int* __restrict__ cnt, double* __restrict__ allwtOut, int* __restrict__ found,
                   double* __restrict__ pvolAcc, double* __restrict__ upAcc, int* __restrict__ stamp)
{
    __shared__ int sId[RANGE_WARPS][RANGE_CAP];
    __shared__ double sW[RANGE_WARPS][RANGE_CAP];
    const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * RANGE_WARPS + wq;
    if (t >= n) return;
    const int p = perm[t];
    const double* rec = pdata + (size_t)p * 10;
    RangeIter it;
    it.init(b, rec[0], rec[1], rec[2]);
    RangeList L{sId[wq], sW[wq], 0};
    const double sw = warpSum(rangeCollect(b, it, gc, L));
    const int hits = L.count;
    if (lane == 0) {
        cnt[t] = hits;
        allwtOut[t] = sw;
        found[p] = hits > 0 ? 1 : -1;
    }
    if (hits == 0) return;
    const double vx = rec[3], vy = rec[4], vz = rec[5];
    const double dia = 2 * rec[9];
    const double vol = M_PI * pow(dia, 3.0) / 6.0;
    auto deposit = [&](int c, double w) {
        const double wj = w / sw;                                                           // F.C:313
        atomicAdd(&pvolAcc[c], vol * wj);                                                   // F.C:271-272 / 278-279
        atomicAdd(&upAcc[3 * (size_t)c], vx * wj * vol);
        atomicAdd(&upAcc[3 * (size_t)c + 1], vy * wj * vol);
        atomicAdd(&upAcc[3 * (size_t)c + 2], vz * wj * vol);
        stamp[c] = serial;
    };
    if (hits <= b.cap) {
        for (int q = lane; q < hits; q += 32) deposit(L.id[q], L.w[q]);
    } else {
        rangeScan(b, it, gc.maxDist, [&](bool hit, int c, double d2) {
            if (hit) deposit(c, exp(-d2 / gc.twoSigmaSq) * gc.interpRangeCu * gc.sigmaPi);
        });
    }
}

// pass 3: hydroDragForce + archimedesForce (+ addedMassForce, Gaussian torque) over the full support
template <bool EXTRA>
__global__ void __launch_bounds__(RANGE_WARPS * 32)
k_range_force(RangeBox b, const double* __restrict__ pdata, int n, const int* __restrict__ perm, const int* __restrict__ cnt,
              const double* __restrict__ allwt, GaussConst gc, ForceConst fc, const double* __restrict__ U,
              const double* __restrict__ alpha, const double* __restrict__ uParticle, const double* __restrict__ gradP,
              const double* __restrict__ divT, const double* __restrict__ V, const double* __restrict__ ddtU,
              const double* __restrict__ vGrad, double* __restrict__ uSourceDrag, double* __restrict__ uSource,
              double* __restrict__ force)
{
    __shared__ int sId[RANGE_WARPS][RANGE_CAP];
    __shared__ double sW[RANGE_WARPS][RANGE_CAP];
    const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * RANGE_WARPS + wq;
    if (t >= n) return;
    const int p = perm[t];
    double* F = force + (size_t)p * 6;
    const int k = cnt[t];
    if (k <= 0) {
        if (lane < 6) F[lane] = 0.0;
        return;
    }
    const double* rec = pdata + (size_t)p * 10;
    const double vx = rec[3], vy = rec[4], vz = rec[5];
    const double dia = 2 * rec[9];
    const double vol = M_PI * pow(dia, 3.0) / 6.0;
    const double rhoF = fc.rhoF, nu = fc.nu, sw = allwt[t];
    const double twoNu = 2.0 * nu;
    RangeIter it;
    it.init(b, rec[0], rec[1], rec[2]);
    RangeList L{sId[wq], sW[wq], 0};
    const bool dense = k <= b.cap;
    if (dense) rangeCollect(b, it, gc, L);                 // the same hits and the same weights as pass 1, bit for bit
    // gathers: uf[3] alpha_f pv divt[3] pg[3] | ddtUf[3] wfluid[3]
    double g[EXTRA ? 17 : 11];
#pragma unroll
    for (int m = 0; m < (EXTRA ? 17 : 11); ++m) g[m] = 0.0;
    auto gather = [&](int c, double w) {
        const double wj = w / sw;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            g[m] += U[3 * (size_t)c + m] * wj;                                            // F.C:360
            g[5 + m] = g[5 + m] + (twoNu * divT[3 * (size_t)c + m] * wj * rhoF);          // F.C:422
            g[8 + m] = g[8 + m] + (gradP[3 * (size_t)c + m] * wj);                        // F.C:423
        }
        g[3] += alpha[c] * wj;
        g[4] += vol * wj;
        if (EXTRA) {
            if (fc.addedMass)
                for (int m = 0; m < 3; ++m) g[11 + m] = g[11 + m] + (ddtU[3 * (size_t)c + m] * wj);
            if (fc.torque) {
                const double* vg = vGrad + 9 * (size_t)c;
                g[14] += ((vg[5] - vg[7]) * wj);
                g[15] += ((vg[6] - vg[2]) * wj);
                g[16] += ((vg[3] - vg[1]) * wj);
            }
        }
    };
    if (dense) {
        for (int q = lane; q < k; q += 32) gather(L.id[q], L.w[q]);
    } else {
        rangeScan(b, it, gc.maxDist, [&](bool hit, int c, double d2) {
            if (hit) gather(c, exp(-d2 / gc.twoSigmaSq) * gc.interpRangeCu * gc.sigmaPi);
        });
    }
#pragma unroll
    for (int m = 0; m < (EXTRA ? 17 : 11); ++m) g[m] = warpSum(g[m]);
    const double alpha_f = g[3], pv = g[4];
    const double alpha_p = 1 - alpha_f;
    const double urx = g[0] - vx, ury = g[1] - vy, urz = g[2] - vz;
    const double magUR = sqrt(urx * urx + ury * ury + urz * urz);
    const double Re = fc.small + ((magUR * dia) / nu);                                     // F.C:370
    const double cd = Re < 1000 ? (24 / (Re)) * (1 + (0.15 * pow(Re, 0.687))) : 0.44;      // F.C:371
    double coeff;
    if (alpha_f > 0.8) {
        coeff = 0.75 * cd * alpha_f * alpha_p * rhoF * magUR * pow(alpha_f, -2.65);        // F.C:374
    } else {
        const double cf1 = 150 * ((alpha_p * alpha_p) / alpha_f) * ((nu * rhoF) / (dia * dia));
        const double cf2 = 1.75 * alpha_p * rhoF * (1 / dia) * magUR;
        coeff = cf1 + cf2;
    }
    const double pc = pv * coeff, ooap = 1 / (alpha_p);
    double Fx = pc * urx * ooap, Fy = pc * ury * ooap, Fz = pc * urz * ooap;               // F.C:381
    const double ax = pv * (-g[8] + g[5]), ay = pv * (-g[9] + g[6]), az = pv * (-g[10] + g[7]);   // F.C:426
    Fx += ax; Fy += ay; Fz += az;
    double am[3] = {0, 0, 0}, T[3] = {0.0, 0.0, 0.0};
    if (EXTRA) {
        if (fc.addedMass) {
            addedMassOf(fc, pv, k, &g[11], vx, vy, vz, am);
            Fx += am[0]; Fy += am[1]; Fz += am[2];
        }
        if (fc.torque) gaussTorqueOf(fc, dia, &g[14], rec, T);
    }
    if (lane == 0) {
        F[0] = Fx; F[1] = Fy; F[2] = Fz;
        F[3] = T[0]; F[4] = T[1]; F[5] = T[2];
    }
    const double oorho = 1 / rhoF;
    auto scatter = [&](int c, double w) {
        const double wj = w / sw;
        const double mcw = -coeff * wj;
        atomicAdd(&uSourceDrag[c], mcw * oorho);                                           // F.C:385
        const double ooCellVol = 1. / (V[c] * rhoF);
        double sx = (mcw * uParticle[3 * (size_t)c]) / rhoF + (-ax * wj * ooCellVol);      // F.C:386, 433
        double sy = (mcw * uParticle[3 * (size_t)c + 1]) / rhoF + (-ay * wj * ooCellVol);
        double sz = (mcw * uParticle[3 * (size_t)c + 2]) / rhoF + (-az * wj * ooCellVol);
        if (EXTRA && fc.addedMass) {
            sx += (-am[0] * wj * ooCellVol);
            sy += (-am[1] * wj * ooCellVol);
            sz += (-am[2] * wj * ooCellVol);
        }
        atomicAdd(&uSource[3 * (size_t)c], sx);
        atomicAdd(&uSource[3 * (size_t)c + 1], sy);
        atomicAdd(&uSource[3 * (size_t)c + 2], sz);
    };
    if (dense) {
        for (int q = lane; q < k; q += 32) scatter(L.id[q], L.w[q]);
    } else {
        rangeScan(b, it, gc.maxDist, [&](bool hit, int c, double d2) {
            if (hit) scatter(c, exp(-d2 / gc.twoSigmaSq) * gc.interpRangeCu * gc.sigmaPi);
        });
    }
}

__global__ void k_unpermute_counts(int n, const int* __restrict__ perm, const int* __restrict__ cnt, int* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[perm[t]] = cnt[t];
}

// ---------------------------------------------------------------------------------------------
// point-force branch: findCell + Stokes drag + Stokes torque, one cell per particle
// ---------------------------------------------------------------------------------------------
struct BoxConst {
    int nx, ny, nz;
    double x0, y0, z0, hx, hy, hz;
};

__global__ void __launch_bounds__(256)
k_point_force(const double* __restrict__ pdata, int n, BoxConst b, ForceConst fc, const double* __restrict__ U,
              const double* __restrict__ vGrad, const double* __restrict__ V, double* __restrict__ uSource,
              int* __restrict__ cellOut, int* __restrict__ found, double* __restrict__ force)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double* rec = pdata + (size_t)p * 10;
    double* F = force + (size_t)p * 6;
    const int c = boxCell(rec[0], rec[1], rec[2], b.nx, b.ny, b.nz, b.x0, b.y0, b.z0, b.hx, b.hy, b.hz);
    cellOut[p] = c;
    if (c < 0) {
        found[p] = -1;
        F[0] = F[1] = F[2] = F[3] = F[4] = F[5] = 0.0;
        return;
    }
    found[p] = 1;
    const double dia = 2 * rec[9];
    const double rhoF = fc.rhoF, nu = fc.nu;
    // stokesDragForce (F.C:437-444)
    const double coeff = 3 * M_PI * (dia) * nu * rhoF;
    const double ooCellVol = 1. / (V[c] * rhoF);
    const double Fx = coeff * (U[3 * (size_t)c] - rec[3]);
    const double Fy = coeff * (U[3 * (size_t)c + 1] - rec[4]);
    const double Fz = coeff * (U[3 * (size_t)c + 2] - rec[5]);
    const double m = -1 * ooCellVol;
    atomicAdd(&uSource[3 * (size_t)c], m * Fx);
    atomicAdd(&uSource[3 * (size_t)c + 1], m * Fy);
    atomicAdd(&uSource[3 * (size_t)c + 2], m * Fz);
    // stokesDragTorque (F.C:446-453): tensor row-major xx xy xz yx yy yz zx zy zz
    const double* g = vGrad + 9 * (size_t)c;
    const double s1 = g[7] - g[5], s2 = g[6] - g[2], s3 = g[3] - g[1];
    const double tc = M_PI * (pow(dia, 3.0));
    F[0] = Fx; F[1] = Fy; F[2] = Fz;
    F[3] = tc * (s1 - rec[6]) * nu * rhoF;
    F[4] = tc * (s2 - rec[7]) * nu * rhoF;
    F[5] = tc * (s3 - rec[8]) * nu * rhoF;
}

__global__ void k_source_zero(int nCells, int gaussian, double* __restrict__ uSource, double* __restrict__ alpha,
                              double* __restrict__ uSourceDrag, double* __restrict__ uParticle)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // one thread per scalar of the [N][3] fields
    if (i >= 3 * nCells) return;
    uSource[i] = 0.0;
    if (gaussian) {
        uParticle[i] = 0.0;
        if (i < nCells) {
            alpha[i] = 1.0;
            uSourceDrag[i] = 0.0;
        }
    }
}

__global__ void k_fill(double* __restrict__ a, size_t n, double v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

}  // namespace

// the packed stack entries hold the subtree depth in 5 bits next to a 26-bit node index
static int kdFuncAttrs(fy_ctx* h, const void*)
{
    if (h->nTree >= (1 << 26)) { h->err = "k-d descent: more than 2^26 cells"; return FY_ERR_UNSUPPORTED; }
    return FY_OK;
}

int fyLaunchLocate(fy_ctx* h, const double* d_xyz, int stride, int n, int* d_ids, int* d_cnt)
{
    if (n <= 0) return FY_OK;
    if (int rc = kdFuncAttrs(h, (const void*)k_locate)) return rc;
    k_locate<<<fyGrid(n, KD_BLOCK), KD_BLOCK, KD_SMEM, h->stream>>>(h->dTree, h->nTree, d_xyz, stride, n, h->maxDist, d_ids, d_cnt);
    FY_CHECK_LAUNCH();
    return FY_OK;
}

int fyLaunchFindCell(fy_ctx* h, const double* d_xyz, int stride, int n, int* d_cell)
{
    if (n <= 0) return FY_OK;
    k_find_cell<<<fyGrid(n, 256), 256, 0, h->stream>>>(d_xyz, stride, n, h->boxN[0], h->boxN[1], h->boxN[2],
                                                      h->boxGeom[0], h->boxGeom[1], h->boxGeom[2], h->boxGeom[3],
                                                      h->boxGeom[4], h->boxGeom[5], d_cell);
    FY_CHECK_LAUNCH();
    return FY_OK;
}

// One YadeProc's worth of FoamYade.C:612-628 on device-resident buffers.
namespace {
// sort key: the particle's cell on a gx x gy x gz grid over the mesh bounding box, x fastest like the cells (the mesh's
// own hex box when it has one, up to 256 per direction; 128^3 otherwise)
__global__ void k_particle_keys(const double* __restrict__ pdata, int n, double x0, double y0, double z0, double sx,
                                double sy, double sz, int gx, int gy, int gz, unsigned int* __restrict__ key,
                                int* __restrict__ idx)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double* r = pdata + (size_t)p * 10;
    const int ix = min(gx - 1, max(0, (int)((r[0] - x0) * sx)));
    const int iy = min(gy - 1, max(0, (int)((r[1] - y0) * sy)));
    const int iz = min(gz - 1, max(0, (int)((r[2] - z0) * sz)));
    key[p] = (unsigned int)(ix + gx * (iy + gy * iz));
    idx[p] = p;
}
// lists of the last buffer back in wire order, array-of-structures (parity hook fy_get_last_lists)
__global__ void k_unpermute_lists(int n, const int* __restrict__ perm, const int* __restrict__ cnt, const int* __restrict__ ids,
                                  const double* __restrict__ wts, int* __restrict__ cntOut, int* __restrict__ idsOut,
                                  double* __restrict__ wtsOut)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = perm[t];
    cntOut[p] = cnt[t];
    for (int j = 0; j < FY_MAXLIST; ++j) {
        idsOut[(size_t)p * FY_MAXLIST + j] = ids[(size_t)j * n + t];
        wtsOut[(size_t)p * FY_MAXLIST + j] = wts[(size_t)j * n + t];
    }
}
}  // namespace

// Sorts the buffer's particles by position (radix sort of 21-bit cell keys; the permutation only -- records stay
// where Yade put them).  Result: h->dPerm.
int fySortParticles(fy_ctx* h, const double* d_pdata, int n)
{
    int rc;
    if ((rc = fyReserve(h, h->dKey, (size_t)n))) return rc;
    if ((rc = fyReserve(h, h->dKey2, (size_t)n))) return rc;
    if ((rc = fyReserve(h, h->dIdx, (size_t)n))) return rc;
    if ((rc = fyReserve(h, h->dPerm, (size_t)n))) return rc;
    const double ex = h->bbox[3] - h->bbox[0], ey = h->bbox[4] - h->bbox[1], ez = h->bbox[5] - h->bbox[2];
    int gq[3] = {128, 128, 128};
    if (h->boxN[0] > 0)
        for (int q = 0; q < 3; ++q) gq[q] = std::max(1, std::min(h->boxN[q], 256));
    int bits = 1;
    while ((1LL << bits) < (long long)gq[0] * gq[1] * gq[2]) ++bits;
    k_particle_keys<<<fyGrid(n, 256), 256, 0, h->stream>>>(d_pdata, n, h->bbox[0], h->bbox[1], h->bbox[2],
                                                          ex > 0 ? gq[0] / ex : 0.0, ey > 0 ? gq[1] / ey : 0.0,
                                                          ez > 0 ? gq[2] / ez : 0.0, gq[0], gq[1], gq[2], h->dKey.p, h->dIdx.p);
    FY_CHECK_LAUNCH();
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->dKey.p, h->dKey2.p, h->dIdx.p, h->dPerm.p, n, 0, bits, h->stream);
    if ((rc = fyReserve(h, h->dSortTmp, bytes))) return rc;
    FY_CUDA(cub::DeviceRadixSort::SortPairs(h->dSortTmp.p, bytes, h->dKey.p, h->dKey2.p, h->dIdx.p, h->dPerm.p, n, 0, bits,
                                            h->stream));
    h->launches += 4;
    return FY_OK;
}

int fyUnpermuteLists(fy_ctx* h, int n, int* d_cnt, int* d_ids, double* d_wts)
{
    k_unpermute_lists<<<fyGrid(n, 128), 128, 0, h->stream>>>(n, h->dPerm.p, h->dCnt.p, h->dIds.p, h->dW.p, d_cnt, d_ids, d_wts);
    FY_CHECK_LAUNCH();
    return FY_OK;
}


namespace {
RangeBox rangeBoxOf(const fy_ctx* h)
{
    int cap = RANGE_CAP;
    if (const char* e = std::getenv("FY_RANGE_CAP")) cap = std::max(0, std::min(RANGE_CAP, std::atoi(e)));
    return RangeBox{h->boxN[0], h->boxN[1], h->boxN[2], h->boxGeom[0], h->boxGeom[1], h->boxGeom[2], h->boxGeom[3],
                    h->boxGeom[4], h->boxGeom[5], h->dAxis, std::sqrt(h->maxDist), cap};
}
ForceConst forceConstOf(const fy_ctx* h)
{
    return ForceConst{h->rhoF, h->nu, 1e-09 /* F.H:67 `small` */, h->rhoP, h->deltaT, h->addedMass ? 1 : 0, h->gaussTorque ? 1 : 0};
}
}  // namespace

// Gaussian branch, pass 0: sort, locate + weights + per-cell accumulate (trail lists or full support)
static int gaussLocate(fy_ctx* h, const double* d_pdata, int n, int* d_found)
{
    int rc;
    const GaussConst gc{h->maxDist, 2 * std::pow(h->sigmaInterp, 2), h->interpRangeCu, h->sigmaPi};
    if ((rc = fyReserve(h, h->dCnt, (size_t)n))) return rc;
    if ((rc = fySortParticles(h, d_pdata, n))) return rc;
    if (h->supportFull) {
        if (!h->dAxis) { h->err = "full-support Gaussian mode needs a hex-box mesh whose centres are a tensor product (mesh.boxN)"; return FY_ERR_UNSUPPORTED; }
        if ((rc = fyReserve(h, h->dAllWt, (size_t)n))) return rc;
        k_range_accumulate<<<fyGrid(n, RANGE_WARPS), RANGE_WARPS * 32, 0, h->stream>>>(
            rangeBoxOf(h), d_pdata, n, h->dPerm.p, gc, h->procSerial, h->dCnt.p, h->dAllWt.p, d_found, h->dPvol, h->dUpAcc, h->dStamp);
        FY_CHECK_LAUNCH();
        return FY_OK;
    }
    if ((rc = fyReserve(h, h->dIds, (size_t)n * FY_MAXLIST))) return rc;
    if ((rc = fyReserve(h, h->dW, (size_t)n * FY_MAXLIST))) return rc;
    if ((rc = kdFuncAttrs(h, nullptr))) return rc;
    // the improvement trail in per-thread local memory (default) or in shared memory (FY_LOCATE_SMEM_TRAIL=1: measured SLOWER on
    // B200, 1.35 against 1.27 ms at C2 and 12.9 against 12.3 ms at 256^3 / 10 M -- profiles/r2o_*; kept as an A/B knob)
    static const bool smemTrail = [] { const char* e = std::getenv("FY_LOCATE_SMEM_TRAIL"); return e && std::atoi(e) != 0; }();
    if (smemTrail)
        k_locate_gauss<true><<<fyGrid(n, KD_BLOCK), KD_BLOCK, KD_SMEM, h->stream>>>(h->dTree, h->nTree, d_pdata, n, h->dPerm.p, gc, h->procSerial,
                                                                                    h->dIds.p, h->dCnt.p, h->dW.p, d_found, h->dPvol,
                                                                                    h->dUpAcc, h->dStamp);
    else
        k_locate_gauss<false><<<fyGrid(n, KD_BLOCK), KD_BLOCK, KD_SMEM, h->stream>>>(h->dTree, h->nTree, d_pdata, n, h->dPerm.p, gc, h->procSerial,
                                                                                     h->dIds.p, h->dCnt.p, h->dW.p, d_found, h->dPvol,
                                                                                     h->dUpAcc, h->dStamp);
    FY_CHECK_LAUNCH();
    return FY_OK;
}

// Gaussian branch, pass 2: forces + reaction scatter
static int gaussForce(fy_ctx* h, const double* d_pdata, int n, double* d_force)
{
    const ForceConst fc = forceConstOf(h);
    const bool extra = h->addedMass || h->gaussTorque;
    double* const* f = h->dField;
    if (h->supportFull) {
        const GaussConst gc{h->maxDist, 2 * std::pow(h->sigmaInterp, 2), h->interpRangeCu, h->sigmaPi};
        if (extra)
            k_range_force<true><<<fyGrid(n, RANGE_WARPS), RANGE_WARPS * 32, 0, h->stream>>>(
                rangeBoxOf(h), d_pdata, n, h->dPerm.p, h->dCnt.p, h->dAllWt.p, gc, fc, f[FY_F_U], f[FY_F_ALPHA], f[FY_F_UPARTICLE],
                f[FY_F_GRADP], f[FY_F_DIVT], h->dV, f[FY_F_DDTU], f[FY_F_VGRAD], f[FY_F_USOURCEDRAG], f[FY_F_USOURCE], d_force);
        else
            k_range_force<false><<<fyGrid(n, RANGE_WARPS), RANGE_WARPS * 32, 0, h->stream>>>(
                rangeBoxOf(h), d_pdata, n, h->dPerm.p, h->dCnt.p, h->dAllWt.p, gc, fc, f[FY_F_U], f[FY_F_ALPHA], f[FY_F_UPARTICLE],
                f[FY_F_GRADP], f[FY_F_DIVT], h->dV, f[FY_F_DDTU], f[FY_F_VGRAD], f[FY_F_USOURCEDRAG], f[FY_F_USOURCE], d_force);
    } else if (extra) {
        k_force_gauss<true><<<fyGrid(n, 128), 128, 0, h->stream>>>(
            d_pdata, n, h->dPerm.p, h->dIds.p, h->dCnt.p, h->dW.p, fc, f[FY_F_U], f[FY_F_ALPHA], f[FY_F_UPARTICLE], f[FY_F_GRADP],
            f[FY_F_DIVT], h->dV, f[FY_F_DDTU], f[FY_F_VGRAD], f[FY_F_USOURCEDRAG], f[FY_F_USOURCE], d_force);
    } else {
        k_force_gauss<false><<<fyGrid(n, 128), 128, 0, h->stream>>>(
            d_pdata, n, h->dPerm.p, h->dIds.p, h->dCnt.p, h->dW.p, fc, f[FY_F_U], f[FY_F_ALPHA], f[FY_F_UPARTICLE], f[FY_F_GRADP],
            f[FY_F_DIVT], h->dV, f[FY_F_DDTU], f[FY_F_VGRAD], f[FY_F_USOURCEDRAG], f[FY_F_USOURCE], d_force);
    }
    FY_CHECK_LAUNCH();
    return FY_OK;
}

int fyUnpermuteCounts(fy_ctx* h, int n, int* d_cnt)
{
    k_unpermute_counts<<<fyGrid(n, 256), 256, 0, h->stream>>>(n, h->dPerm.p, h->dCnt.p, d_cnt);
    FY_CHECK_LAUNCH();
    return FY_OK;
}

int fyCouplingProcDevice(fy_ctx* h, const double* d_pdata, int n, int* d_found, double* d_force)
{
    if (!h->propsSet) { h->err = "fy_set_properties must be called first"; return FY_ERR_INVALID; }
    h->lastN = n;
    if (n <= 0) return FY_OK;
    const ForceConst fc = forceConstOf(h);
    const bool prof = h->profiling;
    if (h->gaussian) {
        int rc;
        h->procSerial++;
        if (prof) cudaEventRecord(h->ev[1], h->stream);
        if ((rc = gaussLocate(h, d_pdata, n, d_found))) return rc;
        if (prof) cudaEventRecord(h->ev[2], h->stream);
        k_void_fraction<<<fyGrid(h->nCells, 256), 256, 0, h->stream>>>(h->nCells, h->procSerial, h->dStamp, h->dPvol,
                                                                       h->dUpAcc, h->dV, h->dField[FY_F_ALPHA],
                                                                       h->dField[FY_F_UPARTICLE]);
        FY_CHECK_LAUNCH();
        if (prof) cudaEventRecord(h->ev[3], h->stream);
        if ((rc = gaussForce(h, d_pdata, n, d_force))) return rc;
        if (prof) cudaEventRecord(h->ev[4], h->stream);
    } else {
        if (h->boxN[0] <= 0) { h->err = "point-force mode needs the hex-box findCell (mesh.boxN)"; return FY_ERR_UNSUPPORTED; }
        int rc;
        if ((rc = fyReserve(h, h->dCell, (size_t)n))) return rc;
        const BoxConst b{h->boxN[0], h->boxN[1], h->boxN[2], h->boxGeom[0], h->boxGeom[1], h->boxGeom[2],
                         h->boxGeom[3], h->boxGeom[4], h->boxGeom[5]};
        if (prof) { cudaEventRecord(h->ev[1], h->stream); cudaEventRecord(h->ev[2], h->stream); cudaEventRecord(h->ev[3], h->stream); }
        k_point_force<<<fyGrid(n, 256), 256, 0, h->stream>>>(d_pdata, n, b, fc, h->dField[FY_F_U],
                                                            h->dField[FY_F_VGRAD], h->dV, h->dField[FY_F_USOURCE],
                                                            h->dCell.p, d_found, d_force);
        FY_CHECK_LAUNCH();
        if (prof) cudaEventRecord(h->ev[4], h->stream);
    }
    return FY_OK;
}

// The three passes of fyCouplingProcDevice separately (particle-sharded multi-GPU mode):
// the caller reduces the per-cell partial sums across ranks between the passes.
//   pass 0  locate + weights + per-cell accumulate of THIS rank's particles        -> dPvol, dUpAcc, dStamp
//   pass 1  void fraction from the (reduced) accumulators                          -> alpha, uParticle
//   pass 2  forces of this rank's particles, reaction scattered                    -> uSource, uSourceDrag partials
int fyCouplingPass(fy_ctx* h, int pass, const double* d_pdata, int n, int* d_found, double* d_force)
{
    if (!h->propsSet) { h->err = "fy_set_properties must be called first"; return FY_ERR_INVALID; }
    const ForceConst fc = forceConstOf(h);
    int rc;
    if (pass == 0) {
        h->lastN = n;
        h->procSerial++;                                   // every rank makes the same calls: same serial everywhere
        if (!h->gaussian || n <= 0) return FY_OK;
        if ((rc = gaussLocate(h, d_pdata, n, d_found))) return rc;
    } else if (pass == 1) {
        if (!h->gaussian) return FY_OK;
        k_void_fraction<<<fyGrid(h->nCells, 256), 256, 0, h->stream>>>(h->nCells, h->procSerial, h->dStamp, h->dPvol, h->dUpAcc,
                                                                       h->dV, h->dField[FY_F_ALPHA], h->dField[FY_F_UPARTICLE]);
        FY_CHECK_LAUNCH();
    } else {
        if (n <= 0) return FY_OK;
        if (h->gaussian) {
            return gaussForce(h, d_pdata, n, d_force);
        } else {
            if (h->boxN[0] <= 0) { h->err = "point-force mode needs the hex-box findCell (mesh.boxN)"; return FY_ERR_UNSUPPORTED; }
            if ((rc = fyReserve(h, h->dCell, (size_t)n))) return rc;
            const BoxConst b{h->boxN[0], h->boxN[1], h->boxN[2], h->boxGeom[0], h->boxGeom[1], h->boxGeom[2],
                             h->boxGeom[3], h->boxGeom[4], h->boxGeom[5]};
            k_point_force<<<fyGrid(n, 256), 256, 0, h->stream>>>(d_pdata, n, b, fc, h->dField[FY_F_U], h->dField[FY_F_VGRAD],
                                                                h->dV, h->dField[FY_F_USOURCE], h->dCell.p, d_found, d_force);
        }
        FY_CHECK_LAUNCH();
    }
    return FY_OK;
}

int fySourceZeroDevice(fy_ctx* h)
{
    k_source_zero<<<fyGrid(3LL * h->nCells, 256), 256, 0, h->stream>>>(h->nCells, h->gaussian ? 1 : 0,
                                                                      h->dField[FY_F_USOURCE], h->dField[FY_F_ALPHA],
                                                                      h->dField[FY_F_USOURCEDRAG],
                                                                      h->dField[FY_F_UPARTICLE]);
    FY_CHECK_LAUNCH();
    return FY_OK;
}

// initFields (F.C:56-73): sources zero; alpha = 1 in BOTH modes (F.C:68)
int fyInitCouplingFields(fy_ctx* h)
{
    const size_t N = (size_t)h->nCells;
    FY_CUDA(cudaMemsetAsync(h->dField[FY_F_USOURCE], 0, 3 * N * sizeof(double), h->stream));
    FY_CUDA(cudaMemsetAsync(h->dField[FY_F_UPARTICLE], 0, 3 * N * sizeof(double), h->stream));
    FY_CUDA(cudaMemsetAsync(h->dField[FY_F_USOURCEDRAG], 0, N * sizeof(double), h->stream));
    k_fill<<<fyGrid((long long)N, 256), 256, 0, h->stream>>>(h->dField[FY_F_ALPHA], N, 1.0);
    FY_CHECK_LAUNCH();
    return FY_OK;
}
