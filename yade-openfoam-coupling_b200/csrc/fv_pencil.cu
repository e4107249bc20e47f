// fv_pencil.cu -- lduMatrix solvers of the fluid half on the device: PCG (DIC / diagonal / none) for the
// pressure equation and smoothSolver + symGaussSeidel for the momentum predictor (the solvers the stock
// fvSolution of icoFoam selects; the reference only calls `solve`, icoFoamYade.C:93,125).
//
// Every sequential recurrence runs as a warp-pencil pipeline in the skewed layout of fv_pencil.cuh; the
// Krylov vector kernels (Amul, dots, axpys) run in the same layout so nothing is converted inside the
// iteration.  The per-cell operation order is OpenFOAM's (see fv_box.cuh), global sums are deterministic
// for a fixed launch geometry.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

#include "fv_pencil.cuh"
#include "fv_solver.h"

#ifndef PEN_UNROLL
#define PEN_UNROLL 2          // row-loop unroll factor of the sweeps (measured: 1 -> 0.39, 2 -> 0.34, 4 -> 0.42 ms per DIC sweep pair at 128^3)
#endif
constexpr int kPenUnroll = PEN_UNROLL;

namespace {
constexpr int BLK = 256;
constexpr double FV_VSMALL = 1e-300;
constexpr int PEN_SPIN_LIMIT = 1 << 20;
constexpr int PEN_WMAX = 8;                   // most warps per pencil group
constexpr int PEN_D = 8;                      // rows per flow-control block / helper ring depth
// input prefetch depth (rows) of a sweep with NIN input streams: as deep as ~22 KB of ring per warp allows
#ifndef PEN_DEPTH
#define PEN_DEPTH 8
#endif
#ifndef PEN_AMUL_R
#define PEN_AMUL_R 8           // rows of a slab per warp and trip in the PCG's Amul (1 = the cell-by-cell kernel)
#endif
#ifndef PEN_YPRED
#define PEN_YPRED 1            // re-arm the y slot with a predicated store instead of a one-lane branch (2 % on B200)
#endif
#ifndef PEN_FBLOCK
#define PEN_FBLOCK 8
#endif
__host__ __device__ constexpr int penDepth(int nin) { return nin > 0 ? PEN_DEPTH : PEN_DEPTH; }   // deeper rings were measured slower (more shared memory, no fewer stalls)
constexpr int PEN_CD = 16;                    // z channel depth (rows), a multiple of D
constexpr int PEN_CY = 32;                    // y channel depth (rows): one slot per y-helper lane
constexpr int PEN_GUARD = 64;                 // guard rows around every pencil array (>= 2D + 31)

// ---------------------------------------------------------------------------------------------
// PTX helpers: cp.async staging, polled loads / chain stores (gpu-scope relaxed), shared-memory and DSMEM channels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemU32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 8-byte asynchronous copy global -> shared (LDGSTS): every lane prefetches the words it will consume itself
__device__ __forceinline__ void cpAsync8(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cpAsyncWait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
// shared-memory accesses of the pencil loop: volatile asm keeps them in program order among themselves
__device__ __forceinline__ double ldSharedV(uint32_t p)
{
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(p));
    return v;
}
__device__ __forceinline__ void stSharedV(uint32_t p, double v)
{
    asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(p), "d"(v));
}
// predicated store: no branch, so a one-lane store does not split the warp
__device__ __forceinline__ void stSharedVIf(bool on, uint32_t p, double v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.volatile.shared.f64 [%1], %2;\n\t}" ::"r"((unsigned)on), "r"(p), "d"(v));
}
__device__ __forceinline__ double ldPoll(const double* p)
{
    unsigned long long v;
#ifdef PEN_SYS_SCOPE
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
#else
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
#endif
    return __longlong_as_double((long long)v);
}
__device__ __forceinline__ bool isSent(double v) { return (unsigned long long)__double_as_longlong(v) == PEN_SENT; }
__device__ __forceinline__ void stChain(double* p, double v)
{
#ifdef PEN_SYS_SCOPE
    asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#else
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v));
#endif
}
// thread-block cluster: rank, barrier, distributed shared memory (the z channel between the CTAs of a cluster)
__device__ __forceinline__ uint32_t clusterRank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t clusterSize()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void clusterSync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapToRank(uint32_t localS, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(localS), "r"(rank));
    return r;
}
__device__ __forceinline__ double ldClusterV(uint32_t p)
{
    double v;
    asm volatile("ld.volatile.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(p));
    return v;
}
__device__ __forceinline__ void stClusterV(uint32_t p, double v)
{
    asm volatile("st.volatile.shared::cluster.f64 [%0], %1;" ::"r"(p), "d"(v));
}
__device__ __forceinline__ uint32_t ldClusterU32(uint32_t p)
{
    uint32_t v;
    asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(p));
    return v;
}
__device__ __forceinline__ double sentValue() { return __longlong_as_double((long long)PEN_SENT); }

// ---------------------------------------------------------------------------------------------
// the recurrences.  in[] are the per-cell input streams, chain is the output array the neighbours'
// values are taken from.  Each op is split into
//   pre(a)                : everything that does not depend on a neighbour's new value (runs one step early)
//   cell(c, vx, vy, vz)   : the dependent chain, in OpenFOAM's order (forward sweeps reach a cell through the
//                           faces owned by c-nx*ny, c-nx, c-1, i.e. z, y, x; backward sweeps through the
//                           cell's own faces in descending order, again z, y, x)
//   post(...)             : side outputs
// There are NO neighbour-exists predicates: the coefficient of a missing neighbour is stored as 0 and every
// value a lane can see is finite (pads are 0, inactive lanes produce 0), so the missing term is an exact
// "- 0*v".  (Only the sign of an exact zero result can differ from the sequential loop.)
// ---------------------------------------------------------------------------------------------
struct OpDicD {            // DICPreconditioner::calcReciprocalD: rD[u] -= upper^2 / rD[l]; then rD = 1/rD
    static constexpr int NIN = 4, NC = 4;
    static constexpr bool DOT = false;
    const double* in[NIN];     // dg lowx lowy lowz
    double* chain;             // D before the reciprocal
    double* rD;
    __device__ __forceinline__ void pre(const double (&a)[NIN], double (&c)[NC]) const
    {
        c[0] = a[0];
        c[1] = a[1] * a[1];
        c[2] = a[2] * a[2];
        c[3] = a[3] * a[3];
    }
    __device__ __forceinline__ double cell(const double (&c)[NC], double vx, double vy, double vz, double&) const
    {
        double r = c[0];
        r -= c[3] / (c[3] == 0.0 ? 1.0 : vz);       // no such face: u*u = 0, and v may be a pad's 0
        r -= c[2] / (c[2] == 0.0 ? 1.0 : vy);
        r -= c[1] / (c[1] == 0.0 ? 1.0 : vx);
        return r;
    }
    __device__ __forceinline__ void post(long long pos, bool active, const double (&)[NC], double res, double, double&) const
    {
        if (active) rD[pos] = 1.0 / res;
    }
    __device__ __forceinline__ void fin(FvSolveDev*, double) const {}
};
struct OpDicFwd {          // wA = rD rA;  wA[u] -= rD[u] upper wA[l]   (faces ascending)
    static constexpr int NIN = 5, NC = 4;
    static constexpr bool DOT = false;
    const double* in[NIN];     // rD rA lowx lowy lowz
    double* chain;             // yA
    __device__ __forceinline__ void pre(const double (&a)[NIN], double (&c)[NC]) const
    {
        c[0] = a[0] * a[1];
        c[1] = a[0] * a[2];
        c[2] = a[0] * a[3];
        c[3] = a[0] * a[4];
    }
    __device__ __forceinline__ double cell(const double (&c)[NC], double vx, double vy, double vz, double&) const
    {
        double w = c[0];
        w -= c[3] * vz;
        w -= c[2] * vy;
        w -= c[1] * vx;
        return w;
    }
    __device__ __forceinline__ void post(long long, bool, const double (&)[NC], double, double, double&) const {}
    __device__ __forceinline__ void fin(FvSolveDev*, double) const {}
};
struct OpDicBwd {          // wA[l] -= rD[l] upper wA[u]   (faces descending); accumulates wA.rA; re-arms yA
    static constexpr int NIN = 6, NC = 5;
    static constexpr bool DOT = true;
    const double* in[NIN];     // yA rD upx upy upz rA
    double* chain;             // zA
    double* y;
    __device__ __forceinline__ void pre(const double (&a)[NIN], double (&c)[NC]) const
    {
        c[0] = a[0];
        c[1] = a[1] * a[2];
        c[2] = a[1] * a[3];
        c[3] = a[1] * a[4];
        c[4] = a[5];
    }
    __device__ __forceinline__ double cell(const double (&c)[NC], double vx, double vy, double vz, double&) const
    {
        double w = c[0];
        w -= c[3] * vz;
        w -= c[2] * vy;
        w -= c[1] * vx;
        return w;
    }
    __device__ __forceinline__ void post(long long pos, bool active, const double (&c)[NC], double res, double, double& acc) const
    {
        acc += res * c[4];
        if (active) y[pos] = sentValue();
    }
    __device__ __forceinline__ void fin(FvSolveDev* st, double tot) const
    {
        if (!st) return;
        st->wArAold = st->wArA;
        st->wArA = tot;
        st->beta = st->wArA / st->wArAold;
    }
};
struct OpGsFwd {           // GaussSeidelSmoother forward sweep: new values below; old values above arrive pre-multiplied (e)
    static constexpr int NIN = 8, NC = 8;
    static constexpr bool DOT = false;
    const double* in[NIN];     // b lowx lowy lowz ex ey ez dg
    double* chain;             // psi after the forward sweep
    double* bPrime;
    double* psiOld;            // consumed by k_pen_gs_upper before the sweep: re-armed here for the backward sweep
    __device__ __forceinline__ void pre(const double (&a)[NIN], double (&c)[NC]) const
    {
#pragma unroll
        for (int x = 0; x < NIN; ++x) c[x] = a[x];
    }
    __device__ __forceinline__ double cell(const double (&c)[NC], double vx, double vy, double vz, double& side) const
    {
        double bp = c[0];
        bp -= c[3] * vz;
        bp -= c[2] * vy;
        bp -= c[1] * vx;
        side = bp;
        double x = bp;
        x -= c[4];                                  // upper[x+] psi_old[c+1]   (0 when there is no such face)
        x -= c[5];
        x -= c[6];
        return x / c[7];
    }
    __device__ __forceinline__ void post(long long pos, bool active, const double (&)[NC], double, double side, double&) const
    {
        if (active) {
            bPrime[pos] = side;
            psiOld[pos] = sentValue();
        }
    }
    __device__ __forceinline__ void fin(FvSolveDev*, double) const {}
};
struct OpGsBwd {           // GaussSeidelSmoother backward sweep (own faces in ascending order: x, y, z)
    static constexpr int NIN = 5, NC = 5;
    static constexpr bool DOT = false;
    const double* in[NIN];     // bPrime upx upy upz dg
    double* chain;             // psi
    double* mid;               // forward-sweep values: dead now, re-armed for the next iteration
    __device__ __forceinline__ void pre(const double (&a)[NIN], double (&c)[NC]) const
    {
#pragma unroll
        for (int x = 0; x < NIN; ++x) c[x] = a[x];
    }
    __device__ __forceinline__ double cell(const double (&c)[NC], double vx, double vy, double vz, double&) const
    {
        double x = c[0];
        x -= c[1] * vx;
        x -= c[2] * vy;
        x -= c[3] * vz;
        return x / c[4];
    }
    __device__ __forceinline__ void post(long long pos, bool active, const double (&)[NC], double, double, double&) const
    {
        if (active) mid[pos] = sentValue();
    }
    __device__ __forceinline__ void fin(FvSolveDev*, double) const {}
};

struct PenCtl {
    unsigned int* ticket;      // [0] next ticket  [1] warps finished
    int* error;                // set when a poll gave up
    double* partial;           // per-warp partial sums
    FvSolveDev* st;            // may be null
    int dbg;                   // debug switches (FY_PENCIL_DBG)
    unsigned long long* trace; // optional [warps][8] time stamps / wait cycles
    double* distOut;           // decomposed solve (NCCL path): the sweep's global sum goes here instead of to Op::fin
    PeerDev* peer;             // decomposed solve (peer-memory path): the sum is all-reduced over the ranks before Op::fin
};

// One CTA per pencil group (j-block jb, plane group kq): W COMPUTE warps, each owning ONE k-plane of the
// 32-lane j-block, plus two HELPER warps.  Per step a compute lane does one cell.  It prefetches its own
// inputs (8-byte cp.async into a private shared-memory ring, D rows ahead) and reads them one step early,
// so that the neighbour-independent products overlap the previous step's dependent chain.  Neighbour
// values that cross warps arrive through shared-memory channels of sentinel-armed slots (data and flag in
// one 8-byte word, per-lane producer/consumer):
//   z channel w : the 32 values of plane w-1's row -- written by compute warp w-1, or, for the group's
//                 first plane, by the Z HELPER, which polls the output array of the group behind in L2;
//   y channel w : the edge lane's y-neighbour -- written by the Y HELPER, which polls the last lane of the
//                 neighbouring j-block's rows in L2 (one helper lane per plane).
// So the compute warps never touch L2 for a dependency and never spin on anything but shared memory.
// Tickets are handed out in dependency order (a group's producers always hold smaller tickets), so the
// pipeline cannot deadlock whatever the residency.  All memory operations of the loops are volatile asm
// (kept in program order, which IS the software pipeline); the arithmetic between them is left to the
// scheduler.  The loops have no bounds tests: Tp is a multiple of 2D and every array carries guard rows.
struct PenWarp {               // per-warp constants of one sweep
    uint32_t ringS, zInS, zOutS, yInS;
    int s0, nx, Tp, row00;
    unsigned napNs;
    bool zOut, zRemote, edge;       // zRemote: the z channel written lives in the next CTA of the cluster (DSMEM)
    long long slab;
};

__device__ __forceinline__ void penStampAt(unsigned long long* tr, int t0, int Tp)
{
    if (!tr) return;
    const int wh = 8 + t0 / PEN_D;
    if (wh < 32 && (threadIdx.x & 31) == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tr[wh] = t;
    }
}

template <class Op, bool REV, bool ZIN, bool YIN>
__device__ __forceinline__ void penSweep(const Op& op, const PenWarp& w, double& acc, unsigned long long* tr)
{
    constexpr int NIN = Op::NIN, NC = Op::NC, D = penDepth(Op::NIN), FB = PEN_FBLOCK, CD = PEN_CD;
    constexpr unsigned int FULL = 0xffffffffu;
    constexpr int RS = REV ? -32 : 32;                     // row stride in sweep order (doubles)
    // running pointers (one 64-bit add per stream and row; no per-row address arithmetic from scratch)
    const double* inP[NIN];
#pragma unroll
    for (int x = 0; x < NIN; ++x) inP[x] = op.in[x] + w.slab + w.row00;
    double* chainP = op.chain + w.slab + w.row00;
    long long pos = w.slab + w.row00;

    // row s of the sweep is fetched by commit group s into ring slot s mod D
#pragma unroll
    for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int x = 0; x < NIN; ++x) cpAsync8(w.ringS + (d * NIN + x) * 256, inP[x] + d * RS);
        cpAsyncCommit();
    }
#pragma unroll
    for (int x = 0; x < NIN; ++x) inP[x] += D * RS;        // from here on: the row D ahead of the current one
    double prev = 0.0, cn[NC];
    cpAsyncWait<D - 1>();
    {
        double a[NIN];
#pragma unroll
        for (int x = 0; x < NIN; ++x) a[x] = ldSharedV(w.ringS + x * 256);
        op.pre(a, cn);
    }
    double vzN = ZIN ? ldSharedV(w.zInS) : 0.0;            // channel reads run one row ahead as well
    double vyN = YIN ? ldSharedV(w.yInS) : 0.0;
    uint32_t rs = 0;                                       // ring slot of the current row (bytes)
    int act = -w.s0;                                       // t - s0: the lane's cell index in sweep order
    // The row loop is deliberately NOT unrolled: one warp executes it alone, and a body that overflows the
    // instruction cache costs more than the few slot-index instructions saved.
#pragma unroll kPenUnroll
    for (int t = 0; t < w.Tp; ++t) {
        const uint32_t rsN = (rs + NIN * 256 == D * NIN * 256) ? 0u : rs + NIN * 256;     // next row's ring slot
        const uint32_t cs = (uint32_t)(t & (CD - 1));                         // channel slots
        const uint32_t csN = (uint32_t)((t + 1) & (CD - 1));
        const uint32_t ys = (uint32_t)(t & (PEN_CY - 1)), ysN = (uint32_t)((t + 1) & (PEN_CY - 1));
        if ((t & (FB - 1)) == 0) {
#ifdef PEN_PROFILE
            penStampAt(tr, t, w.Tp);
#endif
            if (w.zOut) {                                                     // flow control, once per block of D rows
                int spin = 0;
                if (w.zRemote) { while (!isSent(ldClusterV(w.zOutS + (cs + FB - 1) * 256)) && ++spin < PEN_SPIN_LIMIT) {} }
                else { while (!isSent(ldSharedV(w.zOutS + (cs + FB - 1) * 256)) && ++spin < PEN_SPIN_LIMIT) {} }
            }
        }
        // (A) next row's inputs: read them now, use them after this row's chain
        cpAsyncWait<D - 2>();
        double an[NIN];
#pragma unroll
        for (int x = 0; x < NIN; ++x) an[x] = ldSharedV(w.ringS + rsN + x * 256);
        // (B) refill the slot this row came from (its values already sit in cn)
#pragma unroll
        for (int x = 0; x < NIN; ++x) {
            cpAsync8(w.ringS + rs + x * 256, inP[x]);
            inP[x] += RS;
        }
        cpAsyncCommit();
        // (C) the dependent chain
        const bool active = (unsigned)act < (unsigned)w.nx;
        const double vx = prev;
        double vy = REV ? __shfl_down_sync(FULL, prev, 1) : __shfl_up_sync(FULL, prev, 1);
        double vz = 0.0;
        if (ZIN) {
            double v = vzN;
            int spin = 0;
            while (isSent(v) && ++spin < PEN_SPIN_LIMIT) {
                if (w.napNs) __nanosleep(w.napNs);         // the producer is still on this row: leave it the issue slots
                v = ldSharedV(w.zInS + cs * 256);
            }
            stSharedV(w.zInS + cs * 256, sentValue());
            vzN = ldSharedV(w.zInS + csN * 256);
            vz = v;
        }
        if (YIN) {
            double v = vyN;
            int spin = 0;
#if PEN_YPRED
            if (w.edge && isSent(v)) {                     // only the edge lane owns the slot; rare (the helper runs 32 rows ahead)
                do { v = ldSharedV(w.yInS + ys * 8); } while (isSent(v) && ++spin < PEN_SPIN_LIMIT);
            }
            stSharedVIf(w.edge, w.yInS + ys * 8, sentValue());
#else
            while (w.edge && isSent(v) && ++spin < PEN_SPIN_LIMIT) v = ldSharedV(w.yInS + ys * 8);   // only the edge lane owns the slot
            if (w.edge) stSharedV(w.yInS + ys * 8, sentValue());
#endif
            vyN = ldSharedV(w.yInS + ysN * 8);
            vy = w.edge ? v : vy;
        }
        double side = 0.0;
        double res = op.cell(cn, vx, vy, vz, side);
        res = active ? res : 0.0;
        if (w.zOut) {
            if (w.zRemote) stClusterV(w.zOutS + cs * 256, res);
            else stSharedV(w.zOutS + cs * 256, res);
        }
        stChain(chainP, res);
        op.post(pos, active, cn, res, side, acc);
        prev = res;
        // (D) neighbour-independent products of the next row
        op.pre(an, cn);
        chainP += RS;
        pos += RS;
        rs = rsN;
        ++act;
    }
    cpAsyncWait<0>();
}

// Z helper: streams the rows of the plane behind the group (another CTA's output, in L2) into z channel 0.
// Eight rows are kept in flight; when the wanted row is still armed, every armed row is re-requested at once.
template <bool REV, int CD = PEN_CD>
__device__ __forceinline__ void penHelpZ(const double* zRow0, uint32_t chanS, int Tp, int& fail, unsigned long long* tr, int dbg)
{
    constexpr int D = PEN_D;
    constexpr int RS = REV ? -32 : 32;
    const double* zB = zRow0;
    double r[D];
#pragma unroll
    for (int d = 0; d < D; ++d) r[d] = ldPoll(zB + d * RS);
    for (int t0 = 0; t0 < Tp; t0 += D) {
        const uint32_t cb = (uint32_t)(t0 & (CD - 1)) * 256;
        penStampAt(tr, t0, Tp);
        {
            int spin = 0;
            while (!isSent(ldSharedV(chanS + cb + (D - 1) * 256)) && ++spin < PEN_SPIN_LIMIT) {}
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            int spin = 0;
            while (isSent(r[d]) && !fail) {
                // one L2 round trip costs several row times: re-request EVERY row of the ring that has not
                // arrived, so that a round trip brings all the rows produced meanwhile
#pragma unroll
                for (int x = 0; x < D; ++x) {
                    const int off = (x >= d ? x : x + D) * RS;     // rows d.. belong to this block, rows < d to the next
                    if (isSent(r[x])) r[x] = ldPoll(zB + off);
                }
                if (++spin > PEN_SPIN_LIMIT) fail = 1;
            }
            stSharedV(chanS + cb + d * 256, r[d]);
            r[d] = ldPoll(zB + (d + D) * RS);
        }
        zB += D * RS;
    }
}

// Y helpers: one warp per compute warp.  It fetches the y-neighbour of the plane's edge lane -- the last lane
// of the neighbouring j-block's rows, another CTA's output in L2 -- 32 rows at a time: lane l polls row
// 32b + l and drops it into slot l of the plane's y channel as soon as the consumer has taken the previous
// block's row from it.  Poll and hand-over are ONE loop: a lane whose row has arrived delivers it in the same
// trip while the other lanes keep polling (two loops in a row would park the early lanes at the first loop's
// reconvergence point until the block's last row exists, i.e. add up to 31 row times to every j-block hop).
template <bool REV>
__device__ __forceinline__ void penHelpY(const double* yRow0, uint32_t chanS, int Tp, int& fail, unsigned napNs)
{
    constexpr int RS = REV ? -32 : 32;
    const int lane = threadIdx.x & 31;
    const uint32_t slot = chanS + (uint32_t)lane * 8;
    for (int b0 = 0; b0 < Tp; b0 += PEN_CY) {
        const double* a = yRow0 + (long long)(b0 + lane) * RS;
        double v = ldPoll(a);
        int spin = 0;
        bool pending = true;
        while (pending && !fail) {
            if (isSent(v)) {
                if (napNs) __nanosleep(napNs);             // the row is still being produced: do not hog the LSU
                v = ldPoll(a);
            } else if (isSent(ldSharedV(slot))) {
                stSharedV(slot, v);
                pending = false;
            }
            if (++spin > PEN_SPIN_LIMIT) fail = 1;
        }
    }
}

template <class Op, bool REV>
__global__ void __launch_bounds__(32 * (2 * PEN_WMAX + 1), 1) k_pencil(PencilGeom g, Op op, PenCtl ctl)
{
    constexpr int NIN = Op::NIN, D = penDepth(Op::NIN), CD = PEN_CD;
    constexpr unsigned int FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char penSmem[];
    __shared__ unsigned int shTicket;
    if (ctl.st && ctl.st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5, W = (NW - 1) >> 1;
    // warps: [0,W) compute | W: z helper | (W, 2W]: y helpers
    // shared memory: input rings [W][D][NIN][32] | z channels [W][CD][32] (channel w feeds compute warp w) | y channels [W][CY]
    double* const zChan = reinterpret_cast<double*>(penSmem) + (size_t)W * (D * NIN * 32);
    double* const yChan = zChan + (size_t)W * (CD * 32);
    for (int x = threadIdx.x; x < W * (CD * 32 + PEN_CY); x += blockDim.x) zChan[x] = sentValue();
    // one ticket per cluster; the C CTAs of a cluster take C consecutive plane groups of one j-block
    const int C = (int)clusterSize(), rank = (int)clusterRank();
    if (rank == 0 && threadIdx.x == 0) shTicket = atomicAdd(ctl.ticket, 1u);
    clusterSync();                                         // channels armed and the ticket taken, cluster-wide
    const unsigned int tk = ldClusterU32(mapToRank(smemU32(&shTicket), 0));
    const int nKQ = (g.nz + W - 1) / W, nCl = (nKQ + C - 1) / C;
    int cl = (int)tk / g.nJB, jb = (int)tk - cl * g.nJB;
    if (REV) { cl = nCl - 1 - cl; jb = g.nJB - 1 - jb; }
    const int kq = REV ? cl * C + C - 1 - rank : cl * C + rank;           // may be >= nKQ: a CTA without planes
    const int slotId = (kq * g.nJB + jb) * NW + warp;
    constexpr int KS = REV ? -1 : 1;
    auto planeOf = [&](int wq) { return REV ? kq * W + W - 1 - wq : kq * W + wq; };
    auto stamp = [&](int wh) {
        if (ctl.trace && lane == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            ctl.trace[(size_t)slotId * 32 + wh] = t;
        }
    };
    stamp(0);
    unsigned long long* const trw = ctl.trace ? ctl.trace + (size_t)slotId * 32 : nullptr;
    double acc = 0.0;
    int fail = 0;
    const bool yCol = REV ? jb < g.nJB - 1 : jb > 0;       // this j-block has a y-producer block
    // edge lane: (i, j-1) sits in slab jb-1 at row m+31, lane 31; (i, j+1) in slab jb+1 at row m-31, lane 0
    const long long yOff = REV ? ((long long)g.Tp - 31) * 32 - 31 : -(((long long)g.Tp - 31) * 32 - 31);
    const int row00 = REV ? (g.Tp - 1) * 32 : 0;
    constexpr int EDGE = REV ? 31 : 0;
    if (warp < W) {
        const int k = planeOf(warp);
        if (k < g.nz) {
            const int j = jb * 32 + lane;
            const bool jvalid = j < g.ny;
            PenWarp w;
            w.ringS = smemU32(reinterpret_cast<double*>(penSmem) + (size_t)warp * (D * NIN * 32) + lane);
            w.zInS = smemU32(zChan + (size_t)warp * (CD * 32) + lane);
            w.zOutS = w.zInS + CD * 256;
            w.yInS = smemU32(yChan + (size_t)warp * PEN_CY);
            // Sweep row s touches memory row m = s (forward) or Tp-1-s (backward); lane l then sits on the x index
            // i = m - l.  s - s0 counts the lane's cells in sweep order: active for 0 <= s - s0 < nx.
            w.s0 = jvalid ? (REV ? g.Tp - g.nx - lane : lane) : (1 << 30);
            w.nx = g.nx;
            w.Tp = g.Tp;
            w.row00 = row00;
            w.slab = (((long long)k * g.nJB + jb) * g.Tp) * 32 + lane;
            const int kBehind = k - KS, kAhead = k + KS;
            const bool zin = kBehind >= 0 && kBehind < g.nz;
            w.zRemote = warp == W - 1;
            w.zOut = kAhead >= 0 && kAhead < g.nz && (warp < W - 1 || rank < C - 1);
            if (w.zRemote) w.zOutS = mapToRank(smemU32(zChan + lane), (uint32_t)(rank < C - 1 ? rank + 1 : rank));
            w.edge = lane == EDGE;
            w.napNs = (unsigned)(ctl.dbg >> 20);
            if (zin && yCol) penSweep<Op, REV, true, true>(op, w, acc, trw);
            else if (zin) penSweep<Op, REV, true, false>(op, w, acc, trw);
            else if (yCol) penSweep<Op, REV, false, true>(op, w, acc, trw);
            else penSweep<Op, REV, false, false>(op, w, acc, trw);
        }
    } else if (warp == W) {
        const int k0 = planeOf(0), kBehind = k0 - KS;
        if (rank == 0 && k0 < g.nz && kBehind >= 0 && kBehind < g.nz)
            penHelpZ<REV>(op.chain + (((long long)kBehind * g.nJB + jb) * g.Tp) * 32 + lane + row00, smemU32(zChan + lane), g.Tp, fail, trw, ctl.dbg);

    } else if (yCol) {
        const int q = warp - W - 1;
        const int k = planeOf(q);
        if (k < g.nz)
            penHelpY<REV>(op.chain + (((long long)k * g.nJB + jb) * g.Tp) * 32 + EDGE + yOff + row00, smemU32(yChan + (size_t)q * PEN_CY), g.Tp, fail, (unsigned)((ctl.dbg >> 8) & 0xfff));
    }

    stamp(3);
    if (Op::DOT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL, acc, o);
        if (lane == 0) ctl.partial[slotId] = acc;
    }
    fail = __any_sync(FULL, fail);
    unsigned int last = 0;
    if (lane == 0) {
        if (fail) atomicExch(ctl.error, 1);
        __threadfence();
        last = (atomicAdd(ctl.ticket + 1, 1u) == gridDim.x * NW - 1) ? 1u : 0u;
    }
    last = __shfl_sync(FULL, last, 0);
    if (!last) return;
    __threadfence();
    if (Op::DOT) {
        const volatile double* p = ctl.partial;
        double x = 0.0;
        for (unsigned int b = lane; b < gridDim.x * NW; b += 32) x += p[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(FULL, x, o);
        if (lane == 0) op.fin(ctl.st, x);
    }
    if (lane == 0) {
        ctl.ticket[0] = 0u;
        ctl.ticket[1] = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// layout conversion (natural x-fastest cell order <-> pencil layout)
// ---------------------------------------------------------------------------------------------
#define PEN_ROW_LOOP(g, c)                                                                                  \
    for (long long l_ = (long long)blockIdx.x * (BLK / 32) + (threadIdx.x >> 5); l_ < (g).nLoc;            \
         l_ += (long long)gridDim.x * (BLK / 32))                                                           \
        for (PenCell c = penDecode((g), penGlobalRow((g), l_), threadIdx.x & 31); c.valid; c.valid = false)

__device__ __forceinline__ int penNat(const PencilGeom& g, const PenCell& c) { return c.i + g.nx * (c.j + g.ny * c.k); }

#include "fv_pencil2.cuh"

// matrix: dg [N], lo / up owner slots [3N] (natural) -> PenMatrix
__global__ void __launch_bounds__(BLK)
k_pen_matrix(PencilGeom g, const double* __restrict__ dg, const double* __restrict__ lo, const double* __restrict__ up,
             PenMatrix M)
{
    const int sd[3] = {1, g.nx, g.nx * g.ny};
    PEN_ROW_LOOP(g, c) {
        const int n = penNat(g, c);
        const int v[3] = {c.i, c.j, c.k}, nd[3] = {g.nx, g.ny, g.nz};
        if (dg) M.dg[c.pos] = dg[n];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (M.low[d]) M.low[d][c.pos] = v[d] > 0 ? lo[(size_t)d * g.N + n - sd[d]] : 0.0;
            if (M.up[d]) M.up[d][c.pos] = v[d] < nd[d] - 1 ? up[(size_t)d * g.N + n] : 0.0;
        }
    }
}
__global__ void __launch_bounds__(BLK) k_pen_from_nat(PencilGeom g, const double* __restrict__ nat, double* __restrict__ pen)
{
    PEN_ROW_LOOP(g, c) pen[c.pos] = nat[penNat(g, c)];
}
__global__ void __launch_bounds__(BLK) k_pen_to_nat(PencilGeom g, const double* __restrict__ pen, double* __restrict__ nat)
{
    PEN_ROW_LOOP(g, c) nat[penNat(g, c)] = pen[c.pos];
}
// arms up to three chain arrays with the sentinel
__global__ void __launch_bounds__(BLK) k_pen_arm(PencilGeom g, double* a0, double* a1, double* a2, const FvSolveDev* st)
{
    const double sv = sentValue();
    PEN_ROW_LOOP(g, c) {
        if (a0) a0[c.pos] = sv;
        if (a1) a1[c.pos] = sv;
        if (a2) a2[c.pos] = sv;
    }
}

// ---------------------------------------------------------------------------------------------
// lduMatrix kernels in pencil layout
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double penAmulCell(const PencilGeom& g, const PenMatrix& M, const double* __restrict__ x,
                                              const PenCell& c)
{
    const long long p = c.pos;
    double a = M.dg[p] * x[p];
    if (c.k > 0) a += M.low[2][p] * x[p - g.zStride];
    if (c.j > 0) a += M.low[1][p] * x[penYm(g, c)];
    if (c.i > 0) a += M.low[0][p] * x[p - 32];
    if (c.i < g.nx - 1) a += M.up[0][p] * x[p + 32];
    if (c.j < g.ny - 1) a += M.up[1][p] * x[penYp(g, c)];
    if (c.k < g.nz - 1) a += M.up[2][p] * x[p + g.zStride];
    return a;
}
__device__ __forceinline__ double penSumACell(const PencilGeom& g, const PenMatrix& M, const PenCell& c)
{
    const long long p = c.pos;
    double a = M.dg[p];
    if (c.k > 0) a += M.low[2][p];
    if (c.j > 0) a += M.low[1][p];
    if (c.i > 0) a += M.low[0][p];
    if (c.i < g.nx - 1) a += M.up[0][p];
    if (c.j < g.ny - 1) a += M.up[1][p];
    if (c.k < g.nz - 1) a += M.up[2][p];
    return a;
}

// the same sums for a SYMMETRIC matrix from the upper coefficients alone (the coefficient towards a lower neighbour
// is that neighbour's own upper coefficient: same values, same order) -- the `low` arrays then serve the
// preconditioner only, whose coupling across a slab boundary is dropped in a decomposed solve
__device__ __forceinline__ double penAmulCellSym(const PencilGeom& g, const PenMatrix& M, const double* __restrict__ x,
                                                 const PenCell& c);
__device__ __forceinline__ double penSumACellSym(const PencilGeom& g, const PenMatrix& M, const PenCell& c)
{
    const long long p = c.pos;
    double a = M.dg[p];
    if (c.k > 0) a += M.up[2][p - g.zStride];
    if (c.j > 0) a += M.up[1][penYm(g, c)];
    if (c.i > 0) a += M.up[0][p - 32];
    if (c.i < g.nx - 1) a += M.up[0][p];
    if (c.j < g.ny - 1) a += M.up[1][p];
    if (c.k < g.nz - 1) a += M.up[2][p];
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

// what the kernels' global sums feed, as functions: a decomposed solve (fv_dist.cu) all-reduces the ranks' partial sums
// first and then runs them from k_pen_fin
enum { PEN_FIN_AVG = 0, PEN_FIN_INIT, PEN_FIN_WARA, PEN_FIN_WAPA, PEN_FIN_UPDATE };
__device__ __forceinline__ void penFinAvg(FvSolveDev* st, const double* t, int N) { st->avg = t[0] / N; }
__device__ __forceinline__ void penFinInit(FvSolveDev* st, const double* t)
{
    st->normFactor = t[0] + 1e-20;
    st->initRes = t[1] / st->normFactor;
    st->finalRes = st->initRes;
    st->done = fvConverged(st) ? 1 : 0;
}
__device__ __forceinline__ void penFinWArA(FvSolveDev* st, const double* t)
{
    st->wArAold = st->wArA;
    st->wArA = t[0];
    st->beta = st->wArA / st->wArAold;
}
__device__ __forceinline__ void penFinWApA(FvSolveDev* st, const double* t)
{
    st->wApA = t[0];
    if (fabs(t[0]) / st->normFactor < FV_VSMALL) { st->singular = 1; st->done = 1; }
    else st->alpha = st->wArA / t[0];
}
__device__ __forceinline__ void penFinUpdate(FvSolveDev* st, const double* t)
{
    st->finalRes = t[0] / st->normFactor;
    const bool cont = st->nIter < st->maxIter;                    // nIterations++ < maxIter_
    st->nIter += 1;
    if (!cont || fvConverged(st)) st->done = 1;
}
__global__ void k_pen_fin(int which, const double* t, FvSolveDev* st, int N)
{
    if (st->done && which != PEN_FIN_AVG && which != PEN_FIN_INIT) return;
    switch (which) {
    case PEN_FIN_AVG: penFinAvg(st, t, N); break;
    case PEN_FIN_INIT: penFinInit(st, t); break;
    case PEN_FIN_WARA: penFinWArA(st, t); break;
    case PEN_FIN_WAPA: penFinWApA(st, t); break;
    default: penFinUpdate(st, t); break;
    }
}

// gAverage(psi)
__global__ void __launch_bounds__(BLK) k_pen_avg(PencilGeom g, const double* __restrict__ psi, FvRed red, FvSolveDev* st)
{
    double v[1] = {0.0};
    PEN_ROW_LOOP(g, c) v[0] += psi[c.pos];
    const int N = g.N;
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) { penFinAvg(st, t, N); });
}

// rA = source - A psi;  normFactor = sum(|Apsi - sumA avg| + |source - sumA avg|) + 1e-20;
// initialResidual = sum|rA| / normFactor                              [OF-6 lduMatrix::solver::normFactor]
template <bool SYM>
__global__ void __launch_bounds__(BLK)
k_pen_solve_init(PencilGeom g, PenMatrix M, const double* __restrict__ b, const double* __restrict__ psi,
                 double* __restrict__ rA, FvRed red, FvSolveDev* st)
{
    double v[2] = {0.0, 0.0};
    const double avg = st->avg;
    PEN_ROW_LOOP(g, c) {
        const double Apsi = SYM ? penAmulCellSym(g, M, psi, c) : penAmulCell(g, M, psi, c);
        const double t = (SYM ? penSumACellSym(g, M, c) : penSumACell(g, M, c)) * avg;
        const double r = b[c.pos] - Apsi;
        if (rA) rA[c.pos] = r;
        v[0] += fabs(Apsi - t) + fabs(b[c.pos] - t);
        v[1] += fabs(r);
    }
    fvGridReduce<2, false, BLK>(v, red, [=](const double* t) { penFinInit(st, t); });
}

// lduMatrix::residual + gSumMag (smoothSolver's convergence test after each sweep)
__global__ void __launch_bounds__(BLK)
k_pen_residual(PencilGeom g, PenMatrix M, const double* __restrict__ b, const double* __restrict__ psi, FvRed red,
               FvSolveDev* st)
{
    if (st->done) return;
    double v[1] = {0.0};
    PEN_ROW_LOOP(g, c) {
        const long long p = c.pos;
        double r = b[p] - M.dg[p] * psi[p];
        if (c.k > 0) r -= M.low[2][p] * psi[p - g.zStride];
        if (c.j > 0) r -= M.low[1][p] * psi[penYm(g, c)];
        if (c.i > 0) r -= M.low[0][p] * psi[p - 32];
        if (c.i < g.nx - 1) r -= M.up[0][p] * psi[p + 32];
        if (c.j < g.ny - 1) r -= M.up[1][p] * psi[penYp(g, c)];
        if (c.k < g.nz - 1) r -= M.up[2][p] * psi[p + g.zStride];
        v[0] += fabs(r);
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) {
        st->finalRes = t[0] / st->normFactor;
        st->nIter += 1;                                               // nSweeps = 1
        st->done = (!(st->nIter < st->maxIter) || fvConverged(st)) ? 1 : 0;
    });
}

// Gauss-Seidel: the products upper*psi_old of the faces a cell owns, taken before the forward sweep
// overwrites anything (the sweep subtracts them in the same x+, y+, z+ order)
__global__ void __launch_bounds__(BLK)
k_pen_gs_upper(PencilGeom g, PenMatrix M, const double* __restrict__ psi, double* __restrict__ ex, double* __restrict__ ey,
               double* __restrict__ ez, const FvSolveDev* st)
{
    if (st->done) return;
    PEN_ROW_LOOP(g, c) {
        const long long p = c.pos;
        ex[p] = c.i < g.nx - 1 ? M.up[0][p] * psi[p + 32] : 0.0;
        ey[p] = c.j < g.ny - 1 ? M.up[1][p] * psi[penYp(g, c)] : 0.0;
        ez[p] = c.k < g.nz - 1 ? M.up[2][p] * psi[p + g.zStride] : 0.0;
    }
}

// diagonal / no preconditioner: zA = rD rA (or rA) with the wA.rA dot
__global__ void __launch_bounds__(BLK)
k_pen_precond_diag(PencilGeom g, const double* __restrict__ rD, const double* __restrict__ rA, double* __restrict__ zA,
                   FvRed red, FvSolveDev* st)
{
    if (st->done) return;
    double v[1] = {0.0};
    PEN_ROW_LOOP(g, c) {
        const double w = rD ? rD[c.pos] * rA[c.pos] : rA[c.pos];
        zA[c.pos] = w;
        v[0] += w * rA[c.pos];
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) { penFinWArA(st, t); });
}
__global__ void __launch_bounds__(BLK) k_pen_recip(PencilGeom g, const double* __restrict__ a, double* __restrict__ out)
{
    PEN_ROW_LOOP(g, c) out[c.pos] = 1.0 / a[c.pos];
}

// pA = zA (first iteration) | zA + beta pA;  re-arms zA for the next backward sweep
__global__ void __launch_bounds__(BLK)
k_pen_dir(PencilGeom g, double* __restrict__ zA, double* __restrict__ pA, const FvSolveDev* st, PeerDev* peer)
{
    if (st->done) return;
    const bool first = st->nIter == 0;
    const double beta = st->beta;
    const double sv = sentValue();
    if (!peer) {
        PEN_ROW_LOOP(g, c) {
            const double z = zA[c.pos];
            pA[c.pos] = first ? z : z + beta * pA[c.pos];
            zA[c.pos] = sv;
        }
        return;
    }
    // decomposed solve: the boundary cells of this rank's region go into the neighbours' ghost cells as well (same
    // offset: every rank keeps the global layout), then the neighbours are told (fv_peer.cuh)
    double* const zlo = peer->pa[0];
    double* const zhi = peer->pa[1];
    double* const ylo = peer->pa[2];
    double* const yhi = peer->pa[3];
    const int jLo = g.jbLo * 32, jHi = g.jbHi * 32 - 1;
    bool wrote = false;
    PEN_ROW_LOOP(g, c) {
        const double z = zA[c.pos];
        const double v = first ? z : z + beta * pA[c.pos];
        pA[c.pos] = v;
        zA[c.pos] = sv;
        if (zlo && c.k == g.kLo) { zlo[c.pos] = v; wrote = true; }
        if (zhi && c.k == g.kHi - 1) { zhi[c.pos] = v; wrote = true; }
        if (ylo && c.j == jLo) { ylo[c.pos] = v; wrote = true; }
        if (yhi && c.j == jHi) { yhi[c.pos] = v; wrote = true; }
    }
    peerHaloPublish(peer, wrote);
}

// Amul of a SYMMETRIC matrix: the coefficient towards a lower neighbour is that neighbour's own upper coefficient,
// which the neighbouring rows stream anyway -- three coefficient arrays less to fetch from HBM (same values, same sums)
__device__ __forceinline__ double penAmulCellSym(const PencilGeom& g, const PenMatrix& M, const double* __restrict__ x,
                                                 const PenCell& c)
{
    const long long p = c.pos;
    double a = M.dg[p] * x[p];
    if (c.k > 0) a += M.up[2][p - g.zStride] * x[p - g.zStride];
    if (c.j > 0) { const long long q = penYm(g, c); a += M.up[1][q] * x[q]; }
    if (c.i > 0) a += M.up[0][p - 32] * x[p - 32];
    if (c.i < g.nx - 1) a += M.up[0][p] * x[p + 32];
    if (c.j < g.ny - 1) a += M.up[1][p] * x[penYp(g, c)];
    if (c.k < g.nz - 1) a += M.up[2][p] * x[p + g.zStride];
    return a;
}

// wA = A pA; wApA = wA.pA; alpha = wArA/wApA (with the singularity test of PCG.C)
__global__ void __launch_bounds__(BLK)
k_pen_amul(PencilGeom g, PenMatrix M, const double* __restrict__ pA, double* __restrict__ wA, FvRed red, FvSolveDev* st)
{
    if (st->done) return;
    if (red.peer) peerHaloWait(red.peer);                  // the neighbours' boundary cells of pA have arrived
    double v[1] = {0.0};
    PEN_ROW_LOOP(g, c) {
        const double a = penAmulCellSym(g, M, pA, c);
        wA[c.pos] = a;
        v[0] += a * pA[c.pos];
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) { penFinWApA(st, t); });
}

// The same product with R consecutive rows of a slab per warp and trip: the rows share their x-neighbours (R+2 loads of
// pA and R+1 of up[0] instead of 3R and 2R), the row decode is paid once, and ~12 R loads are in flight per warp.
// A missing x-neighbour is a pad (value 0) behind a zero coefficient, so the x terms need no predicate (an exact "+ 0");
// the y / z predicates only keep the addresses inside the arrays.  Same terms in the same order as penAmulCellSym.
template <int R>
__global__ void __launch_bounds__(BLK)
k_pen_amul_rows(PencilGeom g, PenMatrix M, const double* __restrict__ pA, double* __restrict__ wA, FvRed red, FvSolveDev* st)
{
    if (st->done) return;
    if (red.peer) peerHaloWait(red.peer);                  // the neighbours' boundary cells of pA have arrived
    double v[1] = {0.0};
    const int lane = threadIdx.x & 31;
    const int nGroups = (int)(g.nLoc / R);                 // Tp is a multiple of 32: a group never straddles two slabs
    const long long dYm = lane > 0 ? -33 : -(long long)g.Tp * 32 + 31 * 32 + 31;
    const long long dYp = lane < 31 ? 33 : (long long)g.Tp * 32 - 31 * 32 - 31;
    const double* __restrict__ dg = M.dg;
    const double* __restrict__ u0 = M.up[0];
    const double* __restrict__ u1 = M.up[1];
    const double* __restrict__ u2 = M.up[2];
    for (int grp = blockIdx.x * (BLK / 32) + (threadIdx.x >> 5); grp < nGroups; grp += gridDim.x * (BLK / 32)) {
        const int row0 = (int)penGlobalRow(g, (long long)grp * R);
        const int sb = row0 / g.Tp, m0 = row0 - sb * g.Tp;
        const int k = sb / g.nJB, jb = sb - k * g.nJB;
        const int j = jb * 32 + lane;
        if (j >= g.ny) continue;
        const bool hasYm = j > 0, hasYp = j < g.ny - 1, hasZm = k > 0, hasZp = k < g.nz - 1;
        const long long p0 = (long long)row0 * 32 + lane;
        double xr[R + 2], cr[R + 1];
#pragma unroll
        for (int r = 0; r < R + 2; ++r) xr[r] = pA[p0 + (r - 1) * 32];
#pragma unroll
        for (int r = 0; r < R + 1; ++r) cr[r] = u0[p0 + (r - 1) * 32];
        double a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long p = p0 + r * 32;
            double t = dg[p] * xr[r + 1];
            if (hasZm) t += u2[p - g.zStride] * pA[p - g.zStride];
            if (hasYm) t += u1[p + dYm] * pA[p + dYm];
            t += cr[r] * xr[r];
            t += cr[r + 1] * xr[r + 2];
            if (hasYp) t += u1[p] * pA[p + dYp];
            if (hasZp) t += u2[p] * pA[p + g.zStride];
            a[r] = t;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = m0 + r - lane;
            if (i >= 0 && i < g.nx) {
                wA[p0 + r * 32] = a[r];
                v[0] += a[r] * xr[r + 1];
            }
        }
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) { penFinWApA(st, t); });
}

// psi += alpha pA; rA -= alpha wA; finalResidual = sum|rA|/normFactor; PCG.C's loop condition.
// A pure vector update: it runs over the arrays as flat streams of double2 (the pads of every vector are zero and stay
// zero, so they add nothing to the sum), with no cell decoding at all.
__global__ void __launch_bounds__(BLK)
k_pen_update(PencilGeom g, const double2* __restrict__ pA, const double2* __restrict__ wA, double2* __restrict__ psi,
             double2* __restrict__ rA, FvRed red, FvSolveDev* st)
{
    if (st->done) return;
    const double alpha = st->alpha;
    double v[1] = {0.0};
    // (the local rows are ONE contiguous range here: every j-block of the planes [kLo, kHi); a y-decomposed solve uses
    // k_pen_update_rows)
    const long long q0 = (long long)g.kLo * g.nJB * g.Tp * 16, n2 = (long long)g.kHi * g.nJB * g.Tp * 16, stride = (long long)gridDim.x * BLK;
#pragma unroll 2
    for (long long q = q0 + (long long)blockIdx.x * BLK + threadIdx.x; q < n2; q += stride) {
        const double2 p = pA[q], w = wA[q];
        double2 x = psi[q], r = rA[q];
        x.x += alpha * p.x;
        x.y += alpha * p.y;
        r.x = r.x - alpha * w.x;
        r.y = r.y - alpha * w.y;
        psi[q] = x;
        rA[q] = r;
        v[0] += fabs(r.x);
        v[0] += fabs(r.y);
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) { penFinUpdate(st, t); });
}

// the same update over an arbitrary set of local rows (y-decomposed solve: the rows of a rank are not contiguous)
__global__ void __launch_bounds__(BLK)
k_pen_update_rows(PencilGeom g, const double* __restrict__ pA, const double* __restrict__ wA, double* __restrict__ psi,
                  double* __restrict__ rA, FvRed red, FvSolveDev* st)
{
    if (st->done) return;
    const double alpha = st->alpha;
    double v[1] = {0.0};
    PEN_ROW_LOOP(g, c) {
        const long long p = c.pos;
        psi[p] += alpha * pA[p];
        const double r = rA[p] - alpha * wA[p];
        rA[p] = r;
        v[0] += fabs(r);
    }
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) { penFinUpdate(st, t); });
}

// ---------------------------------------------------------------------------------------------
// The tail of a PCG iteration as ONE kernel: direction (pA = zA + beta pA), Amul (wA = A pA, wA.pA -> alpha) and update
// (psi += alpha pA, rA -= alpha wA, sum|rA| -> convergence test).  The three steps are separated by two global sums, so
// as three launches each pays a launch, a drain and a reduction tail for a few microseconds of memory traffic -- and on
// a decomposed solve, where a rank's part of the arrays shrinks with the number of ranks, that fixed cost is all that
// is left of them.  Here the grid is launched cooperatively (every block resident) and the steps are separated by grid
// barriers on a device counter:
//   step 1 | barrier: every block's pA is written (decomposed: boundary cells also into the neighbours' ghost cells, block
//          |   0 then raises the neighbours' halo flags, and every block waits for its own rank's flags)
//   step 2 | block partial sums, arrive; block 0 waits for all, adds the partials in block order (decomposed: all-reduce
//          |   over the ranks, fv_peer.cuh), computes alpha, releases the others
//   step 3 | the usual last-block reduction (decomposed: + all-reduce) and the loop test; the last block re-arms the barrier
// A warp works on groups of R consecutive rows, the same groups in every step.
// Loads of values another block wrote earlier in this launch bypass L1 (__ldcg / volatile).
// ---------------------------------------------------------------------------------------------
struct PenTailCtl {
    unsigned int* bar;         // [0] arrivals after step 1  [1] arrivals after step 2  [2] release (alpha is ready)
    int* error;
};

__device__ __forceinline__ void penGridArrive(unsigned int* c)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(c, 1u);
    }
}
__device__ __forceinline__ void penSpinUntil(const unsigned int* c, unsigned int n, int* error)     // one thread
{
    unsigned int spin = 0;
    while (*(const volatile unsigned int*)c < n)
        if (++spin > FY_PEER_SPIN_LIMIT) { atomicExch(error, 1); break; }
    __threadfence();
}

template <int R>
__global__ void __launch_bounds__(BLK, 3)
k_pen_tail(PencilGeom g, PenMatrix M, double* __restrict__ zA, double* pA, double* wA, double* __restrict__ psi,
           double* __restrict__ rA, FvRed red, PenTailCtl tc, FvSolveDev* st, PeerDev* peer)
{
    if (st->done) return;
    __shared__ double shSum[BLK / 32];
    const bool first = st->nIter == 0;
    const double beta = st->beta;
    const double sv = sentValue();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nGroups = (int)(g.nLoc / R);                 // Tp is a multiple of 32: a group never straddles two slabs
    const int w0 = blockIdx.x * (BLK / 32) + warp, nW = gridDim.x * (BLK / 32);
    const unsigned long long haloTarget = peer ? peer->haloSeq + 1 : 0ull;
    // ---- step 1: the search direction
    {
        double* const zlo = peer ? peer->pa[0] : nullptr;
        double* const zhi = peer ? peer->pa[1] : nullptr;
        double* const ylo = peer ? peer->pa[2] : nullptr;
        double* const yhi = peer ? peer->pa[3] : nullptr;
        const int jLo = g.jbLo * 32, jHi = g.jbHi * 32 - 1;
        bool wrote = false;
        for (int grp = w0; grp < nGroups; grp += nW) {
            const int row0 = (int)penGlobalRow(g, (long long)grp * R);
            const int sb = row0 / g.Tp, m0 = row0 - sb * g.Tp;
            const int k = sb / g.nJB, jb = sb - k * g.nJB;
            const int j = jb * 32 + lane;
            if (j >= g.ny) continue;
            const long long p0 = (long long)row0 * 32 + lane;
            double z[R], o[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = m0 + r - lane;
                const bool act = i >= 0 && i < g.nx;
                z[r] = act ? zA[p0 + r * 32] : 0.0;
                o[r] = (act && !first) ? pA[p0 + r * 32] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = m0 + r - lane;
                if (i < 0 || i >= g.nx) continue;
                const long long p = p0 + r * 32;
                const double v = first ? z[r] : z[r] + beta * o[r];
                pA[p] = v;
                zA[p] = sv;
                if (zlo && k == g.kLo) { zlo[p] = v; wrote = true; }
                if (zhi && k == g.kHi - 1) { zhi[p] = v; wrote = true; }
                if (ylo && j == jLo) { ylo[p] = v; wrote = true; }
                if (yhi && j == jHi) { yhi[p] = v; wrote = true; }
            }
        }
        if (wrote) __threadfence_system();                 // remote stores performed before this block reports
    }
    penGridArrive(tc.bar + 0);
    if (threadIdx.x == 0) {
        penSpinUntil(tc.bar + 0, gridDim.x, tc.error);
        if (peer) {
            if (blockIdx.x == 0) {
                __threadfence_system();
                for (int d = 0; d < 4; ++d)
                    if (peer->nbr[d] >= 0) peerSt(&peer->box[peer->nbr[d]]->halo[d ^ 1], haloTarget);
                peer->haloSeq = haloTarget;
            }
            const unsigned long long* f = peer->box[peer->rank]->halo;
            for (int d = 0; d < 4; ++d) {
                if (peer->nbr[d] < 0) continue;
                unsigned int spin = 0;
                while (peerLd(f + d) < haloTarget)
                    if (++spin > FY_PEER_SPIN_LIMIT) { atomicExch(tc.error, 1); break; }
            }
            __threadfence();
        }
    }
    __syncthreads();
    // ---- step 2: wA = A pA, wA.pA
    double acc = 0.0;
    {
        const long long dYm = lane > 0 ? -33 : -(long long)g.Tp * 32 + 31 * 32 + 31;
        const long long dYp = lane < 31 ? 33 : (long long)g.Tp * 32 - 31 * 32 - 31;
        const double* __restrict__ dg = M.dg;
        const double* __restrict__ u0 = M.up[0];
        const double* __restrict__ u1 = M.up[1];
        const double* __restrict__ u2 = M.up[2];
        for (int grp = w0; grp < nGroups; grp += nW) {
            const int row0 = (int)penGlobalRow(g, (long long)grp * R);
            const int sb = row0 / g.Tp, m0 = row0 - sb * g.Tp;
            const int k = sb / g.nJB, jb = sb - k * g.nJB;
            const int j = jb * 32 + lane;
            if (j >= g.ny) continue;
            const bool hasYm = j > 0, hasYp = j < g.ny - 1, hasZm = k > 0, hasZp = k < g.nz - 1;
            const long long p0 = (long long)row0 * 32 + lane;
            double xr[R + 2], cr[R + 1];
#pragma unroll
            for (int r = 0; r < R + 2; ++r) xr[r] = __ldcg(pA + p0 + (r - 1) * 32);
#pragma unroll
            for (int r = 0; r < R + 1; ++r) cr[r] = u0[p0 + (r - 1) * 32];
            double a[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long p = p0 + r * 32;
                double t = dg[p] * xr[r + 1];
                if (hasZm) t += u2[p - g.zStride] * __ldcg(pA + p - g.zStride);
                if (hasYm) t += u1[p + dYm] * __ldcg(pA + p + dYm);
                t += cr[r] * xr[r];
                t += cr[r + 1] * xr[r + 2];
                if (hasYp) t += u1[p] * __ldcg(pA + p + dYp);
                if (hasZp) t += u2[p] * __ldcg(pA + p + g.zStride);
                a[r] = t;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = m0 + r - lane;
                if (i >= 0 && i < g.nx) {
                    wA[p0 + r * 32] = a[r];
                    acc += a[r] * xr[r + 1];
                }
            }
        }
    }
    {   // block sum -> partial[block]; block 0 finishes the sum for everybody
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) shSum[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double x = shSum[0];
            for (int w = 1; w < BLK / 32; ++w) x += shSum[w];
            red.partial[blockIdx.x] = x;
        }
        penGridArrive(tc.bar + 1);
        if (blockIdx.x == 0) {
            if (threadIdx.x == 0) penSpinUntil(tc.bar + 1, gridDim.x, tc.error);
            __syncthreads();
            const volatile double* p = red.partial;
            double x = 0.0;
            for (unsigned int b = threadIdx.x; b < gridDim.x; b += BLK) x += p[b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            __syncthreads();
            if (lane == 0) shSum[warp] = x;
            __syncthreads();
            double tot[1];
            tot[0] = shSum[0];
            for (int w = 1; w < BLK / 32; ++w) tot[0] += shSum[w];
            if (peer && threadIdx.x < 32) peerAllReduce<1>(peer, tot);
            if (threadIdx.x == 0) {
                penFinWApA(st, tot);
                __threadfence();
                *(volatile unsigned int*)(tc.bar + 2) = 1u;
            }
        }
        if (threadIdx.x == 0) penSpinUntil(tc.bar + 2, 1u, tc.error);
        __syncthreads();
    }
    // ---- step 3: the update and the loop test
    const double alpha = *(const volatile double*)&st->alpha;
    double v[1] = {0.0};
    for (int grp = w0; grp < nGroups; grp += nW) {
        const int row0 = (int)penGlobalRow(g, (long long)grp * R);
        const int sb = row0 / g.Tp, m0 = row0 - sb * g.Tp;
        const int jb = sb % g.nJB;
        const int j = jb * 32 + lane;
        if (j >= g.ny) continue;
        const long long p0 = (long long)row0 * 32 + lane;
        double pp[R], ww[R], xx[R], rr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = m0 + r - lane;
            const bool act = i >= 0 && i < g.nx;
            const long long p = p0 + r * 32;
            pp[r] = act ? pA[p] : 0.0;
            ww[r] = act ? wA[p] : 0.0;
            xx[r] = act ? psi[p] : 0.0;
            rr[r] = act ? rA[p] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = m0 + r - lane;
            if (i < 0 || i >= g.nx) continue;
            const long long p = p0 + r * 32;
            psi[p] = xx[r] + alpha * pp[r];
            const double q = rr[r] - alpha * ww[r];
            rA[p] = q;
            v[0] += fabs(q);
        }
    }
    unsigned int* const bar = tc.bar;
    fvGridReduce<1, false, BLK>(v, red, [=](const double* t) {
        penFinUpdate(st, t);
        bar[0] = 0u;                                       // every block is past the barriers: re-arm them
        bar[1] = 0u;
        bar[2] = 0u;
    });
}

// y-decomposed solve: the edge lanes of a rank's j-block range, packed for the halo exchange of the search direction.
// buf[0 .. n) = lane 0 of the first j-block (goes to the rank below in y), buf[n .. 2n) = lane 31 of the last one (goes up);
// n = (kHi-kLo)*Tp.  Unpacking writes the received values into the neighbours' edge lanes next to the range (the ghosts).
__global__ void __launch_bounds__(BLK) k_pen_pack_yedge(PencilGeom g, const double* __restrict__ v, double* __restrict__ buf)
{
    const long long n = (long long)(g.kHi - g.kLo) * g.Tp;
    for (long long q = (long long)blockIdx.x * BLK + threadIdx.x; q < 2 * n; q += (long long)gridDim.x * BLK) {
        const bool hi = q >= n;
        const long long e = hi ? q - n : q;
        const int kk = (int)(e / g.Tp), m = (int)(e - (long long)kk * g.Tp);
        const int jb = hi ? g.jbHi - 1 : g.jbLo;
        buf[q] = v[(((long long)(g.kLo + kk) * g.nJB + jb) * g.Tp + m) * 32 + (hi ? 31 : 0)];
    }
}
__global__ void __launch_bounds__(BLK) k_pen_unpack_yedge(PencilGeom g, const double* __restrict__ buf, double* __restrict__ v)
{
    // buf[0 .. n) = received from below: lane 31 of j-block jbLo-1 ; buf[n .. 2n) = received from above: lane 0 of j-block jbHi
    const long long n = (long long)(g.kHi - g.kLo) * g.Tp;
    for (long long q = (long long)blockIdx.x * BLK + threadIdx.x; q < 2 * n; q += (long long)gridDim.x * BLK) {
        const bool hi = q >= n;
        if ((hi && g.jbHi >= g.nJB) || (!hi && g.jbLo <= 0)) continue;
        const long long e = hi ? q - n : q;
        const int kk = (int)(e / g.Tp), m = (int)(e - (long long)kk * g.Tp);
        const int jb = hi ? g.jbHi : g.jbLo - 1;
        v[(((long long)(g.kLo + kk) * g.nJB + jb) * g.Tp + m) * 32 + (hi ? 0 : 31)] = buf[q];
    }
}
// the owned region (planes x j-blocks) of a vector as one contiguous buffer, and back from the buffers of all ranks
__global__ void __launch_bounds__(BLK) k_pen_pack_region(PencilGeom g, const double* __restrict__ v, double* __restrict__ buf)
{
    for (long long l_ = (long long)blockIdx.x * (BLK / 32) + (threadIdx.x >> 5); l_ < g.nLoc; l_ += (long long)gridDim.x * (BLK / 32))
        buf[l_ * 32 + (threadIdx.x & 31)] = v[penGlobalRow(g, l_) * 32 + (threadIdx.x & 31)];
}
__global__ void __launch_bounds__(BLK) k_pen_unpack_region(PencilGeom g, const double* __restrict__ buf, double* __restrict__ v)
{
    for (long long l_ = (long long)blockIdx.x * (BLK / 32) + (threadIdx.x >> 5); l_ < g.nLoc; l_ += (long long)gridDim.x * (BLK / 32))
        v[penGlobalRow(g, l_) * 32 + (threadIdx.x & 31)] = buf[l_ * 32 + (threadIdx.x & 31)];
}
// drops a coefficient array's entries on one lane of one j-block's rows (the y face of a slab), owned planes only
__global__ void __launch_bounds__(BLK) k_pen_zero_lane(PencilGeom g, double* __restrict__ a, int jb, int lane)
{
    const long long n = (long long)(g.kHi - g.kLo) * g.Tp;
    for (long long e = (long long)blockIdx.x * BLK + threadIdx.x; e < n; e += (long long)gridDim.x * BLK) {
        const int kk = (int)(e / g.Tp), m = (int)(e - (long long)kk * g.Tp);
        a[(((long long)(g.kLo + kk) * g.nJB + jb) * g.Tp + m) * 32 + lane] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define PEN_LAUNCH(kernel, ...)                                                             \
    do {                                                                                    \
        kernel<<<s->pen.rowGrid, BLK, 0, h->stream>>>(__VA_ARGS__);                         \
        FY_CHECK_LAUNCH();                                                                  \
    } while (0)

// cudaFuncSetAttribute is per (function, device): remember which pairs have been opted in
int penFuncAttrs(fy_ctx* h, const void* fn)
{
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({fn, h->device})) return FY_OK;
    FY_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    FY_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    done.insert({fn, h->device});
    return FY_OK;
}

template <class Op, bool REV>
int launchPencil(fy_ctx* h, FvState* s, const Op& op, FvSolveDev* st)
{
    PenState& P = s->pen;
    // warps per group: as many as the shared-memory budget allows (each owns D*NIN rows + one channel)
    const int perWarp = penDepth(Op::NIN) * Op::NIN * 256 + PEN_CD * 256 + PEN_CY * 8;
    int W = std::max(1, std::min(std::min(P.W, PEN_WMAX), P.smemBudget / perWarp));
    W = std::min(W, P.g.nz);
    const size_t smem = (size_t)W * perWarp;
    if (int rc = penFuncAttrs(h, (const void*)k_pencil<Op, REV>)) return rc;
    PenCtl ctl{P.ticket, P.error, P.partial, st, P.dbg, P.traceOn ? P.trace : nullptr, nullptr};
    // clusters of C consecutive plane groups hand the z-neighbour over through distributed shared memory
    const int nKQ = (P.g.nz + W - 1) / W;
    int C = 1;
    while (C * 2 <= P.cluster && C < nKQ) C *= 2;
    const int nCl = (nKQ + C - 1) / C;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(P.g.nJB * nCl * C));
    cfg.blockDim = dim3(32 * (2 * W + 1));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (C > 8) {                                           // 16-CTA clusters are non-portable: check that one fits
        int nClusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nClusters, k_pencil<Op, REV>, &cfg) != cudaSuccess || nClusters < 1) {
            cudaGetLastError();
            C = 8;
            attr[0].val.clusterDim.x = 8;
            cfg.gridDim = dim3((unsigned)(P.g.nJB * ((nKQ + 7) / 8) * 8));
        }
    }
    FY_CUDA(cudaLaunchKernelEx(&cfg, k_pencil<Op, REV>, P.g, op, ctl));
    h->launches++;
    return FY_OK;
}

// second-generation sweeps (fv_pencil2.cuh): Z planes per compute warp, R rows per TMA stage
template <class Op, bool REV, int Z, int R>
int launchPen2T(fy_ctx* h, FvState* s, const PencilGeom& geom, const Op& op, FvSolveDev* st, double* distOut, PeerDev* peer)
{
    PenState& P = s->pen;
    const int nzL = geom.kHi - geom.kLo;
    const int stageBytes = Z * Op::NA * R * 256;
    const int fixedPerWarp = P2_CD * 256 + Z * P2_YRING * 8 + 16 * P2_MAXSTAGE + 8 * 8 + 16;
    int W = std::max(1, std::min(P.W2, Pen2Max<Z>::W));
    W = std::min(W, (nzL + Z - 1) / Z);
    int nStage = 0;
    for (; W >= 1; --W) {
        nStage = std::min(std::min(P.maxStage2, P2_MAXSTAGE), (P.smemBudget2 / W - fixedPerWarp) / stageBytes);
        if (nStage >= 2) break;
    }
    if (nStage < 2) {
        W = 1;
        nStage = std::min(P2_MAXSTAGE, (216 * 1024 - fixedPerWarp) / stageBytes);
        if (nStage < 2) { h->err = "pencil sweep: stage does not fit shared memory"; return FY_ERR_INVALID; }
        nStage = std::min(nStage, std::max(2, P.maxStage2));
    }
    const size_t smem = (size_t)W * nStage * stageBytes + (size_t)W * P2_CD * 256 + (size_t)W * Z * P2_YRING * 8 + (size_t)W * nStage * 16 +
                        (size_t)W * 8 * 8 + (size_t)W * 4 + 16;
    if (int rc = penFuncAttrs(h, (const void*)k_pen2<Op, REV, Z, R>)) return rc;
    PenCtl ctl{P.ticket, P.error, P.partial, st, P.dbg, P.traceOn ? P.trace : nullptr, distOut, peer};
    const int PZ = W * Z, nKQ = (nzL + PZ - 1) / PZ;
    int C = 1;
    while (C * 2 <= P.cluster && C < nKQ) C *= 2;
    const int nCl = (nKQ + C - 1) / C;
    cudaLaunchConfig_t cfg = {};
    const int nJl = geom.jbHi - geom.jbLo;
    cfg.gridDim = dim3((unsigned)(nJl * nCl * C));
    cfg.blockDim = dim3(32 * (2 * W + 1 + W * Z));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (C > 8) {                                           // 16-CTA clusters are non-portable: check that one fits
        int nClusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nClusters, k_pen2<Op, REV, Z, R>, &cfg) != cudaSuccess || nClusters < 1) {
            cudaGetLastError();
            C = 8;
            attr[0].val.clusterDim.x = 8;
            cfg.gridDim = dim3((unsigned)(nJl * ((nKQ + 7) / 8) * 8));
        }
    }
    FY_CUDA(cudaLaunchKernelEx(&cfg, k_pen2<Op, REV, Z, R>, geom, op, ctl, W, nStage));
    h->launches++;
    return FY_OK;
}
template <class Op, bool REV>
int launchPen2(fy_ctx* h, FvState* s, const PencilGeom& g, const Op& op, FvSolveDev* st, double* distOut = nullptr, PeerDev* peer = nullptr)
{
    const PenState& P = s->pen;
    const int key = P.Z2 * 100 + P.R2;
    switch (key) {
    case 104: return launchPen2T<Op, REV, 1, 4>(h, s, g, op, st, distOut, peer);
    case 204: return launchPen2T<Op, REV, 2, 4>(h, s, g, op, st, distOut, peer);
#ifdef PEN2_ALL_VARIANTS
    case 108: return launchPen2T<Op, REV, 1, 8>(h, s, g, op, st, distOut, peer);
    case 208: return launchPen2T<Op, REV, 2, 8>(h, s, g, op, st, distOut, peer);
    case 408: return launchPen2T<Op, REV, 4, 8>(h, s, g, op, st, distOut, peer);
#endif
    case 404: return launchPen2T<Op, REV, 4, 4>(h, s, g, op, st, distOut, peer);
    default: h->err = "pencil sweep: no kernel for this FY_PEN2_Z / FY_PEN2_R"; return FY_ERR_INVALID;
    }
}

int readSolve(fy_ctx* h, FvState* s)
{
    FY_CUDA(cudaMemcpyAsync(s->hSolve, s->dSolve, sizeof(FvSolveDev), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaMemcpyAsync(s->pen.hError, s->pen.error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    if (*s->pen.hError) {
        h->err = "pencil pipeline: a neighbour value never arrived (poll limit reached)";
        return FY_ERR_CUDA;
    }
    return FY_OK;
}

PenMatrix penMatrixOf(PenState& P, int which)      // 0: p (symmetric: low built from upper)  1: U
{
    PenMatrix M;
    double** c = which == 0 ? P.mP : P.mU;
    M.dg = c[0];
    for (int d = 0; d < 3; ++d) { M.low[d] = c[1 + d]; M.up[d] = c[4 + d]; }
    return M;
}
}  // namespace

int penPackYEdge(fy_ctx* h, FvState* s, const PencilGeom& g, const double* v, double* buf)
{
    const long long n = 2LL * (g.kHi - g.kLo) * g.Tp;
    k_pen_pack_yedge<<<(unsigned)std::min<long long>((n + BLK - 1) / BLK, 4096), BLK, 0, h->stream>>>(g, v, buf);
    FY_CHECK_LAUNCH();
    return FY_OK;
}
int penUnpackYEdge(fy_ctx* h, FvState* s, const PencilGeom& g, const double* buf, double* v)
{
    const long long n = 2LL * (g.kHi - g.kLo) * g.Tp;
    k_pen_unpack_yedge<<<(unsigned)std::min<long long>((n + BLK - 1) / BLK, 4096), BLK, 0, h->stream>>>(g, buf, v);
    FY_CHECK_LAUNCH();
    return FY_OK;
}
int penPackRegion(fy_ctx* h, FvState* s, const PencilGeom& g, const double* v, double* buf)
{
    PEN_LAUNCH(k_pen_pack_region, g, v, buf);
    return FY_OK;
}
int penUnpackRegion(fy_ctx* h, FvState* s, const PencilGeom& g, const double* buf, double* v)
{
    PEN_LAUNCH(k_pen_unpack_region, g, buf, v);
    return FY_OK;
}

int penCreate(fy_ctx* h, FvState* s)
{
    PenState& P = s->pen;
    const BoxGeom& b = s->g;
    PencilGeom& g = P.g;
    g.nx = b.nx; g.ny = b.ny; g.nz = b.nz; g.N = b.N;
    g.nJB = (b.ny + 31) / 32;
    g.Tp = ((b.nx + 31 + PEN_CY - 1) / PEN_CY) * PEN_CY;
    g.nRows = (long long)b.nz * g.nJB * g.Tp;
    if (g.nRows * 32 >= (1LL << 31)) { h->err = "pencil layout: mesh too large for 32-bit row arithmetic"; return FY_ERR_INVALID; }
    g.NP = g.nRows * 32;
    g.zStride = (long long)g.nJB * g.Tp * 32;
    g.kLo = 0; g.kHi = g.nz; g.jbLo = 0; g.jbHi = g.nJB; g.nLoc = g.nRows;
    P.gl = g; P.kLo = 0; P.kHi = g.nz;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int rowMult = 8;
    if (const char* e = std::getenv("FY_ROWGRID_MULT")) { const int m = std::atoi(e); if (m >= 1 && m <= 128) rowMult = m; }
    P.rowGrid = (int)std::max<long long>(1, std::min<long long>((g.nRows + BLK / 32 - 1) / (BLK / 32), (long long)sms * rowMult));
    P.W = 8;
    P.cluster = 16;
    if (const char* e = std::getenv("FY_PENCIL_CLUSTER")) { const int c = std::atoi(e); if (c >= 1 && c <= 16) P.cluster = c; }
    if (const char* e = std::getenv("FY_PENCIL_W")) { const int w = std::atoi(e); if (w >= 1 && w <= PEN_WMAX) P.W = w; }
    P.smemBudget = 216 * 1024;
    if (const char* e = std::getenv("FY_PENCIL_SMEM_KB")) { const int k = std::atoi(e); if (k >= 16 && k <= 216) P.smemBudget = k * 1024; }
    // every array carries PEN_GUARD rows of zeros in front and behind: the sweeps prefetch D rows past either end
    const size_t guard = (size_t)PEN_GUARD * 32, bytes = ((size_t)g.NP + 2 * guard) * sizeof(double);
    auto alloc = [&](double*& p, int mult = 1) -> int {
        double* raw = nullptr;
        FY_CUDA(cudaMalloc((void**)&raw, bytes * mult));
        FY_CUDA(cudaMemsetAsync(raw, 0, bytes * mult, h->stream));
        p = raw + guard * mult;
        return FY_OK;
    };
    int rc;
    if (const char* e = std::getenv("FY_PENCIL_VER")) P.ver = std::atoi(e) == 1 ? 1 : 2;
    if (const char* e = std::getenv("FY_PEN2_Z")) P.Z2 = std::atoi(e);
    if (const char* e = std::getenv("FY_PEN2_R")) P.R2 = std::atoi(e);
    if (const char* e = std::getenv("FY_PEN2_W")) P.W2 = std::atoi(e);
    if (const char* e = std::getenv("FY_PEN2_STAGES")) P.maxStage2 = std::max(2, std::atoi(e));
    if (const char* e = std::getenv("FY_PEN2_SMEM_KB")) { const int k = std::atoi(e); if (k >= 16 && k <= 216) P.smemBudget2 = k * 1024; }
    for (int q = 0; q < 4; ++q) if ((rc = alloc(P.pk[q], q < 3 ? 2 : 1))) return rc;
    for (auto& p : P.mP) if ((rc = alloc(p))) return rc;
    for (auto& p : P.mU) if ((rc = alloc(p))) return rc;
    for (auto& p : P.v) if ((rc = alloc(p))) return rc;
    const int maxWarps = g.nJB * (g.nz + 16 * PEN_WMAX + 64) * 4 + 64;
    FY_CUDA(cudaMalloc((void**)&P.partial, (size_t)maxWarps * sizeof(double)));
    FY_CUDA(cudaMalloc((void**)&P.trace, (size_t)maxWarps * 32 * sizeof(unsigned long long)));
    FY_CUDA(cudaMemsetAsync(P.trace, 0, (size_t)maxWarps * 32 * sizeof(unsigned long long), h->stream));
    P.traceOn = std::getenv("FY_PENCIL_TRACE") != nullptr;
    if (const char* e = std::getenv("FY_PCG_GRAPH")) P.useGraphs = std::atoi(e) != 0;
    if (const char* e = std::getenv("FY_PENCIL_DBG")) P.dbg = std::atoi(e);
    FY_CUDA(cudaMalloc((void**)&P.tailBar, 4 * sizeof(unsigned int)));
    FY_CUDA(cudaMemsetAsync(P.tailBar, 0, 4 * sizeof(unsigned int), h->stream));
    if (const char* e = std::getenv("FY_PCG_FUSED")) P.fusedTail = std::atoi(e) != 0;
    FY_CUDA(cudaMalloc((void**)&P.ticket, 2 * sizeof(unsigned int)));
    FY_CUDA(cudaMemsetAsync(P.ticket, 0, 2 * sizeof(unsigned int), h->stream));
    FY_CUDA(cudaMalloc((void**)&P.error, sizeof(int)));
    FY_CUDA(cudaMemsetAsync(P.error, 0, sizeof(int), h->stream));
    FY_CUDA(cudaHostAlloc((void**)&P.hError, sizeof(int), cudaHostAllocDefault));
    *P.hError = 0;
    return FY_OK;
}

void penDestroy(FvState* s)
{
    PenState& P = s->pen;
    fvDistDestroy(s);
    const size_t guard = (size_t)PEN_GUARD * 32;
    for (int q = 0; q < 4; ++q) if (P.pk[q]) cudaFree(P.pk[q] - guard * (q < 3 ? 2 : 1));
    for (auto p : P.mP) if (p) cudaFree(p - guard);
    for (auto p : P.mU) if (p) cudaFree(p - guard);
    for (auto p : P.v) if (p) cudaFree(p - guard);
    for (auto& ge : P.pcgGraph) if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
    if (P.partial) cudaFree(P.partial);
    if (P.trace) cudaFree(P.trace);
    if (P.tailBar) cudaFree(P.tailBar);
    P.tailBar = nullptr;
    if (P.ticket) cudaFree(P.ticket);
    if (P.error) cudaFree(P.error);
    if (P.hError) cudaFreeHost(P.hError);
}

// vector roles inside the shared pool P.v[]
enum { V_B = 0, V_X, V_RD, V_D, V_RA, V_PA, V_WA, V_YA, V_ZA, V_BPRIME = V_RA, V_MID = V_PA, V_EX = V_WA, V_EY = V_YA, V_EZ = V_ZA };

double* penSearchDir(PenState& P, size_t* guardElems)
{
    *guardElems = (size_t)PEN_GUARD * 32;
    return P.v[V_PA];
}

// the five recurrences, dispatched to the pipeline generation in use
static int sweepDicD(fy_ctx* h, FvState* s, const PencilGeom& g, const PenMatrix& M, FvSolveDev* st)
{
    PenState& P = s->pen;
    double** v = P.v;
    int rc;
    if (P.ver == 1) {
        OpDicD op{{M.dg, M.low[0], M.low[1], M.low[2]}, v[V_D], v[V_RD]};
        return launchPencil<OpDicD, false>(h, s, op, st);
    }
    Op2DicD op{{M.dg, M.low[0], M.low[1], M.low[2]}, v[V_D], v[V_RD]};
    if ((rc = launchPen2<Op2DicD, false>(h, s, g, op, st))) return rc;
    PEN_LAUNCH(k_pen_pack_dic, g, v[V_RD], M, (double2*)P.pk[0], (double2*)P.pk[1], (double2*)P.pk[2], P.pk[3]);
    return FY_OK;
}
static int sweepDicFwd(fy_ctx* h, FvState* s, const PencilGeom& g, const PenMatrix& M, FvSolveDev* st)
{
    PenState& P = s->pen;
    double** v = P.v;
    if (P.ver == 1) {
        OpDicFwd f{{v[V_RD], v[V_RA], M.low[0], M.low[1], M.low[2]}, v[V_YA]};
        return launchPencil<OpDicFwd, false>(h, s, f, st);
    }
    Op2DicFwd f{{P.pk[0], P.pk[1], v[V_RA]}, v[V_YA]};
    return launchPen2<Op2DicFwd, false>(h, s, g, f, st);
}
static int sweepDicBwd(fy_ctx* h, FvState* s, const PencilGeom& g, const PenMatrix& M, FvSolveDev* st, double* distOut = nullptr,
                       PeerDev* peer = nullptr)
{
    PenState& P = s->pen;
    double** v = P.v;
    if (P.ver == 1) {
        OpDicBwd bw{{v[V_YA], v[V_RD], M.up[0], M.up[1], M.up[2], v[V_RA]}, v[V_ZA], v[V_YA]};
        return launchPencil<OpDicBwd, true>(h, s, bw, st);
    }
    Op2DicBwd bw{{v[V_YA], P.pk[2], P.pk[3], v[V_RA]}, v[V_ZA], v[V_YA]};
    return launchPen2<Op2DicBwd, true>(h, s, g, bw, st, distOut, peer);
}
static int sweepGsFwd(fy_ctx* h, FvState* s, const PenMatrix& M, FvSolveDev* st)
{
    PenState& P = s->pen;
    double** v = P.v;
    if (P.ver == 1) {
        OpGsFwd f{{v[V_B], M.low[0], M.low[1], M.low[2], v[V_EX], v[V_EY], v[V_EZ], M.dg}, v[V_MID], v[V_BPRIME], v[V_X]};
        return launchPencil<OpGsFwd, false>(h, s, f, st);
    }
    Op2GsFwd f{{v[V_B], M.low[0], M.low[1], M.low[2], v[V_EX], v[V_EY], v[V_EZ], M.dg}, v[V_MID], v[V_BPRIME], v[V_X]};
    return launchPen2<Op2GsFwd, false>(h, s, P.g, f, st);
}
static int sweepGsBwd(fy_ctx* h, FvState* s, const PenMatrix& M, FvSolveDev* st)
{
    PenState& P = s->pen;
    double** v = P.v;
    if (P.ver == 1) {
        OpGsBwd bw{{v[V_BPRIME], M.up[0], M.up[1], M.up[2], M.dg}, v[V_X], v[V_MID]};
        return launchPencil<OpGsBwd, true>(h, s, bw, st);
    }
    Op2GsBwd bw{{v[V_BPRIME], M.up[0], M.up[1], M.up[2], M.dg}, v[V_X], v[V_MID]};
    return launchPen2<Op2GsBwd, true>(h, s, P.g, bw, st);
}

static int launchTail(fy_ctx* h, FvState* s, const PencilGeom& gl, const PenMatrix& M, const FvRed& red, PeerDev* peer)
{
    PenState& P = s->pen;
    double** v = P.v;
    if (P.tailBlocksPerSm == 0) {
        int nb = 0;
        FY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pen_tail<PEN_AMUL_R>, BLK, 0));
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        P.tailBlocksPerSm = std::max(1, nb);
        P.tailMaxGrid = sms * P.tailBlocksPerSm;
    }
    const long long nGroups = gl.nLoc / PEN_AMUL_R;
    const int grid = (int)std::max<long long>(1, std::min<long long>((nGroups + BLK / 32 - 1) / (BLK / 32), P.tailMaxGrid));
    PenTailCtl tc{P.tailBar, P.error};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(BLK);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;           // every block resident: the kernel holds grid barriers
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FY_CUDA(cudaLaunchKernelEx(&cfg, k_pen_tail<PEN_AMUL_R>, gl, M, v[V_ZA], v[V_PA], v[V_WA], v[V_X], v[V_RA], red, tc, s->dSolve, peer));
    h->launches++;
    return FY_OK;
}

// PCG on owner-slot coefficients (device pointers, natural cell order).  Iteration kernels are queued in
// batches and test the device-side `done` flag themselves; the host looks at the state once per batch.
int fvPcgSolve(fy_ctx* h, FvState* s, const double* dg, const double* up, const double* b, double* psi, double tol,
               double relTol, int maxIter, int precond, fy_solver_perf* perf, bool sameMatrix)
{
    PenState& P = s->pen;
    const PencilGeom& g = P.g;
    int rc;
    PenMatrix M = penMatrixOf(P, 0);
    double** v = P.v;
    // sameMatrix: the caller solved with these very coefficients last time (the PISO correctors of one time step share
    // their pEqn matrix, only the source changes) -- the pencil-layout copy and the preconditioner's reciprocal
    // diagonal are still in place.  OpenFOAM recomputes them for every solve; the values are the same.
    const bool reuse = sameMatrix && P.precondOf == precond;
    P.precondOf = -1;
    // decomposed solve (fv_dist.cu): the iteration works on this rank's rows `gl`; a kernel's sum is this rank's part,
    // all-reduced over the ranks before the finishing kernel runs what the single-domain kernel does in its last block
    const bool dist = P.dist;
    const PencilGeom& gl = dist ? P.gl : P.g;
    // the collectives of the iteration: inside its kernels over peer memory (fv_peer.cuh), else as NCCL calls between them
    PeerDev* const peer = dist ? P.peer : nullptr;
    FvRed redL = s->red;
    if (dist && !peer) redL.distOut = P.distBuf;
    redL.peer = peer;
    auto finish = [&](int which, int nv) -> int {
        if (!dist || peer) return FY_OK;
        if ((rc = fvDistAllReduce(h, s, P.distBuf, nv))) return rc;
        k_pen_fin<<<1, 1, 0, h->stream>>>(which, P.distBuf, s->dSolve, g.N);
        FY_CHECK_LAUNCH();
        return FY_OK;
    };
    if (!reuse) {
        PEN_LAUNCH(k_pen_matrix, g, dg, up, up, M);
        // the preconditioner of a slab sees the slab's own matrix: no coupling through its bottom face (M.low serves
        // the DIC recurrences only; Amul and the residual take every coefficient from M.up)
        if (dist && P.kLo > 0)
            FY_CUDA(cudaMemsetAsync(M.low[2] + (size_t)P.kLo * g.zStride, 0, (size_t)g.zStride * sizeof(double), h->stream));
        if (dist && gl.jbLo > 0) {
            k_pen_zero_lane<<<std::max(1, (int)(((long long)(gl.kHi - gl.kLo) * g.Tp + BLK - 1) / BLK)), BLK, 0, h->stream>>>(gl, M.low[1], gl.jbLo, 0);
            FY_CHECK_LAUNCH();
        }
    }
    PEN_LAUNCH(k_pen_from_nat, g, b, v[V_B]);
    PEN_LAUNCH(k_pen_from_nat, g, psi, v[V_X]);
    k_solve_begin<<<1, 1, 0, h->stream>>>(s->dSolve, tol, relTol, maxIter, precond);
    FY_CHECK_LAUNCH();
    PEN_LAUNCH(k_pen_avg, gl, v[V_X], redL, s->dSolve);
    if ((rc = finish(PEN_FIN_AVG, 1))) return rc;
    PEN_LAUNCH(k_pen_solve_init<true>, gl, M, v[V_B], v[V_X], v[V_RA], redL, s->dSolve);
    if ((rc = finish(PEN_FIN_INIT, 2))) return rc;
    if (precond == FV_PRECOND_DIC) {
        PEN_LAUNCH(k_pen_arm, g, reuse ? (double*)nullptr : v[V_D], v[V_YA], v[V_ZA], s->dSolve);
        if (!reuse && (rc = sweepDicD(h, s, gl, M, s->dSolve))) return rc;
    } else if (precond == FV_PRECOND_DIAGONAL && !reuse) {
        PEN_LAUNCH(k_pen_recip, g, M.dg, v[V_RD]);
    }
    bool sampled = false;
    const bool prof = h->profiling && s->pev[0];
    // direction + Amul + update as one cooperative kernel (k_pen_tail); the NCCL path needs host calls between them
    const bool fusedTail = P.fusedTail && (!dist || peer);
    // one PCG iteration = 5 launches whose arguments never change (the vector pool and the matrix arrays are fixed
    // for the engine's life): a batch of iterations is captured once into a CUDA graph and replayed
    auto enqueueIteration = [&](bool ev) -> int {
        if (ev) cudaEventRecord(s->pev[0], h->stream);
        if (precond == FV_PRECOND_DIC) {
            if ((rc = sweepDicFwd(h, s, gl, M, s->dSolve))) return rc;
            if (ev) cudaEventRecord(s->pev[1], h->stream);
            if ((rc = sweepDicBwd(h, s, gl, M, s->dSolve, dist && !peer ? P.distBuf : nullptr, peer))) return rc;
        } else {
            if (ev) cudaEventRecord(s->pev[1], h->stream);
            PEN_LAUNCH(k_pen_precond_diag, gl, precond == FV_PRECOND_DIAGONAL ? v[V_RD] : (const double*)nullptr, v[V_RA],
                       v[V_ZA], redL, s->dSolve);
        }
        if ((rc = finish(PEN_FIN_WARA, 1))) return rc;
        if (ev) cudaEventRecord(s->pev[2], h->stream);
        if (fusedTail) {
            if ((rc = launchTail(h, s, gl, M, redL, peer))) return rc;
            if (ev) { cudaEventRecord(s->pev[3], h->stream); cudaEventRecord(s->pev[4], h->stream); cudaEventRecord(s->pev[5], h->stream); }
            return FY_OK;
        }
        PEN_LAUNCH(k_pen_dir, gl, v[V_ZA], v[V_PA], s->dSolve, peer);
        if (dist && !peer && (rc = fvDistHalo(h, s, v[V_PA]))) return rc;      // Amul reads the neighbours' boundary planes of pA
        if (ev) cudaEventRecord(s->pev[3], h->stream);
#if PEN_AMUL_R > 1
        PEN_LAUNCH(k_pen_amul_rows<PEN_AMUL_R>, gl, M, v[V_PA], v[V_WA], redL, s->dSolve);
#else
        PEN_LAUNCH(k_pen_amul, gl, M, v[V_PA], v[V_WA], redL, s->dSolve);
#endif
        if ((rc = finish(PEN_FIN_WAPA, 1))) return rc;
        if (ev) cudaEventRecord(s->pev[4], h->stream);
        if (gl.jbHi - gl.jbLo == g.nJB) {
            PEN_LAUNCH(k_pen_update, gl, (const double2*)v[V_PA], (const double2*)v[V_WA], (double2*)v[V_X], (double2*)v[V_RA], redL,
                       s->dSolve);
        } else {
            PEN_LAUNCH(k_pen_update_rows, gl, v[V_PA], v[V_WA], v[V_X], v[V_RA], redL, s->dSolve);
        }
        if ((rc = finish(PEN_FIN_UPDATE, 1))) return rc;
        if (ev) cudaEventRecord(s->pev[5], h->stream);
        return FY_OK;
    };
    const int batch = s->pcgBatch;
    cudaGraphExec_t& gexec = P.pcgGraph[precond];
    const bool useGraph = P.useGraphs && !prof && !P.traceOn && (!dist || peer);      // (NCCL calls inside the iteration: queued eagerly)
    if (useGraph && !gexec && P.graphWarm[precond]) {
        // (every kernel has run eagerly once by now: function attributes set, modules loaded)
        const long long l0 = h->launches;
        cudaGraph_t graph = nullptr;
        FY_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        for (int it = 0; it < batch; ++it)
            if ((rc = enqueueIteration(false))) { cudaStreamEndCapture(h->stream, &graph); if (graph) cudaGraphDestroy(graph); return rc; }
        FY_CUDA(cudaStreamEndCapture(h->stream, &graph));
        FY_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
        cudaGraphDestroy(graph);
        P.graphLaunches = (int)(h->launches - l0);
        h->launches = l0;
    }
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
        if (useGraph && gexec) {
            FY_CUDA(cudaGraphLaunch(gexec, h->stream));
            h->launches += P.graphLaunches;
        } else {
            P.graphWarm[precond] = true;
            for (int it = 0; it < batch; ++it) {
                if ((rc = enqueueIteration(prof && it == 0))) return rc;
                if (prof && it == 0) sampled = true;
            }
        }
    }
    if (dist && (rc = fvDistGatherPlanes(h, s, v[V_X]))) return rc;   // the assembly kernels around the solve run on the whole box
    PEN_LAUNCH(k_pen_to_nat, g, v[V_X], psi);
    s->pcgIterations += s->hSolve->nIter;
    if (perf) {
        perf->initialResidual = s->hSolve->initRes;
        perf->finalResidual = s->hSolve->finalRes;
        perf->nIterations = s->hSolve->nIter;
    }
    P.precondOf = precond;
    return FY_OK;
}

// Uploads the (component-independent) off-diagonals of the momentum matrix into the pencil layout.
int fvSmoothSetMatrix(fy_ctx* h, FvState* s, const double* lo, const double* up)
{
    PenState& P = s->pen;
    PenMatrix M = penMatrixOf(P, 1);
    M.dg = nullptr;
    PEN_LAUNCH(k_pen_matrix, P.g, (const double*)nullptr, lo, up, M);
    return FY_OK;
}

// smoothSolver + symGaussSeidel (nSweeps 1); off-diagonals as set by fvSmoothSetMatrix
int fvSmoothSolve(fy_ctx* h, FvState* s, const double* dg, const double* b, double* psi, double tol, double relTol,
                  int maxIter, fy_solver_perf* perf)
{
    PenState& P = s->pen;
    const PencilGeom& g = P.g;
    int rc;
    PenMatrix M = penMatrixOf(P, 1);
    double** v = P.v;
    PEN_LAUNCH(k_pen_from_nat, g, dg, M.dg);
    PEN_LAUNCH(k_pen_from_nat, g, b, v[V_B]);
    PEN_LAUNCH(k_pen_from_nat, g, psi, v[V_X]);
    k_solve_begin<<<1, 1, 0, h->stream>>>(s->dSolve, tol, relTol, maxIter, 0);
    FY_CHECK_LAUNCH();
    PEN_LAUNCH(k_pen_avg, g, v[V_X], s->red, s->dSolve);
    PEN_LAUNCH(k_pen_solve_init<false>, g, M, v[V_B], v[V_X], (double*)nullptr, s->red, s->dSolve);
    PEN_LAUNCH(k_pen_arm, g, v[V_MID], (double*)nullptr, (double*)nullptr, s->dSolve);
    for (;;) {
        if ((rc = readSolve(h, s))) return rc;
        if (s->hSolve->done) break;
        for (int it = 0; it < s->gsBatch; ++it) {
            PEN_LAUNCH(k_pen_gs_upper, g, M, v[V_X], v[V_EX], v[V_EY], v[V_EZ], s->dSolve);
            if ((rc = sweepGsFwd(h, s, M, s->dSolve))) return rc;
            if ((rc = sweepGsBwd(h, s, M, s->dSolve))) return rc;
            PEN_LAUNCH(k_pen_residual, g, M, v[V_B], v[V_X], s->red, s->dSolve);
        }
    }
    PEN_LAUNCH(k_pen_to_nat, g, v[V_X], psi);
    if (perf) {
        perf->initialResidual = s->hSolve->initRes;
        perf->finalResidual = s->hSolve->finalRes;
        perf->nIterations = s->hSolve->nIter;
    }
    return FY_OK;
}

int fvDicPrecondition(fy_ctx* h, FvState* s, const double* dg, const double* up, const double* rA, double* wA)
{
    PenState& P = s->pen;
    const PencilGeom& g = P.g;
    int rc;
    PenMatrix M = penMatrixOf(P, 0);
    double** v = P.v;
    P.precondOf = -1;
    PEN_LAUNCH(k_pen_matrix, g, dg, up, up, M);
    PEN_LAUNCH(k_pen_from_nat, g, rA, v[V_RA]);
    PEN_LAUNCH(k_pen_arm, g, v[V_D], v[V_YA], v[V_ZA], (const FvSolveDev*)nullptr);
    if ((rc = sweepDicD(h, s, g, M, nullptr))) return rc;
    if ((rc = sweepDicFwd(h, s, g, M, nullptr))) return rc;
    if (!(P.dbg & 64) && (rc = sweepDicBwd(h, s, g, M, nullptr))) return rc;      // dbg 64: dev probe of the forward sweep
    PEN_LAUNCH(k_pen_to_nat, g, v[V_ZA], wA);
    FY_CUDA(cudaMemcpyAsync(P.hError, P.error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    if (*P.hError) {
        h->err = "pencil pipeline: a neighbour value never arrived (poll limit reached)";
        return FY_ERR_CUDA;
    }
    return FY_OK;
}
