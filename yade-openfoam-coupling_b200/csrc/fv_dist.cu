// fv_dist.cu -- z-slab decomposition of the pressure solve over the GPUs of one box (one process per GPU).
//
// What is decomposed: the PCG solve of the pressure equation (icoFoamYade.C:125; 90 % of a coupled time step).  Rank r
// owns the k-planes [kLo, kHi) of the box.  In the pencil layout a k-plane is ONE contiguous block of zStride doubles,
// so every rank keeps the global layout and simply restricts its row loops to its planes; the neighbours' boundary
// planes are its ghosts.  Per PCG iteration:
//     halo exchange of the search direction pA (1 plane up, 1 plane down: ncclSend / ncclRecv, grouped)   before Amul
//     all-reduce of wA.rA (after the preconditioner), wA.pA (after Amul), sum|rA| (after the update): 1 double each
// The preconditioner is OpenFOAM's own decomposed behaviour: DIC works on the rank's local matrix only (the coupling
// coefficients towards the ghost planes are dropped from the factorisation and the substitutions, [OF-6]
// DICPreconditioner over a processor's lduMatrix), so the sweeps of different ranks are independent.  At the end the
// planes of the solution are gathered on every rank (grouped ncclBroadcast, one root per rank), because the FV
// assembly kernels around the solve run replicated on the full box.
// NCCL is resolved with dlopen at fy_dist_init, so that libfycuda.so itself has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "fv_solver.h"

struct FyNccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};

namespace {
FyNccl g_nccl;
std::string g_ncclErr;

bool loadNccl()
{
    if (g_nccl.lib) return true;
    // a process that already holds an NCCL (torch's) gets that one: same SONAME
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { g_ncclErr = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
    FyNccl f;
    f.lib = lib;
#define SYM(field, name)                                                          \
    *(void**)(&f.field) = dlsym(lib, name);                                       \
    if (!f.field) { g_ncclErr = std::string("dlsym(") + name + ") failed"; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Broadcast, "ncclBroadcast")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
    g_nccl = f;
    return true;
}
}  // namespace

#define FY_NCCL(call)                                                                       \
    do {                                                                                    \
        ncclResult_t r_ = (call);                                                           \
        if (r_ != ncclSuccess) {                                                            \
            h->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_);                \
            return FY_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

void fvSlabRange(int nz, int rank, int nranks, int& kLo, int& kHi)
{
    kLo = (int)(((long long)rank * nz) / nranks);
    kHi = (int)(((long long)(rank + 1) * nz) / nranks);
}

int fvDistUniqueId(char out[FY_DIST_ID_BYTES], std::string& err)
{
    static_assert(sizeof(ncclUniqueId) <= FY_DIST_ID_BYTES, "unique id does not fit");
    if (!loadNccl()) { err = g_ncclErr; return FY_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return FY_ERR_CUDA; }
    std::memset(out, 0, FY_DIST_ID_BYTES);
    std::memcpy(out, &id, sizeof(id));
    return FY_OK;
}

int fvDistInit(fy_ctx* h, FvState* s, int rank, int nranks, const char id[FY_DIST_ID_BYTES])
{
    PenState& P = s->pen;
    if (P.dist) { h->err = "fy_dist_init: already initialised"; return FY_ERR_INVALID; }
    if (nranks < 1 || rank < 0 || rank >= nranks) { h->err = "fy_dist_init: bad rank / size"; return FY_ERR_INVALID; }
    if (nranks > P.g.nz) { h->err = "fy_dist_init: more ranks than k-planes"; return FY_ERR_INVALID; }
    if (P.ver != 2) { h->err = "fy_dist_init: the decomposed solve needs the second-generation sweeps (FY_PENCIL_VER=2)"; return FY_ERR_UNSUPPORTED; }
    if (!loadNccl()) { h->err = g_ncclErr; return FY_ERR_UNSUPPORTED; }
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    FY_NCCL(g_nccl.CommInitRank(&comm, nranks, uid, rank));
    P.comm = comm;
    P.rank = rank;
    P.nranks = nranks;
    fvSlabRange(P.g.nz, rank, nranks, P.kLo, P.kHi);
    P.gl = P.g;
    P.gl.kLo = P.kLo;
    P.gl.kHi = P.kHi;
    P.gl.rowLo = (long long)P.kLo * P.g.nJB * P.g.Tp;
    P.gl.rowHi = (long long)P.kHi * P.g.nJB * P.g.Tp;
    FY_CUDA(cudaMalloc((void**)&P.distBuf, 8 * sizeof(double)));
    P.dist = nranks > 1;
    P.precondOf = -1;
    return FY_OK;
}

void fvDistDestroy(FvState* s)
{
    PenState& P = s->pen;
    if (P.comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)P.comm);
    P.comm = nullptr;
    if (P.distBuf) cudaFree(P.distBuf);
    P.distBuf = nullptr;
    P.dist = false;
}

// in-place SUM over the ranks of n doubles at d (on the handle's stream)
int fvDistAllReduce(fy_ctx* h, FvState* s, double* d, int n)
{
    PenState& P = s->pen;
    FY_NCCL(g_nccl.AllReduce(d, d, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)P.comm, h->stream));
    P.distCollectives++;
    return FY_OK;
}

// ghost planes of a pencil-layout vector: plane kLo-1 from the rank below, plane kHi from the rank above
int fvDistHalo(fy_ctx* h, FvState* s, double* v)
{
    PenState& P = s->pen;
    const size_t n = (size_t)P.g.zStride;
    ncclComm_t comm = (ncclComm_t)P.comm;
    FY_NCCL(g_nccl.GroupStart());
    if (P.rank > 0) {
        FY_NCCL(g_nccl.Send(v + (size_t)P.kLo * n, n, ncclDouble, P.rank - 1, comm, h->stream));
        FY_NCCL(g_nccl.Recv(v + (size_t)(P.kLo - 1) * n, n, ncclDouble, P.rank - 1, comm, h->stream));
    }
    if (P.rank < P.nranks - 1) {
        FY_NCCL(g_nccl.Send(v + (size_t)(P.kHi - 1) * n, n, ncclDouble, P.rank + 1, comm, h->stream));
        FY_NCCL(g_nccl.Recv(v + (size_t)P.kHi * n, n, ncclDouble, P.rank + 1, comm, h->stream));
    }
    FY_NCCL(g_nccl.GroupEnd());
    P.distCollectives++;
    P.distHaloBytes += (long long)((P.rank > 0) + (P.rank < P.nranks - 1)) * (long long)n * 8;
    return FY_OK;
}

// every rank ends up with all planes of v (each rank is the root of its own planes)
int fvDistGatherPlanes(fy_ctx* h, FvState* s, double* v)
{
    PenState& P = s->pen;
    const size_t n = (size_t)P.g.zStride;
    ncclComm_t comm = (ncclComm_t)P.comm;
    FY_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < P.nranks; ++r) {
        int lo, hi;
        fvSlabRange(P.g.nz, r, P.nranks, lo, hi);
        double* p = v + (size_t)lo * n;
        FY_NCCL(g_nccl.Broadcast(p, p, (size_t)(hi - lo) * n, ncclDouble, r, comm, h->stream));
    }
    FY_NCCL(g_nccl.GroupEnd());
    P.distCollectives++;
    return FY_OK;
}
