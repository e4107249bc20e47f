// fv_dist.cu -- decomposition of the pressure solve over the GPUs of one box (one process per GPU).
//
// What is decomposed: the PCG solve of the pressure equation (icoFoamYade.C:125; 90 % of a coupled time step), over a
// Py x Pz grid of ranks (decomposePar `simple`, n = (1 Py Pz)): rank = rz * Py + ry owns the j-blocks [jbLo, jbHi) of
// the k-planes [kLo, kHi).  Every rank keeps the GLOBAL pencil layout and restricts its row loops to its region; the
// neighbours' boundary rows are its ghosts.  In that layout a k-plane's j-block range is ONE contiguous block, so the z
// halo is sent in place; the y halo (one lane of one j-block per plane) goes through a small packed buffer.  y comes first
// when the grid is chosen: a wavefront sweep over one rank's region pays a fixed cost per j-block it crosses, so cutting
// in y shortens the critical path of the preconditioner, cutting in z only its width.  Per PCG iteration:
//     halo exchange of the search direction pA (z: 1 plane segment up / down; y: 1 edge lane up / down;
//         ncclSend / ncclRecv, one group)                                                                  before Amul
//     all-reduce of wA.rA (after the preconditioner), wA.pA (after Amul), sum|rA| (after the update): 1 double each
// The preconditioner is OpenFOAM's own decomposed behaviour: DIC works on the rank's local matrix only (the coupling
// coefficients towards the ghost planes are dropped from the factorisation and the substitutions, [OF-6]
// DICPreconditioner over a processor's lduMatrix), so the sweeps of different ranks are independent.  At the end the
// regions of the solution are gathered on every rank (grouped ncclBroadcast, one root per rank), because the FV
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
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    SYM(AllGather, "ncclAllGather")
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

// the region of rank r: planes of z slab r / Py, j-blocks of y slab r % Py
PencilGeom fvDistGeomOf(const PenState& P, int r)
{
    PencilGeom g = P.g;
    const int ry = r % P.Py, rz = r / P.Py;
    fvSlabRange(P.g.nz, rz, P.Pz, g.kLo, g.kHi);
    fvSlabRange(P.g.nJB, ry, P.Py, g.jbLo, g.jbHi);
    g.nLoc = (long long)(g.kHi - g.kLo) * (g.jbHi - g.jbLo) * g.Tp;
    return g;
}

// Maps the peers' mailboxes and search-direction vectors into this process (CUDA IPC; the handles travel through the NCCL
// communicator that exists by now) and builds the device-side peer table of fv_peer.cuh.  Any failure leaves the NCCL
// path in place and the reason in P.peerWhy.
static int peerSetup(fy_ctx* h, FvState* s)
{
    PenState& P = s->pen;
    const int n = P.nranks, me = P.rank;
    if (n > FY_PEER_MAXR) { P.peerWhy = "more ranks than mailbox slots"; return FY_OK; }
    struct Handles { cudaIpcMemHandle_t mail, pa; };
    static_assert(sizeof(Handles) == 128, "two 64-byte IPC handles");
    Handles mine;
    size_t guard = 0;
    double* const pa = penSearchDir(P, &guard);
    FY_CUDA(cudaMalloc((void**)&P.peerMail, sizeof(PeerMail)));
    FY_CUDA(cudaMemsetAsync(P.peerMail, 0, sizeof(PeerMail), h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));           // zeroed before any peer can learn the address
    bool ok = cudaIpcGetMemHandle(&mine.mail, P.peerMail) == cudaSuccess && cudaIpcGetMemHandle(&mine.pa, pa - guard) == cudaSuccess;
    if (!ok) { cudaGetLastError(); std::memset(&mine, 0, sizeof(mine)); }
    // all-gather {handles, ok}: [n][136 bytes]
    const size_t rec = sizeof(Handles) + 8;
    unsigned char* dAll = nullptr;
    FY_CUDA(cudaMalloc((void**)&dAll, rec * n));
    std::vector<unsigned char> hAll(rec * n, 0);
    std::memcpy(hAll.data() + rec * me, &mine, sizeof(mine));
    hAll[rec * me + sizeof(Handles)] = ok ? 1 : 0;
    FY_CUDA(cudaMemcpyAsync(dAll + rec * me, hAll.data() + rec * me, rec, cudaMemcpyHostToDevice, h->stream));
    FY_NCCL(g_nccl.AllGather(dAll + rec * me, dAll, rec, ncclChar, (ncclComm_t)P.comm, h->stream));
    FY_CUDA(cudaMemcpyAsync(hAll.data(), dAll, rec * n, cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(dAll);
    for (int r = 0; r < n; ++r) ok = ok && hAll[rec * r + sizeof(Handles)] == 1;
    PeerDev pd;
    std::memset(&pd, 0, sizeof(pd));
    pd.rank = me;
    pd.nranks = n;
    pd.error = P.error;
    // neighbours on the Py x Pz grid: zlo zhi ylo yhi
    pd.nbr[0] = P.rz > 0 ? me - P.Py : -1;
    pd.nbr[1] = P.rz < P.Pz - 1 ? me + P.Py : -1;
    pd.nbr[2] = P.ry > 0 ? me - 1 : -1;
    pd.nbr[3] = P.ry < P.Py - 1 ? me + 1 : -1;
    auto open = [&](const cudaIpcMemHandle_t& hd, void** out) {
        if (cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            P.peerWhy = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(cudaGetLastError());
            return false;
        }
        P.peerOpened[P.nPeerOpened++] = *out;
        return true;
    };
    if (!ok) P.peerWhy = "cudaIpcGetMemHandle failed on a rank";
    for (int r = 0; r < n && ok; ++r) {
        const Handles* hr = reinterpret_cast<const Handles*>(hAll.data() + rec * r);
        if (r == me) { pd.box[r] = P.peerMail; continue; }
        void* m = nullptr;
        ok = open(hr->mail, &m);
        pd.box[r] = (PeerMail*)m;
        for (int d = 0; d < 4 && ok; ++d) {
            if (pd.nbr[d] != r) continue;
            void* q = nullptr;
            ok = open(hr->pa, &q);                       // (a rank can be the neighbour on two sides only when a grid
            pd.pa[d] = ok ? (double*)q + guard : nullptr; // dimension is 2 and periodic -- it is not: at most one d matches)
        }
    }
    // every rank takes the same path: agree on `ok`
    int* dOk = nullptr;
    FY_CUDA(cudaMalloc((void**)&dOk, sizeof(int)));
    int hOk = ok ? 1 : 0;
    FY_CUDA(cudaMemcpyAsync(dOk, &hOk, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    FY_NCCL(g_nccl.AllReduce(dOk, dOk, 1, ncclInt, ncclMin, (ncclComm_t)P.comm, h->stream));
    FY_CUDA(cudaMemcpyAsync(&hOk, dOk, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(dOk);
    if (!hOk) {
        if (P.peerWhy.empty()) P.peerWhy = "a peer could not map this rank's buffers";
        return FY_OK;
    }
    FY_CUDA(cudaMalloc((void**)&P.peer, sizeof(PeerDev)));
    FY_CUDA(cudaMemcpyAsync(P.peer, &pd, sizeof(pd), cudaMemcpyHostToDevice, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}

// py = 0: chosen here -- as many y slabs as there are j-blocks to give (see the header), the rest in z
int fvDistInit(fy_ctx* h, FvState* s, int rank, int nranks, int py, const char id[FY_DIST_ID_BYTES])
{
    PenState& P = s->pen;
    if (P.dist || P.comm) { h->err = "fy_dist_init: already initialised"; return FY_ERR_INVALID; }
    if (nranks < 1 || rank < 0 || rank >= nranks) { h->err = "fy_dist_init: bad rank / size"; return FY_ERR_INVALID; }
    if (P.ver != 2) { h->err = "fy_dist_init: the decomposed solve needs the second-generation sweeps (FY_PENCIL_VER=2)"; return FY_ERR_UNSUPPORTED; }
    if (py == 0)
        if (const char* e = std::getenv("FY_DIST_PY")) py = std::atoi(e);
    if (py == 0) {
        py = 1;
        for (int c = 1; c <= nranks && c <= P.g.nJB; ++c)
            if (nranks % c == 0) py = c;
    }
    if (py < 1 || nranks % py != 0 || py > P.g.nJB) { h->err = "fy_dist_init: py must divide the number of ranks and not exceed the number of 32-row j-blocks"; return FY_ERR_INVALID; }
    const int pz = nranks / py;
    if (pz > P.g.nz) { h->err = "fy_dist_init: more z slabs than k-planes"; return FY_ERR_INVALID; }
    if (!loadNccl()) { h->err = g_ncclErr; return FY_ERR_UNSUPPORTED; }
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    FY_NCCL(g_nccl.CommInitRank(&comm, nranks, uid, rank));
    P.comm = comm;
    P.rank = rank;
    P.nranks = nranks;
    P.Py = py;
    P.Pz = pz;
    P.ry = rank % py;
    P.rz = rank / py;
    P.gl = fvDistGeomOf(P, rank);
    P.kLo = P.gl.kLo;
    P.kHi = P.gl.kHi;
    FY_CUDA(cudaMalloc((void**)&P.distBuf, 8 * sizeof(double)));
    FY_CUDA(cudaMalloc((void**)&P.perfBuf, 12 * sizeof(double)));
    if (py > 1) {
        int nzMax = 0;
        for (int r = 0; r < pz; ++r) { int lo, hi; fvSlabRange(P.g.nz, r, pz, lo, hi); nzMax = std::max(nzMax, hi - lo); }
        FY_CUDA(cudaMalloc((void**)&P.yBuf, (size_t)4 * nzMax * P.g.Tp * sizeof(double)));
        FY_CUDA(cudaMalloc((void**)&P.gatherBuf, (size_t)P.g.NP * sizeof(double)));
    }
    P.dist = nranks > 1;
    P.precondOf = -1;
    // (graphs captured so far hold the single-domain iteration)
    for (auto& ge : P.pcgGraph) if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
    for (auto& w : P.graphWarm) w = false;
    const char* e = std::getenv("FY_DIST_PEER");
    if (e && std::atoi(e) == 0) { P.peerWhy = "FY_DIST_PEER=0"; return FY_OK; }
    return P.dist ? peerSetup(h, s) : FY_OK;
}

void fvDistDestroy(FvState* s)
{
    PenState& P = s->pen;
    if (P.comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)P.comm);
    P.comm = nullptr;
    for (int q = 0; q < P.nPeerOpened; ++q) cudaIpcCloseMemHandle(P.peerOpened[q]);
    P.nPeerOpened = 0;
    if (P.peer) cudaFree(P.peer);
    if (P.peerMail) cudaFree(P.peerMail);
    P.peer = nullptr;
    P.peerMail = nullptr;
    if (P.distBuf) cudaFree(P.distBuf);
    if (P.perfBuf) cudaFree(P.perfBuf);
    P.perfBuf = nullptr;
    if (P.yBuf) cudaFree(P.yBuf);
    if (P.gatherBuf) cudaFree(P.gatherBuf);
    P.distBuf = P.yBuf = P.gatherBuf = nullptr;
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

// ghosts of a pencil-layout vector: the j-block range of plane kLo-1 / kHi from the ranks below / above in z, and the
// edge lanes next to the j-block range from the ranks below / above in y (a 7-point stencil needs no corners)
int fvDistHalo(fy_ctx* h, FvState* s, double* v)
{
    PenState& P = s->pen;
    const PencilGeom& gl = P.gl;
    const size_t seg = (size_t)(gl.jbHi - gl.jbLo) * gl.Tp * 32;        // one plane's rows of this rank: contiguous
    const size_t n = (size_t)(gl.kHi - gl.kLo) * gl.Tp;                  // one y edge
    auto at = [&](int k) { return v + ((size_t)k * gl.nJB + gl.jbLo) * gl.Tp * 32; };
    ncclComm_t comm = (ncclComm_t)P.comm;
    const bool yLo = P.ry > 0, yHi = P.ry < P.Py - 1, zLo = P.rz > 0, zHi = P.rz < P.Pz - 1;
    int rc;
    if ((yLo || yHi) && (rc = penPackYEdge(h, s, gl, v, P.yBuf))) return rc;
    double* const recv = P.yBuf + 2 * n;
    FY_NCCL(g_nccl.GroupStart());
    if (zLo) {
        FY_NCCL(g_nccl.Send(at(gl.kLo), seg, ncclDouble, P.rank - P.Py, comm, h->stream));
        FY_NCCL(g_nccl.Recv(at(gl.kLo - 1), seg, ncclDouble, P.rank - P.Py, comm, h->stream));
    }
    if (zHi) {
        FY_NCCL(g_nccl.Send(at(gl.kHi - 1), seg, ncclDouble, P.rank + P.Py, comm, h->stream));
        FY_NCCL(g_nccl.Recv(at(gl.kHi), seg, ncclDouble, P.rank + P.Py, comm, h->stream));
    }
    if (yLo) {
        FY_NCCL(g_nccl.Send(P.yBuf, n, ncclDouble, P.rank - 1, comm, h->stream));
        FY_NCCL(g_nccl.Recv(recv, n, ncclDouble, P.rank - 1, comm, h->stream));
    }
    if (yHi) {
        FY_NCCL(g_nccl.Send(P.yBuf + n, n, ncclDouble, P.rank + 1, comm, h->stream));
        FY_NCCL(g_nccl.Recv(recv + n, n, ncclDouble, P.rank + 1, comm, h->stream));
    }
    FY_NCCL(g_nccl.GroupEnd());
    if ((yLo || yHi) && (rc = penUnpackYEdge(h, s, gl, recv, v))) return rc;
    P.distCollectives++;
    P.distHaloBytes += ((long long)(zLo + zHi) * (long long)seg + (long long)(yLo + yHi) * (long long)n) * 8;
    return FY_OK;
}

// every rank ends up with all of v (each rank is the root of its own region)
int fvDistGatherPlanes(fy_ctx* h, FvState* s, double* v)
{
    PenState& P = s->pen;
    ncclComm_t comm = (ncclComm_t)P.comm;
    if (P.Py == 1) {                                       // z slabs: whole planes, contiguous in place
        const size_t n = (size_t)P.g.zStride;
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
    int rc;
    std::vector<size_t> off(P.nranks + 1, 0);
    for (int r = 0; r < P.nranks; ++r) off[r + 1] = off[r] + (size_t)fvDistGeomOf(P, r).nLoc * 32;
    if ((rc = penPackRegion(h, s, P.gl, v, P.gatherBuf + off[P.rank]))) return rc;
    FY_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < P.nranks; ++r) {
        double* p = P.gatherBuf + off[r];
        FY_NCCL(g_nccl.Broadcast(p, p, off[r + 1] - off[r], ncclDouble, r, comm, h->stream));
    }
    FY_NCCL(g_nccl.GroupEnd());
    for (int r = 0; r < P.nranks; ++r)
        if (r != P.rank && (rc = penUnpackRegion(h, s, fvDistGeomOf(P, r), P.gatherBuf + off[r], v))) return rc;
    P.distCollectives++;
    return FY_OK;
}

// cnt buffers, buffer q from rank root[q] to everybody (one NCCL group, on the handle's stream)
int fvDistBroadcastMany(fy_ctx* h, FvState* s, int cnt, double* const* ptr, const size_t* n, const int* root)
{
    PenState& P = s->pen;
    ncclComm_t comm = (ncclComm_t)P.comm;
    FY_NCCL(g_nccl.GroupStart());
    for (int q = 0; q < cnt; ++q) FY_NCCL(g_nccl.Broadcast(ptr[q], ptr[q], n[q], ncclDouble, root[q], comm, h->stream));
    FY_NCCL(g_nccl.GroupEnd());
    P.distCollectives++;
    return FY_OK;
}
