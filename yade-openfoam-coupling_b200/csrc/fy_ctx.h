// fy_ctx.h -- internal state behind the opaque fy_handle of include/fycuda.h.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fycuda.h"

// One k-d tree node, 32 B so that a visit is a single aligned 256-bit load.  The tree is stored
// IMPLICITLY: the node of the index range [lo,hi) sits at lo + (hi-lo)/2, its left subtree is
// [lo,md), its right subtree (md,hi) -- exactly the recursion of meshTree.C:19-37, so no child
// links are needed.
struct __align__(32) FyKdNode {
    double x, y, z;
    int id;
    int pad;
};

static const int FY_MAXLIST = 12;   // meshTree.C:153 `maxelem`

template <class T>
struct FyBuf {               // grow-only device buffer
    T* p = nullptr;
    size_t cap = 0;
};

struct FyPatch {
    int nFaces = 0;
    int start = 0;           // offset into the concatenated boundary-face arrays
    int bcU = 0, bcP = 0;
    double valueU[3] = {0, 0, 0};
    double valueP = 0;
};

struct FvMatrixDev;          // fv_solver.h

struct fy_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // overlapped wire transfers (fy_particles_upload_async / fy_coupling_proc_staged / fy_results_wait)
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evUp = nullptr, evProc = nullptr, evDown = nullptr;
    int stagedN = 0;
    std::string err;
    long long launches = 0;

    // ---- mesh
    int nCells = 0, nFaces = 0, nBFaces = 0;
    double* dC = nullptr;         // [N][3]
    double* dV = nullptr;         // [N]
    int boxN[3] = {0, 0, 0};
    double boxGeom[6] = {0, 0, 0, 1, 1, 1};
    double bbox[6] = {0, 0, 0, 0, 0, 0};
    double V0 = 0;
    // LDU addressing + face geometry (internal faces)
    int *dOwner = nullptr, *dNeigh = nullptr;
    int *dOwnStart = nullptr;     // [N+1] faces owned by cell c: [ownStart[c], ownStart[c+1])
    int *dLosort = nullptr;       // [Fi]  faces sorted by neighbour cell
    int *dLosortStart = nullptr;  // [N+1]
    double *dSf = nullptr, *dMagSf = nullptr, *dWeights = nullptr, *dDeltaCoeffs = nullptr;
    // boundary faces, all patches concatenated
    std::vector<FyPatch> patches;
    int *dBFaceCells = nullptr;   // [nB]
    int *dBPatch = nullptr;       // [nB] patch id of each boundary face
    double *dBSf = nullptr, *dBMagSf = nullptr, *dBDeltaCoeffs = nullptr;
    int *dBStart = nullptr;       // [N+1] boundary faces of cell c in bOrder
    int *dBOrder = nullptr;       // [nB]  boundary faces sorted by cell (stable: patch order kept)
    std::vector<int> hOwner, hNeigh;   // host copies for level scheduling

    // ---- k-d tree
    FyKdNode* dTree = nullptr;
    int nTree = 0;

    // ---- properties / constants (FoamYade.C:9-11, 69-72)
    double rhoP = 0, rhoF = 0, nu = 0;
    bool gaussian = false;
    bool propsSet = false;
    double interpRange = 0, sigmaInterp = 0, interpRangeCu = 0, sigmaPi = 0, maxDist = 0;
    double deltaT = 0;
    // Gaussian-branch options (fy_set_gaussian_options): full-support cell sets instead of the k-d trail;
    // addedMassForce (F.C:392-413) and the Gaussian torque (F.C:467-478), which the reference defines but never calls
    bool supportFull = false, addedMass = false, gaussTorque = false;
    double* dAxis = nullptr;      // [nx + ny + nz] cell-centre coordinates of the hex box (null: centres are no tensor product)

    // ---- coupling fields on the device
    double* dField[FY_F_COUNT] = {nullptr};
    // host bindings (fy_bind_host_fields)
    const double* hIn[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};      // U gradP vGrad divT ddtU
    double* hOut[4] = {nullptr, nullptr, nullptr, nullptr};                     // uSourceDrag alpha uSource uParticle

    // ---- per-cell accumulators of one YadeProc (FoamYade.H:50-51 pVolContrib / uParticleContrib)
    double* dPvol = nullptr;      // [N]
    double* dUpAcc = nullptr;     // [N][3]
    int* dStamp = nullptr;        // [N] serial number of the last proc that touched the cell
    int procSerial = 0;

    // ---- particle buffers
    FyBuf<double> dPdata;         // [n][10]
    FyBuf<int> dFound;            // [n]
    FyBuf<double> dForce;         // [n][6]
    FyBuf<int> dIds;              // [12][n] cell lists, in SORTED particle order (structure of arrays)
    FyBuf<int> dCnt;              // [n]
    FyBuf<double> dW;             // [12][n] normalised weights
    FyBuf<int> dCell;             // [n] point-force cell
    FyBuf<double> dAllWt;         // [n] weight normaliser of the full-support mode (sorted order)
    // position sort of the current buffer (Gaussian mode): keys, identity, permutation (sorted slot -> wire index)
    FyBuf<unsigned int> dKey, dKey2;
    FyBuf<int> dIdx, dPerm;
    FyBuf<char> dSortTmp;
    FyBuf<int> dListCnt, dListIds;   // staging of fy_get_last_lists (wire order)
    FyBuf<double> dListW;
    int lastN = 0;

    // ---- profiling
    bool profiling = false;
    cudaEvent_t ev[8] = {nullptr};
    double phaseMs[8] = {0};

    // ---- FV / PISO state (fv_*.cu)
    struct FvState* fv = nullptr;
};

// every extern "C" entry point runs on its handle's device whatever the caller's current device is (function
// attributes, cudaMalloc and launches are per device); the caller's device is restored on return
struct FyDeviceGuard {
    int prev = -1;
    explicit FyDeviceGuard(const fy_ctx* h)
    {
        if (!h) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != h->device) {
            prev = cur;
            cudaSetDevice(h->device);
        }
    }
    ~FyDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define FY_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
            return FY_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define FY_CHECK_LAUNCH()                                                                   \
    do {                                                                                    \
        h->launches++;                                                                      \
        cudaError_t e_ = cudaGetLastError();                                                \
        if (e_ != cudaSuccess) {                                                            \
            h->err = std::string("kernel launch: ") + cudaGetErrorString(e_);              \
            return FY_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

template <class T>
inline int fyReserve(fy_ctx* h, FyBuf<T>& b, size_t n)
{
    if (n <= b.cap) return FY_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc((void**)&b.p, want * sizeof(T));
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return FY_ERR_ALLOC;
    }
    b.cap = want;
    return FY_OK;
}

static inline int fyGrid(long long n, int block) { return (int)((n + block - 1) / block); }

// kd tree (kdtree_host.cpp)
void fyBuildKdTree(const double* C, int n, std::vector<FyKdNode>& out);

// coupling kernels (coupling.cu)
int fyLaunchLocate(fy_ctx* h, const double* d_xyz, int stride, int n, int* d_ids, int* d_cnt);
int fyLaunchFindCell(fy_ctx* h, const double* d_xyz, int stride, int n, int* d_cell);
int fyCouplingProcDevice(fy_ctx* h, const double* d_pdata, int n, int* d_found, double* d_force);
int fySortParticles(fy_ctx* h, const double* d_pdata, int n);
int fyUnpermuteLists(fy_ctx* h, int n, int* d_cnt, int* d_ids, double* d_wts);
int fyUnpermuteCounts(fy_ctx* h, int n, int* d_cnt);
int fyCouplingPass(fy_ctx* h, int pass, const double* d_pdata, int n, int* d_found, double* d_force);
int fySourceZeroDevice(fy_ctx* h);
int fyInitCouplingFields(fy_ctx* h);
