// fv_solver.h -- internal interface of the finite-volume / PISO half (fv_box.cu).
#pragma once
#include <string>
#include <vector>

#include "fv_box.cuh"
#include "fv_pencil.cuh"
#include "fy_ctx.h"

// solver state shared between the kernels of one linear solve (device memory; the host reads it back
// only every few iterations, the iteration kernels themselves test `done`)
struct FvSolveDev {
    double tol, relTol;
    int maxIter, precond;
    double avg, normFactor, initRes, finalRes;
    double wArA, wArAold, wApA, alpha, beta;
    int nIter, done, singular, pad;
};

// device scalars of one time step (Courant number, continuity errors, adjustPhi)
struct FvStepDev {
    double CoNum, meanCoNum;
    double sumLocal, global;
    double corrSumLocal[8], corrGlobal[8];
    double massIn, fixedMassOut, adjustableMassOut, totalFlux, massCorr;
    int adjustFail, pad;
};

// pencil-layout solver state (fv_pencil.cu): matrices, the shared pool of Krylov / smoother vectors, and
// the bookkeeping of the warp-pencil pipelines
struct PenState {
    PencilGeom g;
    int rowGrid = 0;                    // blocks of the row-structured vector kernels
    int W = 8;                          // compute warps per pencil group (CTA)
    int cluster = 16;                   // largest thread-block cluster (plane groups chained through DSMEM)
    int smemBudget = 200 * 1024;        // bytes of shared memory per pencil group
    // second-generation sweeps (fv_pencil2.cuh)
    int ver = 2;                        // 1: k_pencil (one plane per warp, cp.async), 2: k_pen2 (Z planes per warp, TMA + mbarrier)
    // planes per compute warp, rows per TMA stage, compute warps per CTA: measured best on B200 at 128^3 and 256^3
    // (profiles/r2b_sweep_variants.txt): one plane per warp, eight warps per CTA
    int Z2 = 1, R2 = 4, W2 = 8;
    int maxStage2 = 5;                  // input-ring stages
    int smemBudget2 = 200 * 1024;       // shared memory per CTA the rings may fill
    bool fusedTail = true;              // direction + Amul + update of a PCG iteration as one cooperative kernel (FY_PCG_FUSED=0: three)
    unsigned int* tailBar = nullptr;    // [4] its grid-barrier counters
    int tailBlocksPerSm = 0, tailMaxGrid = 0;
    double* pk[4] = {nullptr};          // premultiplied DIC streams: {rD, rD lowx} {rD lowy, rD lowz} {rD upx, rD upy} rD upz
    // decomposition of the pressure solve over a Py x Pz grid of ranks (fv_dist.cu): rank = rz * Py + ry
    bool dist = false;
    int rank = 0, nranks = 1;
    int Py = 1, Pz = 1, ry = 0, rz = 0;
    int kLo = 0, kHi = 0;               // this rank's k-planes
    PencilGeom gl;                      // the geometry restricted to this rank: planes [kLo, kHi), j-blocks [jbLo, jbHi)
    // the iteration's collectives over NVLink peer memory (fv_peer.cuh; FY_DIST_PEER=0 keeps them on NCCL)
    PeerDev* peer = nullptr;            // device copy of this rank's peer table; null: NCCL path
    PeerMail* peerMail = nullptr;       // this rank's mailbox
    void* peerOpened[2 * FY_PEER_MAXR] = {nullptr};   // IPC mappings to close
    int nPeerOpened = 0;
    std::string peerWhy;                // why the peer path is off, if it is
    double* yBuf = nullptr;             // y-edge halo staging: send[2n] | recv[2n], n = (kHi-kLo)*Tp
    double* gatherBuf = nullptr;
    double* perfBuf = nullptr;          // [3][4] solver statistics of the momentum components travelling with them        // [NP] the ranks' regions back to back (solution gather of a y-decomposed solve)
    void* comm = nullptr;               // ncclComm_t
    double* distBuf = nullptr;          // [8] partial sums handed to the all-reduce
    long long distCollectives = 0, distHaloBytes = 0;
    double* mP[7] = {nullptr};          // pEqn: dg, low[3], up[3]
    double* mU[7] = {nullptr};          // UEqn: dg (current component), low[3], up[3]
    double* v[9] = {nullptr};           // vectors (roles: see fv_pencil.cu)
    double* partial = nullptr;          // per-pencil partial sums
    unsigned long long* trace = nullptr; // [nJB*nz][4] debug time stamps of the last pencil launch (FY_PENCIL_TRACE)
    bool traceOn = false;
    bool useGraphs = true;              // replay a batch of PCG iterations as one CUDA graph (FY_PCG_GRAPH=0 disables)
    cudaGraphExec_t pcgGraph[3] = {nullptr, nullptr, nullptr};     // per preconditioner
    int graphLaunches = 0;              // kernel launches inside one graph replay
    bool graphWarm[3] = {false, false, false};   // the iteration kernels have run eagerly at least once
    int dbg = 0;
    int precondOf = -1;                 // preconditioner whose matrix copy + reciprocal diagonal the last PCG solve left in place
    unsigned int* ticket = nullptr;     // [2]
    int* error = nullptr;
    int* hError = nullptr;              // pinned
};

struct FvState {
    bool supported = false;
    std::string why;
    BoxGeom g;
    int nFi = 0, nB = 0;
    int* dSlotOfFace = nullptr;         // [nFi + nB] OpenFOAM face order -> owner slot
    std::vector<int> hSlotOfFace;
    fy_piso_controls ctl;
    double nu = 0.01;
    double cumulativeContErr = 0;
    fy_ico_stats stats;

    // face fields (owner slots)
    double *phi = nullptr, *phi0 = nullptr, *phiHbyA = nullptr;
    // cell fields
    double *U0 = nullptr, *HbyA = nullptr, *rAU = nullptr, *gradP = nullptr;
    // UEqn: diag, lower/upper in owner slots [3N], source [N][3]; per-component solve arrays (SoA [3][N])
    double *diagU = nullptr, *loU = nullptr, *upU = nullptr, *srcU = nullptr;
    double *dgU = nullptr, *bU = nullptr, *psiU = nullptr;
    // pEqn
    double *upP = nullptr, *dgP = nullptr, *bP = nullptr;
    // pimpleFoamYade extras (allocated by the first fy_pimple_solve): phicForces [slots], explicit stress term [N][3]
    double *phicForces = nullptr, *divDev = nullptr;
    // PIMPLE outer correctors + relaxationFactors (fy_set_pimple_controls; a factor <= 0 = no entry in fvSolution)
    int nOuter = 1;
    double relaxU = 0, relaxUFinal = 0, relaxP = 0, relaxPFinal = 0;
    double *pPrev = nullptr;            // [N] p.prevIter() (allocated when nOuter > 1)
    double *bGradP = nullptr;           // [slots] snGrad(p) of the fixedFluxPressure faces (constrainPressure), else null
    bool hasFluxP = false;
    // scratch for the parity hooks (LDU-order staging)
    double *stage = nullptr;
    size_t stageCap = 0;

    FvRed red{nullptr, nullptr, nullptr, nullptr};
    FvSolveDev* dSolve = nullptr;
    FvSolveDev* hSolve = nullptr;       // pinned
    FvStepDev* dStep = nullptr;
    FvStepDev* hStep = nullptr;         // pinned
    int cellGrid = 0;                   // blocks of the grid-stride cell kernels
    int pcgBatch = 8;
    int gsBatch = 2;
    PenState pen;
    double fluidMs[4] = {0, 0, 0, 0};
    // profiling (fy_set_profiling): device time of each kernel class of the PCG iteration, sampled on the
    // first iteration of every batch: [0] precondition forward [1] backward (+wA.rA) [2] search direction
    // [3] Amul (+wA.pA) [4] update (+|rA|)
    cudaEvent_t pev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double kernelMs[5] = {0, 0, 0, 0, 0};
    long long kernelSamples = 0;
    long long pcgIterations = 0;        // since the last fy_get_kernel_ms reset
};

int fvCreate(fy_ctx* h, const fy_mesh_desc* m);
int fvPcgSolve(fy_ctx* h, FvState* s, const double* dg, const double* up, const double* b, double* psi, double tol,
               double relTol, int maxIter, int precond, fy_solver_perf* perf, bool sameMatrix = false);
int fvSmoothSetMatrix(fy_ctx* h, FvState* s, const double* lo, const double* up);
int fvSmoothSolve(fy_ctx* h, FvState* s, const double* dg, const double* b, double* psi, double tol, double relTol,
                  int maxIter, fy_solver_perf* perf);
void fvSlabRange(int nz, int rank, int nranks, int& kLo, int& kHi);
int fvDistUniqueId(char out[FY_DIST_ID_BYTES], std::string& err);
int fvDistInit(fy_ctx* h, FvState* s, int rank, int nranks, int py, const char id[FY_DIST_ID_BYTES]);
PencilGeom fvDistGeomOf(const PenState& P, int rank);
double* penSearchDir(PenState& P, size_t* guardElems);   // the search-direction vector pA and the guard in front of its allocation
int penPackYEdge(fy_ctx* h, FvState* s, const PencilGeom& g, const double* v, double* buf);
int penUnpackYEdge(fy_ctx* h, FvState* s, const PencilGeom& g, const double* buf, double* v);
int penPackRegion(fy_ctx* h, FvState* s, const PencilGeom& g, const double* v, double* buf);
int penUnpackRegion(fy_ctx* h, FvState* s, const PencilGeom& g, const double* buf, double* v);
void fvDistDestroy(FvState* s);
int fvDistAllReduce(fy_ctx* h, FvState* s, double* d, int n);
int fvDistHalo(fy_ctx* h, FvState* s, double* v);
int fvDistGatherPlanes(fy_ctx* h, FvState* s, double* v);
int fvDistBroadcastMany(fy_ctx* h, FvState* s, int cnt, double* const* ptr, const size_t* n, const int* root);
int penCreate(fy_ctx* h, FvState* s);
void penDestroy(FvState* s);
int fvDicPrecondition(fy_ctx* h, FvState* s, const double* dg, const double* up, const double* rA, double* wA);
int fvCreatePhi(fy_ctx* h, FvState* s);
int fvGradVector(fy_ctx* h, FvState* s, const double* dU, double* dOut);
int fvGradScalar(fy_ctx* h, FvState* s, const double* dP, double* dOut);
int fvDivFlux(fy_ctx* h, FvState* s, const double* dPhiSlots, double* dOut);
int fvFacesToSlots(fy_ctx* h, FvState* s, int n, const double* dFaces, double* dSlots);
int fvSlotsToFaces(fy_ctx* h, FvState* s, int n, const double* dSlots, double* dFaces);
int fvDivPhiVector(fy_ctx* h, FvState* s, const double* dPhiSlots, const double* dU, double* dOut);
int fvLaplacianGammaVector(fy_ctx* h, FvState* s, double scale, const double* dGamma, double gammaB, const double* dU,
                           double* dOut);
int fvIcoPre(fy_ctx* h, FvState* s, double dt);
int fvPimplePre(fy_ctx* h, FvState* s, double dt);
int fvPimpleSolve(fy_ctx* h, FvState* s, double dt, const double gvec[3]);
int fvIcoSolve(fy_ctx* h, FvState* s, double dt);
void fvDestroy(fy_ctx* h);
