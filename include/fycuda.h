/* fycuda.h -- C ABI of the B200-native FoamYade coupling engine (libfycuda.so).
 *
 * This is the drop-in boundary for ONE hot path of dpkn31/Yade-OpenFOAM-coupling:
 * the per-timestep particle<->fluid coupling operator Foam::FoamYade
 * (FoamYade/FoamYade.H:57-161, FoamYade.C) with its k-d-tree mesh search
 * (FoamYade/meshtree/meshTree.{H,C}), and the PISO/PIMPLE pressure-velocity
 * solve of icoFoamYade/icoFoamYade.C:65-149 and pimpleFoamYade/.  The host
 * side (the C++ class Foam::FoamYade in yade-openfoam-coupling_b200/host/, the
 * solver drivers, the Python ctypes binding used by the tests) sits ABOVE this
 * header; everything below it is hand-written sm_100a CUDA.
 *
 * Conventions
 *   - plain C: opaque handle, pointers + sizes, int return code
 *     (0 = FY_OK, < 0 = error; fy_last_error() gives the text). Nothing throws
 *     across the boundary.  There is NO CPU fallback: without a CUDA device
 *     every compute entry point fails with FY_ERR_NO_DEVICE.
 *   - all floating point is IEEE fp64; cell / face indices are int32
 *     (OpenFOAM `label`, `int` in the reference).
 *   - vector fields are arrays of [n][3] doubles, tensors [n][9] row-major
 *     (xx xy xz yx yy yz zx zy zz) -- OpenFOAM's in-memory layout, so a
 *     volVectorField's internal field can be passed as is.
 *   - caller owns host buffers; the library owns device buffers.  One CUDA
 *     stream per handle; a handle is not thread-safe.
 *   - pointers named h_* are host pointers; d_* device pointers.
 */
#ifndef FYCUDA_H
#define FYCUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fy_ctx* fy_handle;

enum {
    FY_OK = 0,
    FY_ERR_INVALID = -1,     /* bad argument / call order */
    FY_ERR_NO_DEVICE = -2,   /* no usable CUDA device: the engine has no CPU path */
    FY_ERR_CUDA = -3,        /* CUDA runtime error, see fy_last_error */
    FY_ERR_ALLOC = -4,
    FY_ERR_NOT_CONVERGED = -5,
    FY_ERR_UNSUPPORTED = -6
};

/* boundary-condition kinds (the subset of OpenFOAM patch fields the synthetic cases use) */
enum {
    FY_BC_FIXED_VALUE = 0,    /* fixedValue / noSlip / movingWall: value given per patch */
    FY_BC_ZERO_GRADIENT = 1,
    FY_BC_EMPTY = 2,          /* OpenFOAM `empty` (2-D cases): the faces take part in nothing */
    FY_BC_FIXED_FLUX_PRESSURE = 3   /* p only: fixedFluxPressure, its gradient set by constrainPressure (pimpleFoamYade/pEqn.H:21;
                                       createFields.H of both solvers declares p so that walls under gravity can carry it) */
};

/* A patch: nFaces boundary faces that all belong to one boundary-condition group.
 * Geometry arrays are per face, in OpenFOAM boundary-face order.                */
typedef struct {
    int nFaces;
    const int* faceCells;        /* [nFaces] owner cell of each boundary face           */
    const double* Sf;            /* [nFaces][3] outward face-area vectors               */
    const double* magSf;         /* [nFaces]                                            */
    const double* deltaCoeffs;   /* [nFaces] 1/|d| (cell centre -> face centre)         */
    int bcU;                     /* FY_BC_* for U                                       */
    double valueU[3];            /* patch-uniform value when bcU == FIXED_VALUE         */
    int bcP;                     /* FY_BC_* for p                                       */
    double valueP;
} fy_patch_desc;

/* Mesh in OpenFOAM LDU form (what fvMesh exposes): cells, internal faces in upper-triangular order
 * (owner < neighbour, sorted by owner then neighbour), face geometry, boundary patches.
 * Replaces: the `const fvMesh&` argument of Foam::FoamYade::FoamYade (FoamYade.H:106) and the mesh
 * every fvm:: / fvc:: call of icoFoamYade.C / pimpleFoamYade.C takes implicitly.                    */
typedef struct {
    int nCells;
    const double* C;             /* [nCells][3] cell centres  (mesh.C(), FoamYade.H:121, meshTree.C:13) */
    const double* V;             /* [nCells]    cell volumes  (mesh.V(), FoamYade.C:69,323)              */
    /* internal faces -- may be 0/NULL when only the coupling operator is used */
    int nInternalFaces;
    const int* owner;            /* [nInternalFaces] */
    const int* neighbour;        /* [nInternalFaces] */
    const double* Sf;            /* [nInternalFaces][3] */
    const double* magSf;         /* [nInternalFaces]    */
    const double* weights;       /* [nInternalFaces] linear-interpolation weight of the owner */
    const double* deltaCoeffs;   /* [nInternalFaces] 1/|C_N - C_P| */
    int nPatches;
    const fy_patch_desc* patches;
    /* "containing cell" query (mesh.findCell, FoamYade.C:251) is index arithmetic on an axis-aligned
     * box of boxN[0] x boxN[1] x boxN[2] hex cells, x fastest; boxN[0] == 0 => point-force mode
     * unavailable on this mesh.  boxGeom = x0 y0 z0 hx hy hz.                                       */
    int boxN[3];
    double boxGeom[6];
    /* mesh vertex bounding box (min xyz, max xyz) -- what sendMeshBbox ships (FoamYade.C:82-96) */
    double bbox[6];
} fy_mesh_desc;

/* ---------------------------------------------------------------------------------------------
 * life cycle
 * ------------------------------------------------------------------------------------------- */

/* Library / device probe. Returns the number of CUDA devices (0 on a CPU-only box, never an error). */
int fy_device_count(void);
const char* fy_version(void);

/* Creates an engine on CUDA device `device`: uploads the mesh, builds the k-d tree over the cell
 * centres on the host with the reference's own construction rule (meshTree.C:9-51: axis = depth%3,
 * std::nth_element at size/2, left = [0,md), right = (md,end)) and uploads it in implicit in-order
 * layout.  Replaces FoamYade::FoamYade + getRankSize()'s mshTree.build_tree() + initFields()
 * (FoamYade.C:18-73).                                                                            */
int fy_create(const fy_mesh_desc* mesh, int device, fy_handle* out);
int fy_destroy(fy_handle h);
const char* fy_last_error(fy_handle h);   /* h may be NULL: last create error */

/* FoamYade::setScalarProperties(rhoP, rhoF, nu) (FoamYade.C:9-11) + the `gaussianInterp` ctor flag
 * (FoamYade.H:117,122).                                                                          */
int fy_set_properties(fy_handle h, double rhoP, double rhoF, double nu, int gaussianInterp);

/* interpRange, sigmaInterp, interpRangeCu, sigmaPi of initFields() (FoamYade.C:69-72), computed on
 * the host exactly as written there.                                                              */
int fy_get_constants(fy_handle h, double out4[4]);

/* Pinned host memory for wire buffers (particle records in, forces out). */
int fy_host_alloc(void** p, size_t bytes);
int fy_host_free(void* p);

/* ---------------------------------------------------------------------------------------------
 * fields the coupling operator binds (FoamYade.H:76-90)
 *   inputs : U gradP vGrad divT ddtU          outputs: uSourceDrag alpha uSource uParticle
 * The engine keeps a device copy of each.  With host binding (an OpenFOAM CPU solver owns the
 * fields) the inputs are uploaded at the start of every fy_set_particle_action and the outputs are
 * downloaded at its end / at fy_set_source_zero; with the built-in GPU solver (fy_ico_*, fy_pimple_*)
 * nothing crosses PCIe.
 * ------------------------------------------------------------------------------------------- */
typedef enum {
    FY_F_U = 0, FY_F_GRADP, FY_F_VGRAD, FY_F_DIVT, FY_F_DDTU,
    FY_F_USOURCEDRAG, FY_F_ALPHA, FY_F_USOURCE, FY_F_UPARTICLE,
    FY_F_P, FY_F_PHI, FY_F_COUNT
} fy_field_id;

/* Host pointers of the solver-owned fields (any may be NULL = not host-bound). */
int fy_bind_host_fields(fy_handle h, const double* U, const double* gradP, const double* vGrad,
                        const double* divT, const double* ddtU, double* uSourceDrag, double* alpha,
                        double* uSource, double* uParticle);
/* Explicit transfers of one field (size is implied by the field id). */
int fy_upload_field(fy_handle h, int field, const double* h_src);
int fy_download_field(fy_handle h, int field, double* h_dst);
/* Device pointer of a field (for callers that already live on the GPU, e.g. torch tensors). */
int fy_device_field(fy_handle h, int field, double** d_ptr);

/* ---------------------------------------------------------------------------------------------
 * the coupling operator
 * ------------------------------------------------------------------------------------------- */

/* meshTree::nnearestCellsRange (meshTree.C:148-179) for n points: h_ids [n][12] nearest-first,
 * -1 padded; h_counts [n] in 0..12 (the canonical form: the last <= 12 strict improvements of the
 * nearest-neighbour descent with d^2 < 1.25 range^2).  Parity hook; bit-exact.                     */
int fy_locate(fy_handle h, const double* h_xyz, int n, int* h_ids, int* h_counts);

/* mesh.findCell (FoamYade.C:251) on the hex box: h_cell [n], -1 outside. */
int fy_find_cell(fy_handle h, const double* h_xyz, int n, int* h_cell);

/* One fluid step of FoamYade::setParticleAction (FoamYade.C:605-632), split so that several Yade
 * ranks' buffers can be processed in the reference's order (each YadeProc: locate -> weights ->
 * per-cell accumulate -> void fraction -> forces; FoamYade.C:612-628):
 *   fy_coupling_begin : deltaT = dt; uploads host-bound input fields
 *   fy_coupling_proc  : one particle buffer: h_pdata [n][10] = x y z vx vy vz wx wy wz radius
 *                       (FoamYade.C:190-219) -> h_found [n] (1 / -1, FoamYade.C:141,222),
 *                       h_force [n][6] = Fx Fy Fz Tx Ty Tz (FoamYade.C:492-498), zero when not found
 *   fy_coupling_end   : downloads host-bound output fields
 * fy_set_particle_action = begin + one proc + end (serial-Yade mode, FoamYade.C:173-184).          */
int fy_coupling_begin(fy_handle h, double dt);
int fy_coupling_proc(fy_handle h, const double* h_pdata, int n, int* h_found, double* h_force);
int fy_coupling_end(fy_handle h);
int fy_set_particle_action(fy_handle h, double dt, const double* h_pdata, int n, int* h_found, double* h_force);

/* Overlapped wire transfers for a solver that runs its fluid step on the device (fy_ico_* / fy_pimple_*): the particle
 * records of the step (caller-owned PINNED host memory, see fy_host_alloc) come up on a second stream while the
 * pre-coupling block (fy_ico_pre: CourantNo, grad(U); icoFoamYade.C:68-71) runs, and found / force go down while the
 * pressure-velocity solve (fy_ico_solve) runs.  Same arithmetic as fy_coupling_proc (FoamYade.C:612-628):
 *   fy_particles_upload_async(h, h_pdata, n)    starts the upload, returns at once
 *   fy_coupling_proc_staged(h, h_found, h_force) the coupling operator on the staged records (after fy_coupling_begin);
 *                                                queues the download, returns without waiting
 *   fy_results_wait(h)                           h_found / h_force are valid after it returns (before the reply to Yade)  */
int fy_particles_upload_async(fy_handle h, const double* h_pdata, int n);
int fy_coupling_proc_staged(fy_handle h, int* h_found, double* h_force);
int fy_results_wait(fy_handle h);

/* Device-resident variant of fy_coupling_proc: all three buffers are device pointers; no PCIe traffic,
 * no host synchronisation.                                                                            */
int fy_coupling_proc_device(fy_handle h, const double* d_pdata, int n, int* d_found, double* d_force);

/* Particle-sharded multi-GPU coupling: the passes of fy_coupling_proc_device one at a time, so
 * that ranks holding different slices of ONE Yade buffer can sum the per-cell partial results between them (the
 * reference does the same sum with MPI_Allreduce over the Foam ranks, FoamYade.C:511-516, for the forces):
 *   pass 0  locate + Gaussian weights + per-cell accumulate (FoamYade.C:187-225, 293-316, 261-290) -> d_found;
 *           then all-reduce SUM the accumulators (fy_device_accumulators: pvol [N], upAcc [N][3]) and MAX the stamps [N]
 *   pass 1  setCellVolFraction (FoamYade.C:318-328) from the reduced accumulators: identical on every rank
 *   pass 2  forces of this rank's particles (FoamYade.C:331-453) -> d_force (+ d_found in point-force mode); then
 *           all-reduce SUM uSource [N][3] and uSourceDrag [N]
 * Every rank must make the same sequence of calls.                                                                  */
int fy_coupling_pass_device(fy_handle h, int pass, const double* d_pdata, int n, int* d_found, double* d_force);
int fy_device_accumulators(fy_handle h, double** d_pvol, double** d_upAcc, int** d_stamp);
/* The handle's CUDA stream (a cudaStream_t), so that a caller can order its own work (e.g. NCCL) with the engine's. */
int fy_stream(fy_handle h, void** cuda_stream);

/* FoamYade::setSourceZero (FoamYade.C:556-566): uSource = 0; Gaussian: alpha = 1, uSourceDrag = 0,
 * uParticle = 0.  Host-bound output fields are reset too.                                          */
int fy_set_source_zero(fy_handle h);

/* Options of the Gaussian branch beyond what the reference's loop runs; defaults = the reference:
 *   support    FY_SUPPORT_TRAIL  the <= 12 cells of meshTree::nnearestCellsRange's improvement trail (meshTree.C:148-238), or
 *              FY_SUPPORT_FULL   "range based search" (README.md:5): EVERY cell whose centre lies within the same bound
 *                                (d^2 < 1.25 interpRange^2, meshTree.C:155) carries a Gaussian weight -- ~370 cells per
 *                                particle; one warp per particle, warp-shuffle reductions over the support; hex-box meshes.
 *                                A particle is found when at least one cell lies inside the bound (the trail mode inherits
 *                                meshTree's quirk that the tree's root cell is never reported, meshTree.C:156,192)
 *   addedMass  also apply addedMassForce (FoamYade.C:392-413: defined, never called by the reference) after archimedesForce
 *   torque     also apply calcHydroTorque's Gaussian branch (FoamYade.C:467-478, commented out at FoamYade.C:618): the
 *              torque slots of the force record are then filled in Gaussian mode too
 * Parity: the unmodified reference's own weight / force functions fed with the full cell sets, and its own
 * addedMassForce / calcHydroTorque (oracle/ref_harness.cpp ref_set_gaussian_options).                                   */
enum { FY_SUPPORT_TRAIL = 0, FY_SUPPORT_FULL = 1 };
int fy_set_gaussian_options(fy_handle h, int support, int addedMass, int torque);

/* Cell lists + normalised Gaussian weights of the last fy_coupling_proc (parity hooks; FY_SUPPORT_FULL: counts only):
 * h_counts [n], h_ids [n][12], h_weights [n][12] (FoamYade.C:293-316). Any pointer may be NULL.     */
int fy_get_last_lists(fy_handle h, int n, int* h_counts, int* h_ids, double* h_weights);

/* ---------------------------------------------------------------------------------------------
 * the fluid half: icoFoamYade's time step (icoFoamYade/icoFoamYade.C:65-149) on the device
 *
 * The arithmetic is OpenFOAM-6's (Euler ddt; Gauss linear grad / div / laplacian; linear interpolation;
 * p: PCG + DIC / diagonal / none, U: smoothSolver + symGaussSeidel -- the stock cavity fvSchemes /
 * fvSolution, which the reference does not ship).  Supported meshes: a uniform hex box in blockMesh
 * order (fy_mesh_desc.boxN set, LDU faces and patches consistent with it) whose boundary sides each
 * carry one boundary condition; fy_fv_supported() says whether the mesh qualified.
 * ------------------------------------------------------------------------------------------- */
enum { FY_PRECOND_DIC = 0, FY_PRECOND_DIAGONAL = 1, FY_PRECOND_NONE = 2 };

/* fvSolution: PISO sub-dictionary (icoFoamYade/createFields.H:166-168 reads pRefCell / pRefValue from it)
 * and the p / pFinal / U solver entries.                                                              */
typedef struct {
    int nCorrectors;                /* PISO nCorrectors                      (stock cavity: 2) */
    int nNonOrthogonalCorrectors;   /*                                       (0)               */
    int momentumPredictor;          /* piso.momentumPredictor()              (1)               */
    int pRefCell;
    double pRefValue;
    double pTol, pRelTol;           /* p      { tolerance 1e-06; relTol 0.05; } */
    double pFinalTol, pFinalRelTol; /* pFinal { $p; relTol 0; }                 */
    double UTol, URelTol;           /* U      { tolerance 1e-05; relTol 0; }    */
    int maxIter;                    /* 1000 */
    int preconditioner;             /* FY_PRECOND_* for p */
} fy_piso_controls;

typedef struct {
    double initialResidual, finalResidual;
    int nIterations;
    int pad_;
} fy_solver_perf;

/* What the solver log of one time step prints: Courant number, the segregated U solves, every
 * pressure solve in order, continuityErrs.H after each PISO corrector.                          */
typedef struct {
    double CoNum, meanCoNum;
    fy_solver_perf U[3];
    fy_solver_perf p[8];
    int nPSolves;
    int pad_;
    double sumLocalContErr, globalContErr, cumulativeContErr;
    double corrSumLocal[8], corrGlobal[8];
} fy_ico_stats;

/* 1 when the mesh given to fy_create qualified for the device FV path, else 0 (fy_last_error says why). */
int fy_fv_supported(fy_handle h);
/* defaults = the stock cavity set above */
int fy_piso_default_controls(fy_piso_controls* c);
int fy_set_piso_controls(fy_handle h, const fy_piso_controls* c);
/* pimpleFoamYade's PIMPLE sub-dictionary and relaxationFactors (fy_pimple_solve only):
 *   nOuterCorrectors   `while (pimple.loop())`, pimpleFoamYade/pimpleFoamYade.C:91-105: UcEqn.H + the PISO loop are repeated;
 *                      alphaPhic stays the one built before the loop (pimpleFoamYade.C:85); the pFinal solver entry is used in
 *                      the last PISO corrector of the LAST outer corrector only (pimple.finalInnerIter(), pEqn.H:35)
 *   relaxU / relaxUFinal   equations { U; UFinal }: UcEqn.relax(), pimpleFoamYade/UcEqn.H:13 (OpenFOAM-6 fvMatrix::relax(alpha))
 *   relaxP / relaxPFinal   fields { p; pFinal }:    p.relax(),     pimpleFoamYade/pEqn.H:41  (p = prevIter + alpha (p - prevIter))
 * A factor <= 0 means "no entry in fvSolution" (the call is a no-op, the default); the ...Final factor applies on the last
 * outer corrector when given.  As in OpenFOAM, prevIter fields exist only when nOuterCorrectors > 1: a p factor < 1 with
 * one outer corrector is refused by fy_pimple_solve.  nOuterCorrectors x nCorrectors <= 8 (fy_ico_stats slots).          */
int fy_set_pimple_controls(fy_handle h, int nOuterCorrectors, double relaxU, double relaxUFinal, double relaxP, double relaxPFinal);
/* transportProperties nu of the fluid solve (fy_set_properties sets it too) */
int fy_set_viscosity(fy_handle h, double nu);
/* phi = linearInterpolate(U) & Sf   (createPhi.H, icoFoamYade/createFields.H:152) from the U on the device */
int fy_create_phi(fy_handle h);
/* CourantNo.H + vGrad = fvc::grad(U)   (icoFoamYade.C:68-71): everything before setParticleAction */
int fy_ico_pre(fy_handle h, double dt);
/* pimpleFoamYade's pre-coupling block (pimpleFoamYade/pimpleFoamYade.C:71-76): CourantNo.H, then on the device
 * fields  FY_F_DDTU  = fvc::ddt(Uc) + fvc::div(phic, Uc)     FY_F_GRADP = fvc::grad(p)
 *         FY_F_DIVT  = 2 nu fvc::laplacian(alphac, Uc)       FY_F_VGRAD = fvc::grad(Uc)
 * from U (FY_F_U), p (FY_F_P), phi (FY_F_PHI) and the void fraction field as it stands (FY_F_ALPHA: setSourceZero reset
 * it to 1 at the end of the previous step, as in the reference's loop; its patches hold 1.0) -- the inputs the Gaussian branch of setParticleAction reads, so that with the
 * fluid state resident no host field crosses PCIe.  The Euler fvc::ddt term is identically zero at that point of
 * the time step (Uc.oldTime() has just been stored from Uc).                                                    */
int fy_pimple_pre(fy_handle h, double dt);
/* pimpleFoamYade's fluid step after setParticleAction (pimpleFoamYade/pimpleFoamYade.C:82-104 with UcEqn.H, pEqn.H,
 * continuityErrs.H; laminar; nOuterCorrectors / relaxation: fy_set_pimple_controls): alphacf = interpolate(alphac), alphaPhic = alphacf*phic,
 *   UcEqn = ddt(alphac,Uc) + div(alphaPhic,Uc) - Sp(ddt(alphac) + div(alphaPhic),Uc) + divDevRhoReff(Uc) == Sp(uSourceDrag,Uc)
 *   phicForces = flux(rAUc*uSource) + rAUcf*(g & Sf);  momentum predictor == reconstruct(phicForces/rAUcf - snGrad(p)*magSf)
 *   PISO: phiHbyA = flux(HbyA) + alphacf*rAUcf*ddtCorr + phicForces;  laplacian(alphacf*rAUcf, p) == ddt(alphac) + div(alphacf*phiHbyA)
 *         phic = phiHbyA - flux/alphacf;  Uc = HbyA + rAUc*reconstruct((phicForces - flux/alphacf)/rAUcf)
 * Reads the device fields the coupling pass left (FY_F_ALPHA, FY_F_USOURCE, FY_F_USOURCEDRAG), updates U, p, phi;
 * g = gravitational acceleration (NULL = 0).  fy_piso_controls / fy_get_ico_stats serve this solver too.  Patch types
 * as for icoFoam plus fixedFluxPressure on p (constrainPressure, pEqn.H:21: closed boxes under gravity).           */
int fy_pimple_solve(fy_handle h, double dt, const double g[3]);
/* UEqn assembly with uSource, momentum predictor, PISO correctors (icoFoamYade.C:79-140).  Reads and
 * updates the device fields U (FY_F_U), p (FY_F_P), phi (FY_F_PHI); reads uSource (FY_F_USOURCE).    */
int fy_ico_solve(fy_handle h, double dt);
int fy_get_ico_stats(fy_handle h, fy_ico_stats* out);

/* Parity hooks: single operators on host data, in OpenFOAM's own layouts (cell fields [N][..], face
 * fields and matrix coefficients in LDU face order: internal faces then boundary faces patch by patch). */
int fy_fvc_grad_vector(fy_handle h, const double* h_U, double* h_out9);      /* fvc::grad(U), icoFoamYade.C:71  */
int fy_fvc_grad_scalar(fy_handle h, const double* h_p, double* h_out3);      /* fvc::grad(p), icoFoamYade.C:136 */
int fy_fvc_div_flux(fy_handle h, const double* h_phi, double* h_out);        /* fvc::div(phi), icoFoamYade.C:120 */
int fy_fvc_div_phi_vector(fy_handle h, const double* h_phi, const double* h_U, double* h_out3);   /* fvc::div(phic, Uc), pimpleFoamYade.C:73 */
/* fvc::laplacian(gamma, U), pimpleFoamYade.C:75; gammaB = value of gamma on the patches (alphac: 1.0) */
int fy_fvc_laplacian_gamma_vector(fy_handle h, const double* h_gamma, double gammaB, const double* h_U, double* h_out3);
/* lduMatrix solves over the mesh's addressing: out3 = initialResidual, finalResidual, nIterations */
int fy_pcg_solve(fy_handle h, const double* h_diag, const double* h_upper, const double* h_source, double* h_psi,
                 double tol, double relTol, int maxIter, int preconditioner, double out3[3]);
int fy_smooth_solve(fy_handle h, const double* h_diag, const double* h_lower, const double* h_upper,
                    const double* h_source, double* h_psi, double tol, double relTol, int maxIter, double out3[3]);
int fy_dic_precondition(fy_handle h, const double* h_diag, const double* h_upper, const double* h_rA, double* h_wA);
/* intermediates of the last PISO corrector: "rAU" [N], "HbyA" [N][3], "phiHbyA" [faces], "gradP" [N][3];
 * after fy_pimple_solve also "phicForces" [faces], "divDev" [N][3]                                         */
int fy_fv_get(fy_handle h, const char* name, double* h_dst);
/* Device time (ms) of the phases of the last fy_ico_solve: [0] UEqn assembly + momentum predictor
 * [1] pressure solves (PCG) [2] the rest of the correctors; and the last PCG's iteration time.      */
int fy_get_fluid_ms(fy_handle h, double out[4]);
/* Average device time (ms, CUDA events on the handle's stream) of each kernel class of the PCG iteration,
 * sampled while profiling is on (fy_set_profiling): [0] preconditioner forward sweep [1] backward sweep
 * (+ wA.rA) [2] search direction [3] Amul (+ wA.pA) [4] solution/residual update (+ sum|rA|);
 * [5] samples taken [6] PCG iterations run since the last reset.                                        */
int fy_get_kernel_ms(fy_handle h, double out[8], int reset);

/* ---------------------------------------------------------------------------------------------
 * multi-GPU: decomposition of the pressure solve over the ranks of one job (one process per GPU)
 *
 * The reference runs its fluid side decomposed (`mpiexec ... -n 2 icoFoamYade -parallel`, README.md:29; the bbox routing
 * of FoamYade.C:77-155 exists for that); OpenFOAM's Pstream does the halo exchange and the global sums of the linear
 * solvers.  Here the ranks form a Py x Pz grid over the hex box (decomposePar `simple`, n (1 Py Pz)): rank rz*Py + ry
 * owns rows [32 jbLo, 32 jbHi) of the k-planes [kLo, kHi) in the PCG solve of the pressure equation (the y cuts fall on
 * multiples of 32 rows, the unit of the solver's memory layout).  Per iteration: one halo exchange of the search
 * direction (ncclSend / ncclRecv of the region's boundary rows) and three 1-double all-reduces; the DIC preconditioner
 * is rank-local, which is what OpenFOAM's decomposed runs do (DICPreconditioner sees a processor's own lduMatrix).  The
 * FV assembly kernels around the solve run on the whole box on every rank, so the solution is gathered at the end of
 * each solve.
 *   fy_dist_unique_id : rank 0 creates the NCCL id (FY_DIST_ID_BYTES bytes) and hands it to the others (any transport)
 *   fy_dist_init      : every rank, with the same id; call once, after fy_create and before the first solve.  The grid
 *                       is chosen by the library: Py = the largest divisor of nranks that is <= ceil(ny / 32)  (cutting
 *                       in y shortens the critical path of the wavefront preconditioner sweeps, cutting in z does not;
 *                       DESIGN.md section 6), Pz = nranks / Py.  Environment FY_DIST_PY overrides Py.
 *   fy_dist_init_grid : the same with Py given (py = 1: z slabs only)
 *   fy_dist_info      : out[0] rank [1] nranks [2] kLo [3] kHi [4] collectives issued [5] halo bytes sent
 *   fy_dist_grid      : out[0] Py [1] Pz [2] ry [3] rz [4] jLo [5] jHi (rows of y) [6] kLo [7] kHi
 *                       [8] 1: the iteration's collectives run inside its kernels over NVLink peer memory (CUDA IPC
 *                       mappings made at fy_dist_init; environment FY_DIST_PEER=0 or an IPC failure: 0, NCCL calls) [9] 0
 * ------------------------------------------------------------------------------------------- */
#define FY_DIST_ID_BYTES 128
int fy_dist_unique_id(char id[FY_DIST_ID_BYTES]);
int fy_dist_init(fy_handle h, int rank, int nranks, const char id[FY_DIST_ID_BYTES]);
int fy_dist_init_grid(fy_handle h, int rank, int nranks, int py, const char id[FY_DIST_ID_BYTES]);
int fy_dist_info(fy_handle h, long long out[6]);
int fy_dist_grid(fy_handle h, long long out[10]);

/* Blocks until all work queued on the handle's stream is complete. */
int fy_synchronize(fy_handle h);

/* Per-phase device times (ms) of the last coupling_proc, measured with CUDA events on the handle's
 * stream: [0] h2d [1] locate [2] weights+accumulate [3] void fraction [4] forces [5] d2h.
 * Only recorded when profiling is on (fy_set_profiling(h, 1)); synchronises.                        */
int fy_set_profiling(fy_handle h, int on);
int fy_get_phase_ms(fy_handle h, double out[8]);
/* Stopwatch on the handle's own stream (CUDA events): bench.py brackets its timed region with these,
 * because events recorded on another stream would not see this handle's kernels.  stop synchronises. */
int fy_timer_start(fy_handle h);
int fy_timer_stop(fy_handle h, double* ms);
/* Kernel launches issued by this handle since creation (for bench.py's gpu_launches). */
long long fy_launch_count(fy_handle h);

#ifdef __cplusplus
}
#endif
#endif /* FYCUDA_H */
