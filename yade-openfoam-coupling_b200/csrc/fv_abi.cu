// fv_abi.cu -- extern "C" entry points of the fluid half (include/fycuda.h, "the fluid half" section).
#include <cstring>
#include <string>

#include "fv_solver.h"

namespace {
int needFv(fy_handle h, FvState** out)
{
    if (!h) return FY_ERR_INVALID;
    FvState* s = h->fv;
    if (!s || !s->supported) {
        h->err = "finite-volume path unavailable on this mesh: " + (s ? s->why : std::string("not initialised"));
        return FY_ERR_UNSUPPORTED;
    }
    *out = s;
    return FY_OK;
}

// device staging area for the parity hooks (host arrays in OpenFOAM layouts)
int stage(fy_ctx* h, FvState* s, size_t doubles, double** p)
{
    if (doubles > s->stageCap) {
        if (s->stage) cudaFree(s->stage);
        s->stage = nullptr;
        s->stageCap = 0;
        FY_CUDA(cudaMalloc((void**)&s->stage, doubles * sizeof(double)));
        s->stageCap = doubles;
    }
    *p = s->stage;
    return FY_OK;
}
int h2d(fy_ctx* h, double* d, const double* src, size_t n)
{
    FY_CUDA(cudaMemcpyAsync(d, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return FY_OK;
}
int d2h(fy_ctx* h, double* dst, const double* d, size_t n)
{
    FY_CUDA(cudaMemcpyAsync(dst, d, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}
}  // namespace

extern "C" {

int fy_fv_supported(fy_handle h)
{
    FyDeviceGuard guard_(h);
    if (!h || !h->fv) return 0;
    if (!h->fv->supported) h->err = "finite-volume path unavailable on this mesh: " + h->fv->why;
    return h->fv->supported ? 1 : 0;
}

int fy_piso_default_controls(fy_piso_controls* c)
{
    if (!c) return FY_ERR_INVALID;
    c->nCorrectors = 2; c->nNonOrthogonalCorrectors = 0; c->momentumPredictor = 1; c->pRefCell = 0; c->pRefValue = 0.0;
    c->pTol = 1e-6; c->pRelTol = 0.05; c->pFinalTol = 1e-6; c->pFinalRelTol = 0.0; c->UTol = 1e-5; c->URelTol = 0.0;
    c->maxIter = 1000; c->preconditioner = FY_PRECOND_DIC;
    return FY_OK;
}

int fy_set_piso_controls(fy_handle h, const fy_piso_controls* c)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!c || c->nCorrectors < 1 || c->nNonOrthogonalCorrectors < 0 || c->pRefCell < 0 || c->pRefCell >= s->g.N ||
        c->preconditioner < 0 || c->preconditioner > 2 || c->maxIter < 0) {
        h->err = "fy_set_piso_controls: bad controls";
        return FY_ERR_INVALID;
    }
    s->ctl = *c;
    return FY_OK;
}

int fy_set_pimple_controls(fy_handle h, int nOuterCorrectors, double relaxU, double relaxUFinal, double relaxP, double relaxPFinal)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (nOuterCorrectors < 1 || relaxU > 1 || relaxUFinal > 1 || relaxP > 1 || relaxPFinal > 1) {
        h->err = "fy_set_pimple_controls: nOuterCorrectors >= 1 and relaxation factors <= 1 (<= 0: none)";
        return FY_ERR_INVALID;
    }
    s->nOuter = nOuterCorrectors;
    s->relaxU = relaxU; s->relaxUFinal = relaxUFinal; s->relaxP = relaxP; s->relaxPFinal = relaxPFinal;
    return FY_OK;
}

int fy_set_viscosity(fy_handle h, double nu)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    s->nu = nu;
    return FY_OK;
}

int fy_create_phi(fy_handle h)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    return fvCreatePhi(h, s);
}

int fy_ico_pre(fy_handle h, double dt)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    return fvIcoPre(h, s, dt);
}

int fy_pimple_pre(fy_handle h, double dt)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!(dt > 0)) { h->err = "fy_pimple_pre: dt must be positive"; return FY_ERR_INVALID; }
    return fvPimplePre(h, s, dt);
}

int fy_pimple_solve(fy_handle h, double dt, const double g[3])
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!(dt > 0)) { h->err = "fy_pimple_solve: dt must be positive"; return FY_ERR_INVALID; }
    const double zero[3] = {0, 0, 0};
    return fvPimpleSolve(h, s, dt, g ? g : zero);
}

int fy_ico_solve(fy_handle h, double dt)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!(dt > 0)) { h->err = "fy_ico_solve: dt must be positive"; return FY_ERR_INVALID; }
    return fvIcoSolve(h, s, dt);
}

int fy_get_ico_stats(fy_handle h, fy_ico_stats* out)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!out) return FY_ERR_INVALID;
    *out = s->stats;
    return FY_OK;
}

int fy_fvc_grad_vector(fy_handle h, const double* U, double* out9)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!U || !out9) return FY_ERR_INVALID;
    const size_t N = (size_t)s->g.N;
    double* d;
    if ((rc = stage(h, s, 12 * N, &d))) return rc;
    if ((rc = h2d(h, d, U, 3 * N))) return rc;
    if ((rc = fvGradVector(h, s, d, d + 3 * N))) return rc;
    return d2h(h, out9, d + 3 * N, 9 * N);
}

int fy_fvc_grad_scalar(fy_handle h, const double* p, double* out3)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!p || !out3) return FY_ERR_INVALID;
    const size_t N = (size_t)s->g.N;
    double* d;
    if ((rc = stage(h, s, 4 * N, &d))) return rc;
    if ((rc = h2d(h, d, p, N))) return rc;
    if ((rc = fvGradScalar(h, s, d, d + N))) return rc;
    return d2h(h, out3, d + N, 3 * N);
}

int fy_fvc_div_flux(fy_handle h, const double* phi, double* out)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!phi || !out) return FY_ERR_INVALID;
    const size_t N = (size_t)s->g.N, nF = (size_t)s->nFi + s->nB, NS = (size_t)s->g.nSlots;
    double* d;
    if ((rc = stage(h, s, nF + NS + N, &d))) return rc;
    if ((rc = h2d(h, d, phi, nF))) return rc;
    if ((rc = fvFacesToSlots(h, s, (int)nF, d, d + nF))) return rc;
    if ((rc = fvDivFlux(h, s, d + nF, d + nF + NS))) return rc;
    return d2h(h, out, d + nF + NS, N);
}

int fy_fvc_div_phi_vector(fy_handle h, const double* phi, const double* U, double* out3)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!phi || !U || !out3) return FY_ERR_INVALID;
    const size_t N = (size_t)s->g.N, nF = (size_t)s->nFi + s->nB, NS = (size_t)s->g.nSlots;
    double* d;
    if ((rc = stage(h, s, nF + NS + 6 * N, &d))) return rc;
    double *dSlots = d + nF, *dU = dSlots + NS, *dOut = dU + 3 * N;
    if ((rc = h2d(h, d, phi, nF))) return rc;
    if ((rc = h2d(h, dU, U, 3 * N))) return rc;
    if ((rc = fvFacesToSlots(h, s, (int)nF, d, dSlots))) return rc;
    if ((rc = fvDivPhiVector(h, s, dSlots, dU, dOut))) return rc;
    return d2h(h, out3, dOut, 3 * N);
}

int fy_fvc_laplacian_gamma_vector(fy_handle h, const double* gamma, double gammaB, const double* U, double* out3)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!gamma || !U || !out3) return FY_ERR_INVALID;
    const size_t N = (size_t)s->g.N;
    double* d;
    if ((rc = stage(h, s, 7 * N, &d))) return rc;
    if ((rc = h2d(h, d, gamma, N))) return rc;
    if ((rc = h2d(h, d + N, U, 3 * N))) return rc;
    if ((rc = fvLaplacianGammaVector(h, s, 1.0, d, gammaB, d + N, d + 4 * N))) return rc;
    return d2h(h, out3, d + 4 * N, 3 * N);
}

// matrix coefficients arrive in LDU face order; the kernels want owner slots
static int stageMatrix(fy_ctx* h, FvState* s, const double* diag, const double* lower, const double* upper,
                       const double* source, const double* psi, double** dDiag, double** dLo, double** dUp, double** dB,
                       double** dPsi)
{
    const size_t N = (size_t)s->g.N, Fi = (size_t)s->nFi, N3 = 3 * N;
    double* d;
    int rc;
    if ((rc = stage(h, s, 3 * N + 2 * N3 + Fi, &d))) return rc;
    *dDiag = d; *dB = d + N; *dPsi = d + 2 * N; *dUp = d + 3 * N; *dLo = d + 3 * N + N3;
    double* tmp = d + 3 * N + 2 * N3;
    if ((rc = h2d(h, *dDiag, diag, N))) return rc;
    if (source && (rc = h2d(h, *dB, source, N))) return rc;
    if (psi && (rc = h2d(h, *dPsi, psi, N))) return rc;
    FY_CUDA(cudaMemsetAsync(*dUp, 0, 2 * N3 * sizeof(double), h->stream));
    if ((rc = h2d(h, tmp, upper, Fi))) return rc;
    if ((rc = fvFacesToSlots(h, s, (int)Fi, tmp, *dUp))) return rc;
    if (lower) {
        if ((rc = h2d(h, tmp, lower, Fi))) return rc;
        if ((rc = fvFacesToSlots(h, s, (int)Fi, tmp, *dLo))) return rc;
    } else {
        *dLo = *dUp;
    }
    return FY_OK;
}

int fy_pcg_solve(fy_handle h, const double* diag, const double* upper, const double* source, double* psi, double tol,
                 double relTol, int maxIter, int preconditioner, double out3[3])
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!diag || !upper || !source || !psi || preconditioner < 0 || preconditioner > 2) return FY_ERR_INVALID;
    double *dD, *dLo, *dUp, *dB, *dPsi;
    if ((rc = stageMatrix(h, s, diag, nullptr, upper, source, psi, &dD, &dLo, &dUp, &dB, &dPsi))) return rc;
    fy_solver_perf perf{0, 0, 0, 0};
    if ((rc = fvPcgSolve(h, s, dD, dUp, dB, dPsi, tol, relTol, maxIter, preconditioner, &perf))) return rc;
    if (out3) { out3[0] = perf.initialResidual; out3[1] = perf.finalResidual; out3[2] = perf.nIterations; }
    return d2h(h, psi, dPsi, (size_t)s->g.N);
}

int fy_smooth_solve(fy_handle h, const double* diag, const double* lower, const double* upper, const double* source,
                    double* psi, double tol, double relTol, int maxIter, double out3[3])
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!diag || !lower || !upper || !source || !psi) return FY_ERR_INVALID;
    double *dD, *dLo, *dUp, *dB, *dPsi;
    if ((rc = stageMatrix(h, s, diag, lower, upper, source, psi, &dD, &dLo, &dUp, &dB, &dPsi))) return rc;
    fy_solver_perf perf{0, 0, 0, 0};
    if ((rc = fvSmoothSetMatrix(h, s, dLo, dUp))) return rc;
    if ((rc = fvSmoothSolve(h, s, dD, dB, dPsi, tol, relTol, maxIter, &perf))) return rc;
    if (out3) { out3[0] = perf.initialResidual; out3[1] = perf.finalResidual; out3[2] = perf.nIterations; }
    return d2h(h, psi, dPsi, (size_t)s->g.N);
}

int fy_dic_precondition(fy_handle h, const double* diag, const double* upper, const double* rA, double* wA)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!diag || !upper || !rA || !wA) return FY_ERR_INVALID;
    double *dD, *dLo, *dUp, *dB, *dPsi;
    if ((rc = stageMatrix(h, s, diag, nullptr, upper, rA, nullptr, &dD, &dLo, &dUp, &dB, &dPsi))) return rc;
    if ((rc = fvDicPrecondition(h, s, dD, dUp, dB, dPsi))) return rc;
    return d2h(h, wA, dPsi, (size_t)s->g.N);
}

int fy_fv_get(fy_handle h, const char* name, double* dst)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!name || !dst) return FY_ERR_INVALID;
    const std::string k(name);
    const size_t N = (size_t)s->g.N;
    if (k == "rAU") return d2h(h, dst, s->rAU, N);
    if (k == "HbyA") return d2h(h, dst, s->HbyA, 3 * N);
    if (k == "gradP") return d2h(h, dst, s->gradP, 3 * N);
    if (k == "diagU") return d2h(h, dst, s->diagU, N);
    if (k == "sourceU") return d2h(h, dst, s->srcU, 3 * N);
    if (k == "divDev" && s->divDev) return d2h(h, dst, s->divDev, 3 * N);
    if (k == "phicForces" && s->phicForces) {
        const size_t nF = (size_t)s->nFi + s->nB;
        double* d;
        if ((rc = stage(h, s, nF, &d))) return rc;
        if ((rc = fvSlotsToFaces(h, s, (int)nF, s->phicForces, d))) return rc;
        return d2h(h, dst, d, nF);
    }
    if (k == "bGradP") {             // snGrad(p) constrainPressure left on the fixedFluxPressure faces, [faces] (0 elsewhere)
        const size_t nF = (size_t)s->nFi + s->nB;
        if (!s->bGradP) { h->err = "fy_fv_get: the mesh has no fixedFluxPressure patch"; return FY_ERR_INVALID; }
        double* d;
        if ((rc = stage(h, s, nF, &d))) return rc;
        if ((rc = fvSlotsToFaces(h, s, (int)nF, s->bGradP, d))) return rc;
        return d2h(h, dst, d, nF);
    }
    if (k == "phiHbyA" || k == "phi" || k == "upperP" || k == "upperU" || k == "lowerU") {
        const size_t nF = (k == "phiHbyA" || k == "phi") ? (size_t)s->nFi + s->nB : (size_t)s->nFi;
        double* d;
        if ((rc = stage(h, s, nF, &d))) return rc;
        const double* src = k == "phiHbyA" ? s->phiHbyA : (k == "phi" ? s->phi : (k == "upperP" ? s->upP : (k == "upperU" ? s->upU : s->loU)));
        if ((rc = fvSlotsToFaces(h, s, (int)nF, src, d))) return rc;
        return d2h(h, dst, d, nF);
    }
    if (k == "pencilTraceRaw") {     // debug: the same buffer, raw 64-bit values as doubles (section counters of a PEN2_TIMING build)
        const size_t n = ((size_t)s->pen.g.nJB * (s->pen.g.nz + 16 * 8 + 64) * 4 + 64) * 32;
        std::vector<unsigned long long> t(n);
        FY_CUDA(cudaStreamSynchronize(h->stream));
        FY_CUDA(cudaMemcpy(t.data(), s->pen.trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (size_t q = 0; q < n; ++q) dst[q] = (double)t[q];
        return FY_OK;
    }
    if (k == "pencilTrace") {        // debug: [nJB*nz][4] time stamps (ns, relative to the earliest) of the last pencil launch
        const size_t n = ((size_t)s->pen.g.nJB * (s->pen.g.nz + 16 * 8 + 64) * 4 + 64) * 32;
        std::vector<unsigned long long> t(n);
        FY_CUDA(cudaStreamSynchronize(h->stream));
        FY_CUDA(cudaMemcpy(t.data(), s->pen.trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (size_t q = 0; q < n; ++q) if (t[q] && t[q] < t0) t0 = t[q];
        for (size_t q = 0; q < n; ++q) dst[q] = t[q] ? (double)(t[q] - t0) : -1.0;
        return FY_OK;
    }
    h->err = "fy_fv_get: unknown field " + k;
    return FY_ERR_INVALID;
}

int fy_get_fluid_ms(fy_handle h, double out[4])
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    for (int i = 0; i < 4; ++i) out[i] = s->fluidMs[i];
    return FY_OK;
}

int fy_get_kernel_ms(fy_handle h, double out[8], int reset)
{
    FyDeviceGuard guard_(h);
    FvState* s;
    int rc = needFv(h, &s);
    if (rc) return rc;
    if (!out) return FY_ERR_INVALID;
    const double n = s->kernelSamples > 0 ? (double)s->kernelSamples : 1.0;
    for (int q = 0; q < 5; ++q) out[q] = s->kernelMs[q] / n;
    out[5] = (double)s->kernelSamples;
    out[6] = (double)s->pcgIterations;
    out[7] = s->pen.fusedTail && (!s->pen.dist || s->pen.peer) ? 1.0 : 0.0;      // 1: out[2] is the fused direction + Amul + update kernel, out[3] = out[4] = 0
    if (reset) {
        for (int q = 0; q < 5; ++q) s->kernelMs[q] = 0;
        s->kernelSamples = 0;
        s->pcgIterations = 0;
    }
    return FY_OK;
}

int fy_dist_unique_id(char id[FY_DIST_ID_BYTES])
{
    std::string err;
    return fvDistUniqueId(id, err);
}

int fy_dist_init(fy_handle h, int rank, int nranks, const char id[FY_DIST_ID_BYTES])
{
    FyDeviceGuard guard_(h);
    if (!h || !id) return FY_ERR_INVALID;
    if (!h->fv || !h->fv->supported) { h->err = "fy_dist_init: the mesh did not qualify for the device FV path"; return FY_ERR_UNSUPPORTED; }
    return fvDistInit(h, h->fv, rank, nranks, 0, id);
}

int fy_dist_init_grid(fy_handle h, int rank, int nranks, int py, const char id[FY_DIST_ID_BYTES])
{
    FyDeviceGuard guard_(h);
    if (!h || !id) return FY_ERR_INVALID;
    if (!h->fv || !h->fv->supported) { h->err = "fy_dist_init_grid: the mesh did not qualify for the device FV path"; return FY_ERR_UNSUPPORTED; }
    return fvDistInit(h, h->fv, rank, nranks, py, id);
}

int fy_dist_grid(fy_handle h, long long out[10])
{
    if (!h || !out || !h->fv) return FY_ERR_INVALID;
    const PenState& P = h->fv->pen;
    out[0] = P.Py; out[1] = P.Pz; out[2] = P.ry; out[3] = P.rz;
    out[4] = P.gl.jbLo * 32; out[5] = std::min(P.gl.jbHi * 32, P.g.ny); out[6] = P.gl.kLo; out[7] = P.gl.kHi;
    out[8] = P.peer ? 1 : 0; out[9] = 0;
    return FY_OK;
}

int fy_dist_info(fy_handle h, long long out[6])
{
    if (!h || !out || !h->fv) return FY_ERR_INVALID;
    const PenState& P = h->fv->pen;
    out[0] = P.rank; out[1] = P.nranks; out[2] = P.kLo; out[3] = P.kHi; out[4] = P.distCollectives; out[5] = P.distHaloBytes;
    return FY_OK;
}

}  // extern "C"
