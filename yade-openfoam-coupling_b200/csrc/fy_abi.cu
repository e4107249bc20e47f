// fy_abi.cu -- the extern "C" entry points of include/fycuda.h (life cycle, fields, coupling operator).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "fy_ctx.h"
#include "fv_solver.h"

namespace {
std::string g_createErr;

size_t fieldWidth(int f)
{
    switch (f) {
        case FY_F_VGRAD: return 9;
        case FY_F_USOURCEDRAG:
        case FY_F_ALPHA:
        case FY_F_P: return 1;
        case FY_F_PHI: return 0;     // face field, sized separately
        default: return 3;
    }
}
size_t fieldCount(const fy_ctx* h, int f)
{
    if (f == FY_F_PHI) return (size_t)h->nFaces + (size_t)h->nBFaces;
    return fieldWidth(f) * (size_t)h->nCells;
}

template <class T>
int upload(fy_ctx* h, T** d, const T* src, size_t n)
{
    *d = nullptr;
    if (n == 0) return FY_OK;
    FY_CUDA(cudaMalloc((void**)d, n * sizeof(T)));
    if (src) FY_CUDA(cudaMemcpyAsync(*d, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    else FY_CUDA(cudaMemsetAsync(*d, 0, n * sizeof(T), h->stream));
    return FY_OK;
}

int uploadMesh(fy_ctx* h, const fy_mesh_desc* m)
{
    int rc;
    const size_t N = (size_t)m->nCells, Fi = (size_t)m->nInternalFaces;
    if ((rc = upload(h, &h->dC, m->C, 3 * N))) return rc;
    if ((rc = upload(h, &h->dV, m->V, N))) return rc;
    h->V0 = m->V[0];
    if (Fi > 0) {
        if (!m->owner || !m->neighbour || !m->Sf || !m->magSf || !m->weights || !m->deltaCoeffs) {
            h->err = "mesh: internal-face arrays missing";
            return FY_ERR_INVALID;
        }
        for (size_t f = 0; f < Fi; ++f) {
            if (m->owner[f] < 0 || m->neighbour[f] >= m->nCells || m->owner[f] >= m->neighbour[f] ||
                (f > 0 && (m->owner[f] < m->owner[f - 1] ||
                           (m->owner[f] == m->owner[f - 1] && m->neighbour[f] <= m->neighbour[f - 1])))) {
                h->err = "mesh: internal faces must be in upper-triangular order (owner < neighbour, sorted)";
                return FY_ERR_INVALID;
            }
        }
        h->hOwner.assign(m->owner, m->owner + Fi);
        h->hNeigh.assign(m->neighbour, m->neighbour + Fi);
        if ((rc = upload(h, &h->dOwner, m->owner, Fi))) return rc;
        if ((rc = upload(h, &h->dNeigh, m->neighbour, Fi))) return rc;
        if ((rc = upload(h, &h->dSf, m->Sf, 3 * Fi))) return rc;
        if ((rc = upload(h, &h->dMagSf, m->magSf, Fi))) return rc;
        if ((rc = upload(h, &h->dWeights, m->weights, Fi))) return rc;
        if ((rc = upload(h, &h->dDeltaCoeffs, m->deltaCoeffs, Fi))) return rc;
        // owner-start and losort addressing (lduAddressing::ownerStartAddr / losortAddr)
        std::vector<int> ownStart(N + 1, 0), losortStart(N + 1, 0), losort(Fi);
        for (size_t f = 0; f < Fi; ++f) { ownStart[m->owner[f] + 1]++; losortStart[m->neighbour[f] + 1]++; }
        for (size_t c = 0; c < N; ++c) { ownStart[c + 1] += ownStart[c]; losortStart[c + 1] += losortStart[c]; }
        {
            std::vector<int> pos(losortStart.begin(), losortStart.end() - 1);
            for (size_t f = 0; f < Fi; ++f) losort[pos[m->neighbour[f]]++] = (int)f;   // ascending face order per cell
        }
        if ((rc = upload(h, &h->dOwnStart, ownStart.data(), N + 1))) return rc;
        if ((rc = upload(h, &h->dLosortStart, losortStart.data(), N + 1))) return rc;
        if ((rc = upload(h, &h->dLosort, losort.data(), Fi))) return rc;
        FY_CUDA(cudaStreamSynchronize(h->stream));      // host vectors above go out of scope
    }
    // boundary
    size_t nB = 0;
    for (int p = 0; p < m->nPatches; ++p) nB += (size_t)m->patches[p].nFaces;
    h->nBFaces = (int)nB;
    h->patches.clear();
    if (nB > 0) {
        std::vector<int> fc(nB), pid(nB);
        std::vector<double> sf(3 * nB), msf(nB), dc(nB);
        size_t o = 0;
        for (int p = 0; p < m->nPatches; ++p) {
            const fy_patch_desc& pd = m->patches[p];
            FyPatch fp;
            fp.nFaces = pd.nFaces; fp.start = (int)o; fp.bcU = pd.bcU; fp.bcP = pd.bcP;
            fp.valueU[0] = pd.valueU[0]; fp.valueU[1] = pd.valueU[1]; fp.valueU[2] = pd.valueU[2];
            fp.valueP = pd.valueP;
            h->patches.push_back(fp);
            if (pd.nFaces < 0 || (pd.nFaces > 0 && (!pd.faceCells || !pd.Sf || !pd.magSf || !pd.deltaCoeffs))) {
                h->err = "fy_create: patch arrays missing";
                return FY_ERR_INVALID;
            }
            for (int i = 0; i < pd.nFaces; ++i, ++o) {
                if (pd.faceCells[i] < 0 || (size_t)pd.faceCells[i] >= N) {
                    h->err = "fy_create: patch faceCells entry outside [0, nCells)";
                    return FY_ERR_INVALID;
                }
                fc[o] = pd.faceCells[i]; pid[o] = p;
                sf[3 * o] = pd.Sf[3 * (size_t)i]; sf[3 * o + 1] = pd.Sf[3 * (size_t)i + 1]; sf[3 * o + 2] = pd.Sf[3 * (size_t)i + 2];
                msf[o] = pd.magSf[i]; dc[o] = pd.deltaCoeffs[i];
            }
        }
        std::vector<int> bStart(N + 1, 0), bOrder(nB);
        for (size_t b = 0; b < nB; ++b) bStart[fc[b] + 1]++;
        for (size_t c = 0; c < N; ++c) bStart[c + 1] += bStart[c];
        {
            std::vector<int> pos(bStart.begin(), bStart.end() - 1);
            for (size_t b = 0; b < nB; ++b) bOrder[pos[fc[b]]++] = (int)b;
        }
        if ((rc = upload(h, &h->dBFaceCells, fc.data(), nB))) return rc;
        if ((rc = upload(h, &h->dBPatch, pid.data(), nB))) return rc;
        if ((rc = upload(h, &h->dBSf, sf.data(), 3 * nB))) return rc;
        if ((rc = upload(h, &h->dBMagSf, msf.data(), nB))) return rc;
        if ((rc = upload(h, &h->dBDeltaCoeffs, dc.data(), nB))) return rc;
        if ((rc = upload(h, &h->dBStart, bStart.data(), N + 1))) return rc;
        if ((rc = upload(h, &h->dBOrder, bOrder.data(), nB))) return rc;
        FY_CUDA(cudaStreamSynchronize(h->stream));
    }
    return FY_OK;
}
}  // namespace

extern "C" {

int fy_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* fy_version(void) { return "fycuda 0.1 (sm_100a)"; }

const char* fy_last_error(fy_handle h) { return h ? h->err.c_str() : g_createErr.c_str(); }

int fy_create(const fy_mesh_desc* m, int device, fy_handle* out)
{
    if (!m || !out || m->nCells <= 0 || !m->C || !m->V) { g_createErr = "fy_create: bad mesh descriptor"; return FY_ERR_INVALID; }
    *out = nullptr;
    if (fy_device_count() <= device) {
        g_createErr = "fy_create: no CUDA device " + std::to_string(device) + " (this engine has no CPU path)";
        return FY_ERR_NO_DEVICE;
    }
    fy_ctx* h = new fy_ctx();
    h->device = device;
    auto fail = [&](int rc) { g_createErr = h->err; fy_destroy(h); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return fail(FY_ERR_CUDA); }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "stream create failed"; return fail(FY_ERR_CUDA); }
    for (auto& e : h->ev) cudaEventCreate(&e);
    h->nCells = m->nCells;
    h->nFaces = m->nInternalFaces;
    for (int i = 0; i < 3; ++i) h->boxN[i] = m->boxN[i];
    for (int i = 0; i < 6; ++i) { h->boxGeom[i] = m->boxGeom[i]; h->bbox[i] = m->bbox[i]; }
    int rc = uploadMesh(h, m);
    if (rc) return fail(rc);

    // k-d tree over the cell centres (meshTree.C:9-17), host build, implicit layout
    {
        std::vector<FyKdNode> tree;
        fyBuildKdTree(m->C, m->nCells, tree);
        h->nTree = m->nCells;
        if (cudaMalloc((void**)&h->dTree, tree.size() * sizeof(FyKdNode)) != cudaSuccess) { h->err = "cudaMalloc tree"; return fail(FY_ERR_ALLOC); }
        if (cudaMemcpy(h->dTree, tree.data(), tree.size() * sizeof(FyKdNode), cudaMemcpyHostToDevice) != cudaSuccess) { h->err = "tree upload"; return fail(FY_ERR_CUDA); }
    }
    // hex box: the centre coordinates per axis, taken from mesh.C() itself and checked to BE a tensor product bit for bit
    // (the full-support Gaussian mode evaluates distances from them; any other mesh leaves dAxis null)
    if (h->boxN[0] > 0 && (long long)h->boxN[0] * h->boxN[1] * h->boxN[2] == m->nCells) {
        const int nx = h->boxN[0], ny = h->boxN[1], nz = h->boxN[2];
        std::vector<double> ax((size_t)nx + ny + nz);
        for (int i = 0; i < nx; ++i) ax[i] = m->C[3 * (size_t)i];
        for (int j = 0; j < ny; ++j) ax[nx + j] = m->C[3 * ((size_t)j * nx) + 1];
        for (int k = 0; k < nz; ++k) ax[nx + ny + k] = m->C[3 * ((size_t)k * nx * ny) + 2];
        bool ok = true;
        for (int k = 0; k < nz && ok; ++k)
            for (int j = 0; j < ny && ok; ++j)
                for (int i = 0; i < nx; ++i) {
                    const double* c = m->C + 3 * ((size_t)i + (size_t)nx * (j + (size_t)ny * k));
                    if (c[0] != ax[i] || c[1] != ax[nx + j] || c[2] != ax[nx + ny + k]) { ok = false; break; }
                }
        if (ok) {
            if (cudaMalloc((void**)&h->dAxis, ax.size() * sizeof(double)) != cudaSuccess) { h->err = "cudaMalloc axis"; return fail(FY_ERR_ALLOC); }
            if (cudaMemcpy(h->dAxis, ax.data(), ax.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { h->err = "axis upload"; return fail(FY_ERR_CUDA); }
        }
    }
    // initFields constants, host arithmetic exactly as FoamYade.C:69-72 and meshTree.C:155
    h->interpRange = 4 * std::pow(h->V0, 1.0 / 3.0);
    h->sigmaInterp = h->interpRange * 0.42460;
    h->interpRangeCu = std::pow(h->interpRange, 3.0);
    h->sigmaPi = 1.0 / (std::pow(2 * M_PI * h->sigmaInterp * h->sigmaInterp, 1.5));
    {
        const double range = h->interpRange;
        h->maxDist = (range * range) + (0.25 * range * range);
    }
    // fields
    const size_t N = (size_t)h->nCells;
    for (int f = 0; f < FY_F_COUNT; ++f) {
        const size_t cnt = fieldCount(h, f);
        if (cnt == 0) continue;
        if (cudaMalloc((void**)&h->dField[f], cnt * sizeof(double)) != cudaSuccess) { h->err = "cudaMalloc field"; return fail(FY_ERR_ALLOC); }
        cudaMemsetAsync(h->dField[f], 0, cnt * sizeof(double), h->stream);
    }
    if (cudaMalloc((void**)&h->dPvol, N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&h->dUpAcc, 3 * N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&h->dStamp, N * sizeof(int)) != cudaSuccess) { h->err = "cudaMalloc accumulators"; return fail(FY_ERR_ALLOC); }
    cudaMemsetAsync(h->dPvol, 0, N * sizeof(double), h->stream);
    cudaMemsetAsync(h->dUpAcc, 0, 3 * N * sizeof(double), h->stream);
    cudaMemsetAsync(h->dStamp, 0, N * sizeof(int), h->stream);
    rc = fyInitCouplingFields(h);
    if (rc) return fail(rc);
    rc = fvCreate(h, m);                 // decides whether the mesh qualifies for the device FV path
    if (rc) return fail(rc);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "sync after create failed"; return fail(FY_ERR_CUDA); }
    *out = h;
    return FY_OK;
}

int fy_destroy(fy_handle h)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    fvDestroy(h);
    void* ptrs[] = {h->dC, h->dV, h->dOwner, h->dNeigh, h->dOwnStart, h->dLosort, h->dLosortStart, h->dSf, h->dMagSf,
                    h->dWeights, h->dDeltaCoeffs, h->dBFaceCells, h->dBPatch, h->dBSf, h->dBMagSf, h->dBDeltaCoeffs,
                    h->dBStart, h->dBOrder, h->dTree, h->dPvol, h->dUpAcc, h->dStamp, h->dPdata.p, h->dFound.p,
                    h->dForce.p, h->dIds.p, h->dCnt.p, h->dW.p, h->dCell.p, h->dKey.p, h->dKey2.p, h->dIdx.p, h->dPerm.p,
                    h->dSortTmp.p, h->dListCnt.p, h->dListIds.p, h->dListW.p, h->dAxis, h->dAllWt.p};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& f : h->dField) if (f) cudaFree(f);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->copyStream) {
        cudaStreamSynchronize(h->copyStream);
        cudaEventDestroy(h->evUp); cudaEventDestroy(h->evProc); cudaEventDestroy(h->evDown);
        cudaStreamDestroy(h->copyStream);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return FY_OK;
}

int fy_set_properties(fy_handle h, double rhoP, double rhoF, double nu, int gaussianInterp)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    h->rhoP = rhoP; h->rhoF = rhoF; h->nu = nu;
    h->gaussian = gaussianInterp != 0;
    h->propsSet = true;
    if (h->fv) h->fv->nu = nu;
    return FY_OK;
}

int fy_get_constants(fy_handle h, double out4[4])
{
    FyDeviceGuard guard_(h);
    if (!h || !out4) return FY_ERR_INVALID;
    out4[0] = h->interpRange; out4[1] = h->sigmaInterp; out4[2] = h->interpRangeCu; out4[3] = h->sigmaPi;
    return FY_OK;
}

int fy_host_alloc(void** p, size_t bytes)
{
    if (!p) return FY_ERR_INVALID;
    if (cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); *p = nullptr; return FY_ERR_ALLOC; }
    return FY_OK;
}
int fy_host_free(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return FY_ERR_CUDA; }
    return FY_OK;
}

int fy_bind_host_fields(fy_handle h, const double* U, const double* gradP, const double* vGrad, const double* divT,
                        const double* ddtU, double* uSourceDrag, double* alpha, double* uSource, double* uParticle)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    h->hIn[0] = U; h->hIn[1] = gradP; h->hIn[2] = vGrad; h->hIn[3] = divT; h->hIn[4] = ddtU;
    h->hOut[0] = uSourceDrag; h->hOut[1] = alpha; h->hOut[2] = uSource; h->hOut[3] = uParticle;
    return FY_OK;
}

int fy_upload_field(fy_handle h, int f, const double* src)
{
    FyDeviceGuard guard_(h);
    if (!h || f < 0 || f >= FY_F_COUNT || !src || !h->dField[f]) return FY_ERR_INVALID;
    FY_CUDA(cudaMemcpyAsync(h->dField[f], src, fieldCount(h, f) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (f == FY_F_PHI && h->fv && h->fv->supported) {      // OpenFOAM face order -> owner slots
        int rc = fvFacesToSlots(h, h->fv, h->fv->nFi + h->fv->nB, h->dField[f], h->fv->phi);
        if (rc) return rc;
    }
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}
int fy_download_field(fy_handle h, int f, double* dst)
{
    FyDeviceGuard guard_(h);
    if (!h || f < 0 || f >= FY_F_COUNT || !dst || !h->dField[f]) return FY_ERR_INVALID;
    if (f == FY_F_PHI && h->fv && h->fv->supported) {
        int rc = fvSlotsToFaces(h, h->fv, h->fv->nFi + h->fv->nB, h->fv->phi, h->dField[f]);
        if (rc) return rc;
    }
    FY_CUDA(cudaMemcpyAsync(dst, h->dField[f], fieldCount(h, f) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}
int fy_device_field(fy_handle h, int f, double** d)
{
    FyDeviceGuard guard_(h);
    if (!h || f < 0 || f >= FY_F_COUNT || !d) return FY_ERR_INVALID;
    *d = h->dField[f];
    return FY_OK;
}

int fy_locate(fy_handle h, const double* xyz, int n, int* ids, int* counts)
{
    FyDeviceGuard guard_(h);
    if (!h || n < 0 || (n > 0 && (!xyz || !ids || !counts))) return FY_ERR_INVALID;
    if (n == 0) return FY_OK;
    int rc;
    // scratch of its own (the staging buffers of fy_get_last_lists): the particle buffer and the cell lists of the last
    // fy_coupling_proc stay intact
    if ((rc = fyReserve(h, h->dListW, (size_t)n * FY_MAXLIST))) return rc;
    if ((rc = fyReserve(h, h->dListIds, (size_t)n * FY_MAXLIST))) return rc;
    if ((rc = fyReserve(h, h->dListCnt, (size_t)n))) return rc;
    FY_CUDA(cudaMemcpyAsync(h->dListW.p, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if ((rc = fyLaunchLocate(h, h->dListW.p, 3, n, h->dListIds.p, h->dListCnt.p))) return rc;
    FY_CUDA(cudaMemcpyAsync(ids, h->dListIds.p, (size_t)n * FY_MAXLIST * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaMemcpyAsync(counts, h->dListCnt.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}

int fy_find_cell(fy_handle h, const double* xyz, int n, int* cell)
{
    FyDeviceGuard guard_(h);
    if (!h || n < 0 || (n > 0 && (!xyz || !cell))) return FY_ERR_INVALID;
    if (h->boxN[0] <= 0) { h->err = "fy_find_cell: mesh has no hex-box descriptor"; return FY_ERR_UNSUPPORTED; }
    if (n == 0) return FY_OK;
    int rc;
    if ((rc = fyReserve(h, h->dPdata, (size_t)n * 10))) return rc;
    if ((rc = fyReserve(h, h->dCell, (size_t)n))) return rc;
    FY_CUDA(cudaMemcpyAsync(h->dPdata.p, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if ((rc = fyLaunchFindCell(h, h->dPdata.p, 3, n, h->dCell.p))) return rc;
    FY_CUDA(cudaMemcpyAsync(cell, h->dCell.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}

int fy_coupling_begin(fy_handle h, double dt)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    if (!h->propsSet) { h->err = "fy_set_properties must be called first"; return FY_ERR_INVALID; }
    h->deltaT = dt;
    // host-bound inputs: what this branch reads (Gaussian: U gradP divT; point-force: U vGrad)
    static const int inId[5] = {FY_F_U, FY_F_GRADP, FY_F_VGRAD, FY_F_DIVT, FY_F_DDTU};
    for (int i = 0; i < 5; ++i) {
        if (!h->hIn[i]) continue;
        const bool needed = (i == 0) || (h->gaussian ? (i == 1 || i == 3) : (i == 2));
        if (!needed) continue;
        FY_CUDA(cudaMemcpyAsync(h->dField[inId[i]], h->hIn[i], fieldCount(h, inId[i]) * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    }
    return FY_OK;
}

int fy_coupling_proc_device(fy_handle h, const double* d_pdata, int n, int* d_found, double* d_force)
{
    FyDeviceGuard guard_(h);
    if (!h || n < 0) return FY_ERR_INVALID;
    return fyCouplingProcDevice(h, d_pdata, n, d_found, d_force);
}

int fy_coupling_pass_device(fy_handle h, int pass, const double* d_pdata, int n, int* d_found, double* d_force)
{
    FyDeviceGuard guard_(h);
    if (!h || n < 0 || pass < 0 || pass > 2) return FY_ERR_INVALID;
    return fyCouplingPass(h, pass, d_pdata, n, d_found, d_force);
}

int fy_device_accumulators(fy_handle h, double** d_pvol, double** d_upAcc, int** d_stamp)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    if (d_pvol) *d_pvol = h->dPvol;
    if (d_upAcc) *d_upAcc = h->dUpAcc;
    if (d_stamp) *d_stamp = h->dStamp;
    return FY_OK;
}

int fy_stream(fy_handle h, void** cuda_stream)
{
    FyDeviceGuard guard_(h);
    if (!h || !cuda_stream) return FY_ERR_INVALID;
    *cuda_stream = (void*)h->stream;
    return FY_OK;
}

int fy_coupling_proc(fy_handle h, const double* pdata, int n, int* found, double* force)
{
    FyDeviceGuard guard_(h);
    if (!h || n < 0 || (n > 0 && (!pdata || !found || !force))) return FY_ERR_INVALID;
    if (n == 0) { h->lastN = 0; return FY_OK; }
    int rc;
    if ((rc = fyReserve(h, h->dPdata, (size_t)n * 10))) return rc;
    if ((rc = fyReserve(h, h->dFound, (size_t)n))) return rc;
    if ((rc = fyReserve(h, h->dForce, (size_t)n * 6))) return rc;
    if (h->profiling) cudaEventRecord(h->ev[0], h->stream);
    FY_CUDA(cudaMemcpyAsync(h->dPdata.p, pdata, (size_t)n * 10 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if ((rc = fyCouplingProcDevice(h, h->dPdata.p, n, h->dFound.p, h->dForce.p))) return rc;
    FY_CUDA(cudaMemcpyAsync(found, h->dFound.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaMemcpyAsync(force, h->dForce.p, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (h->profiling) cudaEventRecord(h->ev[5], h->stream);
    FY_CUDA(cudaStreamSynchronize(h->stream));
    if (h->profiling) {
        for (int i = 0; i < 5; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
            h->phaseMs[i] = ms;
        }
    }
    return FY_OK;
}

// ---- overlapped wire transfers (second stream + events): the records come up while the solver's pre-coupling block
// runs, the forces go down while the pressure-velocity solve runs
static int copyStream(fy_ctx* h)
{
    if (h->copyStream) return FY_OK;
    FY_CUDA(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
    FY_CUDA(cudaEventCreateWithFlags(&h->evUp, cudaEventDisableTiming));
    FY_CUDA(cudaEventCreateWithFlags(&h->evProc, cudaEventDisableTiming));
    FY_CUDA(cudaEventCreateWithFlags(&h->evDown, cudaEventDisableTiming));
    return FY_OK;
}

int fy_particles_upload_async(fy_handle h, const double* pdata, int n)
{
    FyDeviceGuard guard_(h);
    if (!h || n < 0 || (n > 0 && !pdata)) return FY_ERR_INVALID;
    int rc;
    if ((rc = copyStream(h))) return rc;
    h->stagedN = n;
    if (n == 0) return FY_OK;
    if ((rc = fyReserve(h, h->dPdata, (size_t)n * 10))) return rc;
    if ((rc = fyReserve(h, h->dFound, (size_t)n))) return rc;
    if ((rc = fyReserve(h, h->dForce, (size_t)n * 6))) return rc;
    // (the previous step's kernels that read dPdata are long done: fy_results_wait has returned)
    FY_CUDA(cudaMemcpyAsync(h->dPdata.p, pdata, (size_t)n * 10 * sizeof(double), cudaMemcpyHostToDevice, h->copyStream));
    FY_CUDA(cudaEventRecord(h->evUp, h->copyStream));
    return FY_OK;
}

int fy_coupling_proc_staged(fy_handle h, int* found, double* force)
{
    FyDeviceGuard guard_(h);
    if (!h || !h->copyStream) return FY_ERR_INVALID;
    const int n = h->stagedN;
    if (n > 0 && (!found || !force)) return FY_ERR_INVALID;
    if (n == 0) { h->lastN = 0; return FY_OK; }
    int rc;
    FY_CUDA(cudaStreamWaitEvent(h->stream, h->evUp, 0));
    if ((rc = fyCouplingProcDevice(h, h->dPdata.p, n, h->dFound.p, h->dForce.p))) return rc;
    FY_CUDA(cudaEventRecord(h->evProc, h->stream));
    FY_CUDA(cudaStreamWaitEvent(h->copyStream, h->evProc, 0));
    FY_CUDA(cudaMemcpyAsync(found, h->dFound.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->copyStream));
    FY_CUDA(cudaMemcpyAsync(force, h->dForce.p, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->copyStream));
    FY_CUDA(cudaEventRecord(h->evDown, h->copyStream));
    return FY_OK;
}

int fy_results_wait(fy_handle h)
{
    FyDeviceGuard guard_(h);
    if (!h || !h->copyStream) return FY_ERR_INVALID;
    FY_CUDA(cudaEventSynchronize(h->evDown));
    return FY_OK;
}

int fy_coupling_end(fy_handle h)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    static const int outId[4] = {FY_F_USOURCEDRAG, FY_F_ALPHA, FY_F_USOURCE, FY_F_UPARTICLE};
    bool any = false;
    for (int i = 0; i < 4; ++i) {
        if (!h->hOut[i]) continue;
        if (!h->gaussian && i != 2) continue;      // point-force writes uSource only
        FY_CUDA(cudaMemcpyAsync(h->hOut[i], h->dField[outId[i]], fieldCount(h, outId[i]) * sizeof(double),
                                cudaMemcpyDeviceToHost, h->stream));
        any = true;
    }
    if (any) FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}

int fy_set_particle_action(fy_handle h, double dt, const double* pdata, int n, int* found, double* force)
{
    FyDeviceGuard guard_(h);
    int rc;
    if ((rc = fy_coupling_begin(h, dt))) return rc;
    if ((rc = fy_coupling_proc(h, pdata, n, found, force))) return rc;
    return fy_coupling_end(h);
}

int fy_set_source_zero(fy_handle h)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    int rc = fySourceZeroDevice(h);
    if (rc) return rc;
    // host-bound outputs: the reference zeroes the solver's own arrays (F.C:557-563)
    const size_t N = (size_t)h->nCells;
    if (h->hOut[2]) std::memset(h->hOut[2], 0, 3 * N * sizeof(double));
    if (h->gaussian) {
        if (h->hOut[0]) std::memset(h->hOut[0], 0, N * sizeof(double));
        if (h->hOut[1]) std::fill(h->hOut[1], h->hOut[1] + N, 1.0);
        if (h->hOut[3]) std::memset(h->hOut[3], 0, 3 * N * sizeof(double));
    }
    return FY_OK;
}

int fy_set_gaussian_options(fy_handle h, int support, int addedMass, int torque)
{
    FyDeviceGuard guard_(h);
    if (!h || (support != FY_SUPPORT_TRAIL && support != FY_SUPPORT_FULL)) return FY_ERR_INVALID;
    if (support == FY_SUPPORT_FULL && !h->dAxis) {
        h->err = "fy_set_gaussian_options: the full-support mode needs a hex-box mesh (fy_mesh_desc.boxN) whose cell centres are a tensor product";
        return FY_ERR_UNSUPPORTED;
    }
    h->supportFull = support == FY_SUPPORT_FULL;
    h->addedMass = addedMass != 0;
    h->gaussTorque = torque != 0;
    return FY_OK;
}

int fy_get_last_lists(fy_handle h, int n, int* counts, int* ids, double* weights)
{
    FyDeviceGuard guard_(h);
    if (!h || n < 0 || n > h->lastN) return FY_ERR_INVALID;
    if (n == 0) return FY_OK;
    if (!h->gaussian) { h->err = "fy_get_last_lists: Gaussian mode only"; return FY_ERR_INVALID; }
    if (h->supportFull) {
        // full-support mode: the cell sets are never materialised -- the counts are all there is
        if (ids || weights) { h->err = "fy_get_last_lists: the full-support mode keeps no cell lists (counts only)"; return FY_ERR_UNSUPPORTED; }
        int rc0;
        if ((rc0 = fyReserve(h, h->dListCnt, (size_t)h->lastN))) return rc0;
        if ((rc0 = fyUnpermuteCounts(h, h->lastN, h->dListCnt.p))) return rc0;
        if (counts) FY_CUDA(cudaMemcpyAsync(counts, h->dListCnt.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        FY_CUDA(cudaStreamSynchronize(h->stream));
        return FY_OK;
    }
    // the lists live in sorted order, structure-of-arrays: back to wire order first (all lastN of them)
    const int m = h->lastN;
    int rc;
    if ((rc = fyReserve(h, h->dListCnt, (size_t)m))) return rc;
    if ((rc = fyReserve(h, h->dListIds, (size_t)m * FY_MAXLIST))) return rc;
    if ((rc = fyReserve(h, h->dListW, (size_t)m * FY_MAXLIST))) return rc;
    if ((rc = fyUnpermuteLists(h, m, h->dListCnt.p, h->dListIds.p, h->dListW.p))) return rc;
    if (counts) FY_CUDA(cudaMemcpyAsync(counts, h->dListCnt.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (ids) FY_CUDA(cudaMemcpyAsync(ids, h->dListIds.p, (size_t)n * FY_MAXLIST * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (weights) FY_CUDA(cudaMemcpyAsync(weights, h->dListW.p, (size_t)n * FY_MAXLIST * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}

int fy_synchronize(fy_handle h)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    FY_CUDA(cudaStreamSynchronize(h->stream));
    return FY_OK;
}

int fy_set_profiling(fy_handle h, int on)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    h->profiling = on != 0;
    return FY_OK;
}
int fy_get_phase_ms(fy_handle h, double out[8])
{
    FyDeviceGuard guard_(h);
    if (!h || !out) return FY_ERR_INVALID;
    for (int i = 0; i < 8; ++i) out[i] = h->phaseMs[i];
    return FY_OK;
}
int fy_timer_start(fy_handle h)
{
    FyDeviceGuard guard_(h);
    if (!h) return FY_ERR_INVALID;
    FY_CUDA(cudaEventRecord(h->ev[6], h->stream));
    return FY_OK;
}
int fy_timer_stop(fy_handle h, double* ms)
{
    FyDeviceGuard guard_(h);
    if (!h || !ms) return FY_ERR_INVALID;
    FY_CUDA(cudaEventRecord(h->ev[7], h->stream));
    FY_CUDA(cudaEventSynchronize(h->ev[7]));
    float f = 0;
    FY_CUDA(cudaEventElapsedTime(&f, h->ev[6], h->ev[7]));
    *ms = f;
    return FY_OK;
}
long long fy_launch_count(fy_handle h) { return h ? h->launches : 0; }

}  // extern "C"
