// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference coupling operator
//   /root/reference/FoamYade/FoamYade.{H,C}  and  FoamYade/meshtree/meshTree.{H,C}
// (compiled where they lie, see oracle/Makefile) through a plain C API that the
// tests and bench.py's cpu_baseline leg load with ctypes as
// oracle/_ref/libfoamyade_ref.so.
//
// It supplies what the reference expects from its environment:
//   * the dozen MPI calls of FoamYade.C (single process; this file plays the
//     Yade peer: world ranks 0..Y-1; the Foam side is one rank, world rank Y),
//     logging every call so the wire sequence can be asserted;
//   * an fvMesh (cell centres / volumes / vertices handed in by the caller and a
//     "containing cell" query for an axis-aligned box of hex cells);
//   * the ten solver-owned fields FoamYade binds (F.H:106-122).
//
// Two ways of running a coupling step:
//   ref_step         -> FoamYade::setParticleAction(dt), untouched (F.C:605-632)
//   ref_step_pieces  -> the same sequence issued through the class's public
//                       methods, with (a) the canonical <=12-cell truncation of
//                       cellIds between locateAllParticles() and the weights
//                       (meshTree.H:66-68 reads one past the container when it
//                       holds 12 entries, so longer lists depend on stale heap
//                       bytes) and (b) optionally a dense per-cell accumulate in
//                       the same particle-then-cell visiting order in place of
//                       the quadratic list scan of buildCellPartList
//                       (F.C:265-289) -- same additions in the same order, so
//                       the same bits, but O(pairs).
#ifdef FY_HOST_CLASS
// second build of this harness (oracle/_build/libhost_harness.so): the SAME fake Yade peer and fields drive
// the product's OpenFOAM-side host class (yade-openfoam-coupling_b200/host/FoamYadeB200.H) instead of the
// reference, so a GPU test can compare wire traces and replies message for message.
#include "icoFoamYadeB200.H"
typedef Foam::FoamYadeB200 FyClass;
#else
#include "FoamYade.H"
typedef Foam::FoamYade FyClass;
#endif
#include "PstreamGlobals.H"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>
#include <sstream>

namespace Foam
{
InfoStream Info;
namespace PstreamGlobals { MPI_Comm MPI_COMM_FOAM = MPI_COMM_FOAM_SHIM; }
}

// ----------------------------------------------------------------------------
// fake Yade peer
// ----------------------------------------------------------------------------
namespace
{
struct Peer
{
    int nYade = 1;                  // Y: 1 => serial-Yade mode (F.C:31)
    // per-step inputs
    const double* pdata = nullptr;  // [n][10]
    int n = 0;
    std::vector<int> split;         // parallel: particle range per worker rank r=1..Y-1: [split[r-1], split[r])
    double yadeDT = 0.0;
    // per-step captures
    std::vector<int> ownerRank;     // serial: result of the per-particle Allreduce(MAX)
    std::vector<double> sumForce;   // serial Gaussian: the 6n Allreduce(SUM) inputs in call order
    std::vector<double> p2pForce;   // serial point-force: 6 doubles per Send(tag 1005)
    std::vector<std::vector<int> > foundMsg;     // parallel: tag 1004 per worker
    std::vector<std::vector<double> > forceMsg;  // parallel: tag 1005 per worker
    std::vector<double> bbox;       // tag 1001 payloads
    double fluidDT = 0.0;
    bool logging = false;
    std::vector<std::string> trace;
    long nBcast = 0, nAllreduce = 0, nSend = 0, nRecv = 0, nIsend = 0;
    void clearStep()
    {
        ownerRank.clear(); sumForce.clear(); p2pForce.clear();
        foundMsg.assign(nYade, std::vector<int>());
        forceMsg.assign(nYade, std::vector<double>());
        nBcast = nAllreduce = nSend = nRecv = nIsend = 0;
    }
    void log(const char* what, const char* ty, int cnt, int peer, int tag, const char* comm)
    {
        if (!logging) return;
        char b[160];
        std::snprintf(b, sizeof b, "%s %s[%d] peer=%d tag=%d %s", what, ty, cnt, peer, tag, comm);
        trace.push_back(b);
    }
};
Peer g_peer;
const char* tyName(MPI_Datatype t) { return t == MPI_DOUBLE ? "f64" : "i32"; }
const char* commName(MPI_Comm c) { return c == MPI_COMM_WORLD ? "WORLD" : "FOAM"; }
}

extern "C" {
int MPI_Comm_rank(MPI_Comm c, int* r) { *r = (c == MPI_COMM_WORLD) ? g_peer.nYade : 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { *s = (c == MPI_COMM_WORLD) ? g_peer.nYade + 1 : 1; return 0; }

int MPI_Isend(const void* buf, int cnt, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* rq)
{
    g_peer.nIsend++;
    g_peer.log("Isend", tyName(t), cnt, dst, tag, commName(c));
    if (tag == 1001) { const double* d = (const double*)buf; g_peer.bbox.insert(g_peer.bbox.end(), d, d + cnt); }
    *rq = 0;
    return 0;
}
int MPI_Wait(MPI_Request*, MPI_Status*) { g_peer.log("Wait", "-", 0, -1, -1, "-"); return 0; }

int MPI_Recv(void* buf, int cnt, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status*)
{
    g_peer.nRecv++;
    g_peer.log("Recv", tyName(t), cnt, src, tag, commName(c));
    if (tag == 1003) {          // numParticlesProc[F] from worker src (F.C:124); F == 1 here
        int* ib = (int*)buf;
        for (int i = 0; i < cnt; ++i) ib[i] = g_peer.split[src] - g_peer.split[src - 1];
    } else if (tag == 1002) {   // particle records of worker src (F.C:152)
        std::memcpy(buf, g_peer.pdata + 10*(size_t)g_peer.split[src - 1], sizeof(double)*(size_t)cnt);
    } else if (tag == 1060) {   // Yade dt (F.C:544)
        *(double*)buf = g_peer.yadeDT;
    }
    return 0;
}

int MPI_Send(const void* buf, int cnt, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{
    g_peer.nSend++;
    g_peer.log("Send", tyName(t), cnt, dst, tag, commName(c));
    if (tag == 1004) {
        const int* b = (const int*)buf; g_peer.foundMsg[dst].assign(b, b + cnt);
    } else if (tag == 1005) {
        const double* b = (const double*)buf;
        if (g_peer.nYade == 1) g_peer.p2pForce.insert(g_peer.p2pForce.end(), b, b + cnt);
        else g_peer.forceMsg[dst].assign(b, b + cnt);
    } else if (tag == 1050) {
        g_peer.fluidDT = *(const double*)buf;
    }
    return 0;
}

int MPI_Bcast(void* buf, int cnt, MPI_Datatype t, int root, MPI_Comm c)
{
    g_peer.nBcast++;
    g_peer.log("Bcast", tyName(t), cnt, root, -1, commName(c));
    if (c != MPI_COMM_WORLD) return 0;          // F.C:547: Foam-side rebroadcast, one Foam rank
    if (t == MPI_INT) { *(int*)buf = g_peer.n; }                        // F.C:176
    else if (cnt == 1) { *(double*)buf = g_peer.yadeDT; }               // F.C:549
    else if (cnt > 0) { std::memcpy(buf, g_peer.pdata, sizeof(double)*(size_t)cnt); }   // F.C:181
    return 0;
}

int MPI_Allreduce(const void* in, void* out, int cnt, MPI_Datatype t, MPI_Op, MPI_Comm c)
{
    g_peer.nAllreduce++;
    g_peer.log("Allreduce", tyName(t), cnt, -1, -1, commName(c));
    // (cnt == 1: the reference's per-particle / per-component collectives, F.C:228, 514; cnt > 1: the product's batched
    // wire mode, one message per direction and step)
    if (t == MPI_INT) { for (int i = 0; i < cnt; ++i) { ((int*)out)[i] = ((const int*)in)[i]; g_peer.ownerRank.push_back(((const int*)in)[i]); } }
    else { for (int i = 0; i < cnt; ++i) { ((double*)out)[i] = ((const double*)in)[i]; g_peer.sumForce.push_back(((const double*)in)[i]); } }
    return 0;
}
int MPI_Finalize(void) { return 0; }
}

// ----------------------------------------------------------------------------
// the reference object and its environment
// ----------------------------------------------------------------------------
namespace
{
struct Box { int nx, ny, nz; double x0, y0, z0, hx, hy, hz; };

Foam::label boxFindCell(const void* ctx, const Foam::point& p)
{
    const Box& b = *(const Box*)ctx;
    if (b.nx <= 0) return -1;
    const double fi = std::floor((p.x() - b.x0)/b.hx);
    const double fj = std::floor((p.y() - b.y0)/b.hy);
    const double fk = std::floor((p.z() - b.z0)/b.hz);
    if (!(fi >= 0 && fi < b.nx && fj >= 0 && fj < b.ny && fk >= 0 && fk < b.nz)) return -1;
    return (Foam::label)fi + b.nx*((Foam::label)fj + b.ny*(Foam::label)fk);
}

struct Ref
{
    Box box;
    Foam::fvMesh mesh;
    Foam::volVectorField U, gradP, divT, ddtU, uSource, uParticle;
    Foam::volTensorField vGrad;
    Foam::volScalarField uSourceDrag, alpha, p;
    Foam::uniformDimensionedVectorField g;
    FyClass* fy = nullptr;
    void* fyStorage = nullptr;
    bool gaussian = false;
    // captured cell lists of the last step (after canonical truncation when applied)
    std::vector<int> listCount, listIds;   // [n], [n][16]
    double tLocate = 0, tWeights = 0, tAccum = 0, tForce = 0, tSend = 0;
    // SURVEY 8(f)3 options of the Gaussian branch (ref_set_gaussian_options): full-support cell lists instead of the
    // k-d descent's improvement trail; the forces the class defines but never calls
    bool supportFull = false, addedMass = false, gaussTorque = false;
};

// every cell whose centre lies within the search bound of meshTree::nnearestCellsRange (MT.C:155: d^2 < range^2 +
// 0.25 range^2), ascending by distance (ties by id) -- "range based search" (README.md:5).  Box arithmetic only bounds the
// candidates; the test uses the mesh's own centres in meshTree::distance's operation order.
void fullSupportList(const Ref* r, const Foam::vector& pos, double range, std::vector<int>& ids)
{
    const Box& b = r->box;
    const double maxDist = (range*range) + (0.25*range*range);
    const double R = std::sqrt(maxDist);
    int lo[3], hi[3];
    const double pp[3] = {pos.x(), pos.y(), pos.z()}, x0[3] = {b.x0, b.y0, b.z0}, hh[3] = {b.hx, b.hy, b.hz};
    const int nn[3] = {b.nx, b.ny, b.nz};
    for (int d = 0; d < 3; ++d) {
        lo[d] = std::max(0, (int)std::floor((pp[d] - R - x0[d])/hh[d]) - 1);
        hi[d] = std::min(nn[d] - 1, (int)std::floor((pp[d] + R - x0[d])/hh[d]) + 1);
    }
    std::vector<std::pair<double, int> > hit;
    for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) {
                const int c = i + b.nx*(j + b.ny*k);
                const Foam::vector& C = r->mesh.C()[c];
                double dist = 0;
                const double d0 = C.x() - pos.x(), d1 = C.y() - pos.y(), d2 = C.z() - pos.z();
                dist += d0*d0; dist += d1*d1; dist += d2*d2;                     // MT.C:54-64
                if (dist < maxDist) hit.push_back(std::make_pair(dist, c));
            }
    std::sort(hit.begin(), hit.end());
    ids.clear();
    for (const auto& q : hit) ids.push_back(q.second);
}

double nowSec()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void collectOutputs(Ref* r, int n, int* found, double* force6)
{
    Peer& P = g_peer;
    for (int i = 0; i < n; ++i) { found[i] = -1; for (int j = 0; j < 6; ++j) force6[6*(size_t)i + j] = 0.0; }
    if (P.nYade == 1) {
        for (int i = 0; i < n && i < (int)P.ownerRank.size(); ++i) found[i] = (P.ownerRank[i] > 0) ? 1 : -1;
        if (r->gaussian) {
            for (size_t i = 0; i < P.sumForce.size() && i < 6*(size_t)n; ++i) force6[i] = P.sumForce[i];
        } else if (P.p2pForce.size() == 6*(size_t)n && P.nSend <= 2) {
            // batched wire mode: ONE message with all 6n doubles (zeros for particles this rank did not find)
            for (size_t i = 0; i < 6*(size_t)n; ++i) force6[i] = P.p2pForce[i];
        } else {
            size_t m = 0;     // one 6-double message per found particle, in index order (F.C:519-531)
            for (int i = 0; i < n; ++i) {
                if (found[i] != 1) continue;
                for (int j = 0; j < 6; ++j) force6[6*(size_t)i + j] = P.p2pForce[m + j];
                m += 6;
            }
        }
    } else {
        for (int w = 1; w < P.nYade; ++w) {
            const int lo = P.split[w - 1];
            for (size_t i = 0; i < P.foundMsg[w].size(); ++i) found[lo + (int)i] = P.foundMsg[w][i];
            for (size_t i = 0; i < P.forceMsg[w].size(); ++i) force6[6*(size_t)lo + i] = P.forceMsg[w][i];
        }
    }
}

void setupStep(Ref* r, double yadeDT, const double* pdata, int n, const int* split)
{
    Peer& P = g_peer;
    P.clearStep();
    P.pdata = pdata; P.n = n; P.yadeDT = yadeDT;
    P.split.assign(1, 0);
    if (P.nYade > 1) {
        // worker w in 1..Y-1 owns [split[w-1], split[w]); default: equal contiguous chunks
        for (int w = 1; w < P.nYade; ++w)
            P.split.push_back(split ? split[w] : (int)((long long)n*w/(P.nYade - 1)));
    }
    (void)r;
}
}

extern "C" {

// nYade == 1 => serial-Yade protocol; >= 2 => parallel protocol with nYade-1 workers.
// box: nx,ny,nz,x0,y0,z0,hx,hy,hz used ONLY by findCell (point-force mode); pass nx=0 for "no findCell".
void* ref_create(int nCells, const double* C, const double* V, int nPoints, const double* points,
                 const int* boxN, const double* boxGeom, int gaussian, int nYade)
{
    Ref* r = new Ref();
    r->gaussian = gaussian != 0;
    r->box = Box{boxN[0], boxN[1], boxN[2], boxGeom[0], boxGeom[1], boxGeom[2], boxGeom[3], boxGeom[4], boxGeom[5]};
    r->mesh.C_.setSize(nCells); r->mesh.V_.setSize(nCells); r->mesh.points_.setSize(nPoints);
    for (int i = 0; i < nCells; ++i) {
        r->mesh.C_[i] = Foam::vector(C[3*(size_t)i], C[3*(size_t)i + 1], C[3*(size_t)i + 2]);
        r->mesh.V_[i] = V[i];
    }
    for (int i = 0; i < nPoints; ++i)
        r->mesh.points_[i] = Foam::point(points[3*(size_t)i], points[3*(size_t)i + 1], points[3*(size_t)i + 2]);
    r->mesh.findCellFn_ = boxFindCell; r->mesh.findCellCtx_ = &r->box;
    r->U.setSize(nCells); r->gradP.setSize(nCells); r->divT.setSize(nCells); r->ddtU.setSize(nCells);
    r->uSource.setSize(nCells); r->uParticle.setSize(nCells); r->vGrad.setSize(nCells);
    r->uSourceDrag.setSize(nCells); r->alpha.setSize(nCells); r->p.setSize(nCells);

    const bool keepLogging = g_peer.logging;     // ref_set_logging(1) before ref_create records the ctor's sends
    g_peer = Peer();
    g_peer.logging = keepLogging;
    g_peer.nYade = nYade;
    g_peer.clearStep();
    // `bool serialYade` (F.H:91) is only ever set to true (F.C:31): give the object zeroed storage so
    // that the parallel protocol is what runs when Y > 1.
    r->fyStorage = std::calloc(1, sizeof(FyClass));
    r->fy = new (r->fyStorage) FyClass(r->mesh, r->U, r->gradP, r->vGrad, r->divT, r->ddtU, r->g,
                                               r->uSourceDrag, r->alpha, r->uSource, r->uParticle, gaussian != 0);
#ifdef FY_HOST_CLASS
    if (r->box.nx > 0)
        r->fy->setHexBox(r->box.nx, r->box.ny, r->box.nz, Foam::point(r->box.x0, r->box.y0, r->box.z0),
                         Foam::vector(r->box.hx, r->box.hy, r->box.hz));
#endif
    return r;
}

void ref_destroy(void* h)
{
    Ref* r = (Ref*)h;
    if (!r) return;
    r->fy->~FyClass();
    std::free(r->fyStorage);
    delete r;     // the k-d nodes are leaked by the reference itself (MT.C:28)
}

void ref_set_properties(void* h, double rhoP, double rhoF, double nu) { ((Ref*)h)->fy->setScalarProperties(rhoP, rhoF, nu); }

// name: U gradP divT ddtU uSource uParticle (3/cell) vGrad (9/cell) uSourceDrag alpha (1/cell)
double* ref_field(void* h, const char* name)
{
    Ref* r = (Ref*)h;
    const std::string s(name);
    if (s == "U") return (double*)r->U.data();
    if (s == "gradP") return (double*)r->gradP.data();
    if (s == "divT") return (double*)r->divT.data();
    if (s == "ddtU") return (double*)r->ddtU.data();
    if (s == "uSource") return (double*)r->uSource.data();
    if (s == "uParticle") return (double*)r->uParticle.data();
    if (s == "vGrad") return (double*)r->vGrad.data();
    if (s == "uSourceDrag") return r->uSourceDrag.data();
    if (s == "alpha") return r->alpha.data();
    if (s == "p") return r->p.data();
    return nullptr;
}

// every solver-owned field moves to new storage (as OpenFOAM's `vGrad = fvc::grad(U)` etc. do every time step)
void ref_realloc_fields(void* h)
{
    Ref* r = (Ref*)h;
    r->U.reallocate(); r->gradP.reallocate(); r->divT.reallocate(); r->ddtU.reallocate(); r->vGrad.reallocate();
    r->uSource.reallocate(); r->uParticle.reallocate(); r->uSourceDrag.reallocate(); r->alpha.reallocate();
}

void ref_get_constants(void* h, double* out4)
{
    FyClass* fy = ((Ref*)h)->fy;
    out4[0] = fy->interpRange; out4[1] = fy->sigmaInterp; out4[2] = fy->interpRangeCu; out4[3] = fy->sigmaPi;
}

void ref_set_logging(int on) { g_peer.logging = on != 0; if (!on) g_peer.trace.clear(); }
// copies the call trace, '\n'-separated, into buf; returns the full length
int ref_get_trace(char* buf, int cap)
{
    std::string s;
    for (const auto& l : g_peer.trace) { s += l; s += '\n'; }
    if (buf && cap > 0) { std::strncpy(buf, s.c_str(), (size_t)cap - 1); buf[cap - 1] = 0; }
    return (int)s.size();
}
void ref_clear_trace() { g_peer.trace.clear(); }
void ref_get_counts(long* out5)
{
    out5[0] = g_peer.nBcast; out5[1] = g_peer.nAllreduce; out5[2] = g_peer.nSend; out5[3] = g_peer.nRecv; out5[4] = g_peer.nIsend;
}
void ref_get_dt(double* out2, void* h) { out2[0] = g_peer.fluidDT; out2[1] = ((Ref*)h)->fy->yadeDT; }
int ref_get_bbox(double* out, int cap)
{
    int m = (int)g_peer.bbox.size();
    for (int i = 0; i < m && i < cap; ++i) out[i] = g_peer.bbox[i];
    return m;
}

#ifndef FY_HOST_CLASS
// raw k-d "range" query of the reference (MT.C:148-179); ids is [n][stride], counts [n] (count may exceed 12)
void ref_locate(void* h, const double* xyz, int n, int* ids, int* counts, int stride)
{
    Ref* r = (Ref*)h;
    for (int i = 0; i < n; ++i) {
        Foam::vector p(xyz[3*(size_t)i], xyz[3*(size_t)i + 1], xyz[3*(size_t)i + 2]);
        std::vector<int> l = r->fy->mshTree.nnearestCellsRange(p, r->fy->interpRange, true);
        counts[i] = (int)l.size();
        for (int j = 0; j < stride; ++j) ids[(size_t)i*stride + j] = (j < (int)l.size()) ? l[j] : -1;
    }
}

#endif

int ref_find_cell(void* h, const double* xyz)
{
    Ref* r = (Ref*)h;
    return r->mesh.findCell(Foam::point(xyz[0], xyz[1], xyz[2]));
}

#ifdef FY_HOST_CLASS
// ---- the fluid step on the device, driven through the host class and the C++ loop bodies of icoFoamYadeB200.H.
// The harness fills the shim fvMesh's LDU addressing / face geometry / patches and the patch types of U and p, which
// the host class reads through the OpenFOAM accessors (owner(), Sf().boundaryField()[patchi], U.boundaryField()[patchi].type()).
void ref_set_fv_mesh(void* h, int nF, const int* owner, const int* neigh, const double* Sf, const double* magSf,
                     const double* w, const double* dc, int nPatches, const int* sizes, const int* faceCells,
                     const double* bSf, const double* bMagSf, const double* bDc, const int* bcU, const double* valU,
                     const int* bcP, const double* valP)
{
    Ref* r = (Ref*)h;
    Foam::fvMesh& m = r->mesh;
    m.owner_.setSize(nF); m.neighbour_.setSize(nF); m.Sf_.setSize(nF); m.magSf_.setSize(nF); m.weights_.setSize(nF);
    m.deltaCoeffs_.setSize(nF);
    for (int f = 0; f < nF; ++f) {
        m.owner_[f] = owner[f]; m.neighbour_[f] = neigh[f];
        m.Sf_[f] = Foam::vector(Sf[3*(size_t)f], Sf[3*(size_t)f + 1], Sf[3*(size_t)f + 2]);
        m.magSf_[f] = magSf[f]; m.weights_[f] = w[f]; m.deltaCoeffs_[f] = dc[f];
    }
    static const char* typeOf[] = {"fixedValue", "zeroGradient", "empty", "fixedFluxPressure"};
    m.boundary_.assign(nPatches, Foam::fvPatch());
    m.Sf_.bf_.resize(nPatches); m.magSf_.bf_.resize(nPatches); m.deltaCoeffs_.bf_.resize(nPatches);
    r->U.bf_.resize(nPatches); r->p.bf_.resize(nPatches);
    size_t o = 0;
    for (int pI = 0; pI < nPatches; ++pI) {
        const int n = sizes[pI];
        m.boundary_[pI].name_ = "patch" + std::to_string(pI);
        m.boundary_[pI].faceCells_.setSize(n);
        r->U.bf_[pI].type_ = typeOf[bcU[pI]];
        r->p.bf_[pI].type_ = typeOf[bcP[pI]];
        for (int q = 0; q < n; ++q, ++o) {
            m.boundary_[pI].faceCells_[q] = faceCells[o];
            m.Sf_.bf_[pI].v_.push_back(Foam::vector(bSf[3*o], bSf[3*o + 1], bSf[3*o + 2]));
            m.magSf_.bf_[pI].v_.push_back(bMagSf[o]);
            m.deltaCoeffs_.bf_[pI].v_.push_back(bDc[o]);
            r->U.bf_[pI].v_.push_back(Foam::vector(valU[3*pI], valU[3*pI + 1], valU[3*pI + 2]));
            r->p.bf_[pI].v_.push_back(valP[pI]);
        }
    }
    r->fy->enableFluidSolve(r->U, r->p);
}

// F1: one message per direction and step instead of 7n collectives (needs the matching Yade-side change)
void ref_set_batched_wire(void* h, int on) { ((Ref*)h)->fy->batchedWire = on != 0; }

// PISO controls of the device fluid step (defaults: the stock cavity set)
void ref_set_piso(void* h, int nCorrectors, int nNonOrth, int momentumPredictor)
{
    Ref* r = (Ref*)h;
    fy_piso_controls c;
    fy_piso_default_controls(&c);
    c.nCorrectors = nCorrectors; c.nNonOrthogonalCorrectors = nNonOrth; c.momentumPredictor = momentumPredictor;
    fy_set_piso_controls(r->fy->engine(), &c);
}

// one pass of the solver's loop body (solver 0: icoFoamYade.C:65-149, 1: pimpleFoamYade.C:65-110) with the fake Yade
// peer on the wire; out: pressure-solve iteration counts (up to 8), their number, the Courant number
void ref_fluid_step(void* h, int solver, double dt, double yadeDT, const double* pdata, int n, const int* split,
                    const double* g3, int* found, double* force6, double* out10)
{
    Ref* r = (Ref*)h;
    setupStep(r, yadeDT, pdata, n, split);
    Foam::FluidStepLog lg = solver == 0 ? Foam::icoFoamYadeTimeStep(*r->fy, dt)
                                        : Foam::pimpleFoamYadeTimeStep(*r->fy, dt, Foam::vector(g3[0], g3[1], g3[2]));
    collectOutputs(r, n, found, force6);
    for (int q = 0; q < 8; ++q) out10[q] = lg.stats.p[q].nIterations;
    out10[8] = lg.stats.nPSolves;
    out10[9] = lg.stats.CoNum;
}

void ref_download_fluid(void* h) { ((Ref*)h)->fy->downloadFluid(); }

// the host class's setters for the options beyond the reference's loop (fycuda.h: fy_set_gaussian_options, fy_set_pimple_controls)
void ref_host_gaussian_options(void* h, int full, int addedMass, int torque)
{
    ((Ref*)h)->fy->setGaussianOptions(full != 0, addedMass != 0, torque != 0);
}
void ref_host_pimple_controls(void* h, int nOuter, double relaxU, double relaxUFinal, double relaxP, double relaxPFinal)
{
    ((Ref*)h)->fy->setPimpleControls(nOuter, relaxU, relaxUFinal, relaxP, relaxPFinal);
}
#endif

// the unmodified driver, F.C:605-632
void ref_step(void* h, double dt, double yadeDT, const double* pdata, int n, const int* split, int* found, double* force6)
{
    Ref* r = (Ref*)h;
    setupStep(r, yadeDT, pdata, n, split);
    r->fy->setParticleAction(dt);
    collectOutputs(r, n, found, force6);
}

#ifndef FY_HOST_CLASS
// Gaussian-branch options, used by ref_step_pieces: supportFull = cell lists hold every cell within the
// search bound (fullSupportList) and are fed to the reference's OWN calcInterpWeightGaussian / hydroDragForce /
// archimedesForce; addedMass = addedMassForce (F.C:392-413) after archimedesForce for every found particle; torque =
// calcHydroTorque's Gaussian branch (F.C:467-478, commented out at F.C:618) after the forces.
void ref_set_gaussian_options(void* h, int supportFull, int addedMass, int torque)
{
    Ref* r = (Ref*)h;
    r->supportFull = supportFull != 0; r->addedMass = addedMass != 0; r->gaussTorque = torque != 0;
}

// same sequence through the public pieces (see header comment). truncate12: canonical list form;
// dense: order-preserving dense accumulate instead of the quadratic scan. Records cell lists and phase times.
void ref_step_pieces(void* h, double dt, double yadeDT, const double* pdata, int n, const int* split,
                     int truncate12, int dense, int* found, double* force6)
{
    Ref* r = (Ref*)h;
    Foam::FoamYade* fy = r->fy;
    setupStep(r, yadeDT, pdata, n, split);
    fy->deltaT = dt;
    if (!fy->serialYade) fy->recvYadeIntrs();
    double t0 = nowSec();
    fy->locateAllParticles();
    r->tLocate = nowSec() - t0;

    r->listCount.assign(n, 0);
    r->listIds.assign((size_t)n*16, -1);
    std::vector<double> pvol, upx, upy, upz;
    std::vector<char> touched;
    r->tWeights = r->tAccum = r->tForce = 0;
    for (const auto& yProc : fy->inCommProcs) {
        int base = 0;
        if (!fy->serialYade) base = g_peer.split[yProc->yRank - 1];
        for (auto& prt : yProc->foundParticles) {
            if (r->gaussian && r->supportFull) fullSupportList(r, prt->pos, fy->interpRange, prt->cellIds);
            else if (truncate12 && prt->cellIds.size() > 12) prt->cellIds.resize(12);
            const int gi = base + prt->indx;
            r->listCount[gi] = (int)prt->cellIds.size();
            for (size_t j = 0; j < prt->cellIds.size() && j < 16; ++j) r->listIds[(size_t)gi*16 + j] = prt->cellIds[j];
        }
        if (r->gaussian) {
            if (!dense) {
                t0 = nowSec();
                fy->buildCellPartList(yProc.get());
                fy->setCellVolFraction(yProc.get());
                r->tAccum += nowSec() - t0;
            } else if (yProc->foundParticles.size()) {
                t0 = nowSec();
                fy->calcInterpWeightGaussian(yProc->foundParticles);
                r->tWeights += nowSec() - t0;
                t0 = nowSec();
                const int N = r->alpha.size();
                pvol.assign(N, 0.0); upx.assign(N, 0.0); upy.assign(N, 0.0); upz.assign(N, 0.0); touched.assign(N, 0);
                for (auto& prt : yProc->foundParticles) {
                    for (size_t i = 0; i < prt->interpCellWeight.size(); ++i) {
                        const int c = prt->interpCellWeight[i].first;
                        const double w = prt->interpCellWeight[i].second;
                        const Foam::vector uc = prt->linearVelocity*w*prt->vol;   // F.C:279
                        if (!touched[c]) { touched[c] = 1; pvol[c] = prt->vol*w; upx[c] = uc.x(); upy[c] = uc.y(); upz[c] = uc.z(); }
                        else { pvol[c] += prt->vol*w; upx[c] += uc.x(); upy[c] += uc.y(); upz[c] += uc.z(); }
                    }
                }
                for (int c = 0; c < N; ++c) {
                    if (!touched[c]) continue;
                    const double pvolC = 1.0 - (pvol[c]/r->mesh.V()[c]);          // F.C:323-325
                    r->alpha[c] = (pvolC > 0.10) ? pvolC : 0.10;
                    r->uParticle[c] = Foam::vector(upx[c], upy[c], upz[c])/(r->mesh.V()[c]);
                }
                r->tAccum += nowSec() - t0;
            }
            t0 = nowSec();
            if (r->addedMass || r->gaussTorque) {
                for (const auto& prt : yProc->foundParticles) {              // calcHydroForce's loop (F.C:331-344) + the dormant force
                    fy->initParticleForce(prt.get());
                    fy->hydroDragForce(prt.get());
                    fy->archimedesForce(prt.get());
                    if (r->addedMass) fy->addedMassForce(prt.get());
                }
                if (r->gaussTorque) fy->calcHydroTorque(yProc.get());
            } else {
                fy->calcHydroForce(yProc.get());
            }
            r->tForce += nowSec() - t0;
        } else {
            t0 = nowSec();
            fy->calcHydroForce(yProc.get());
            fy->calcHydroTorque(yProc.get());
            r->tForce += nowSec() - t0;
        }
    }
    t0 = nowSec();
    fy->sendHydroForceYadeMPI();
    fy->exchangeDT();
    r->tSend = nowSec() - t0;
    collectOutputs(r, n, found, force6);
}

void ref_get_lists(void* h, int n, int* counts, int* ids16)
{
    Ref* r = (Ref*)h;
    for (int i = 0; i < n && i < (int)r->listCount.size(); ++i) {
        counts[i] = r->listCount[i];
        for (int j = 0; j < 16; ++j) ids16[(size_t)i*16 + j] = r->listIds[(size_t)i*16 + j];
    }
}
void ref_get_times(void* h, double* out5)
{
    Ref* r = (Ref*)h;
    out5[0] = r->tLocate; out5[1] = r->tWeights; out5[2] = r->tAccum; out5[3] = r->tForce; out5[4] = r->tSend;
}

#endif

void ref_set_source_zero(void* h) { ((Ref*)h)->fy->setSourceZero(); }

}

// Synthetic-input helper shared by tests and bench: std::mt19937_64(seed) +
// uniform_real_distribution<double>(0,1), the recipe SURVEY.md section 8(c) used for its known answers.
#include <random>
extern "C" void ref_mt19937_64_uniform(unsigned long long seed, long n, double* out)
{
    std::mt19937_64 g(seed);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    for (long i = 0; i < n; ++i) out[i] = u(g);
}
