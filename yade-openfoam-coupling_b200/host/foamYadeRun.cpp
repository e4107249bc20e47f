// foamYadeRun.cpp -- the reference's two solvers as ONE standalone C++ program over the C ABI, for a box case directory
// and no OpenFOAM installation:
//
//     foamYadeB200 -case <dir> [-solver icoFoamYade|pimpleFoamYade] [-steps N] [-particles records.bin] [-gaussian]
//                  [-noWrite] [-device D]          run the time loop on the GPU
//     foamYadeB200 -case <dir> [-solver ...] -dump  print what was read as JSON and exit (no GPU needed)
//
// main() is the solvers' own: read the case (createFields.H), construct the operator, setScalarProperties, then
// `while (runTime.loop())` with the loop body of icoFoamYade.C:65-149 / pimpleFoamYade.C:65-110 -- here fy_ico_pre |
// fy_pimple_pre, fy_set_particle_action, fy_ico_solve | fy_pimple_solve, fy_set_source_zero on resident fields -- the solver
// log in OpenFOAM's format and runTime.write() (foamCase.H).  -particles: [P][10] float64 wire records held fixed, standing
// in for the Yade side (a Yade peer drives the MPI host class FoamYadeB200.H instead; without MPI in this image the wire
// is not opened here); without it the coupling call runs with zero particles.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "foamCase.H"

static void ck(fy_handle h, const char* what, int rc)
{
    if (rc == FY_OK) return;
    std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, fy_last_error(h));
    std::exit(2);
}

static unsigned long long fnv(const void* p, size_t n)
{
    unsigned long long h = 1469598103934665603ull;
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

static void dump(const fycase::Case& c, const fycase::MeshData& m)
{
    std::printf("{\"n\": [%d, %d, %d], \"origin\": [%.17g, %.17g, %.17g], \"L\": [%.17g, %.17g, %.17g], \"Uname\": \"%s\", \"nu\": %.17g, "
                "\"rhoP\": %.17g, \"rhoF\": %.17g, \"g\": [%.17g, %.17g, %.17g], \"deltaT\": %.17g, \"startTime\": %.17g, \"endTime\": %.17g, "
                "\"writeInterval\": %.17g, \"writePrecision\": %d,\n",
                c.box.n[0], c.box.n[1], c.box.n[2], c.box.origin[0], c.box.origin[1], c.box.origin[2], c.box.L[0], c.box.L[1], c.box.L[2],
                c.Uname.c_str(), c.nu, c.rhoP, c.rhoF, c.g[0], c.g[1], c.g[2], c.deltaT, c.startTime, c.endTime, c.writeInterval, c.writePrecision);
    const fy_piso_controls& p = c.piso;
    std::printf(" \"piso\": {\"nCorrectors\": %d, \"nNonOrthogonalCorrectors\": %d, \"momentumPredictor\": %d, \"pRefCell\": %d, \"pRefValue\": %.17g, "
                "\"pTol\": %.17g, \"pRelTol\": %.17g, \"pFinalTol\": %.17g, \"pFinalRelTol\": %.17g, \"UTol\": %.17g, \"URelTol\": %.17g, \"maxIter\": %d, "
                "\"preconditioner\": %d},\n",
                p.nCorrectors, p.nNonOrthogonalCorrectors, p.momentumPredictor, p.pRefCell, p.pRefValue, p.pTol, p.pRelTol, p.pFinalTol,
                p.pFinalRelTol, p.UTol, p.URelTol, p.maxIter, p.preconditioner);
    std::printf(" \"pimple\": {\"nOuterCorrectors\": %d, \"relaxU\": %.17g, \"relaxUFinal\": %.17g, \"relaxP\": %.17g, \"relaxPFinal\": %.17g},\n",
                c.nOuterCorrectors, c.relaxU, c.relaxUFinal, c.relaxP, c.relaxPFinal);
    std::printf(" \"patches\": [");
    for (size_t q = 0; q < c.patches.size(); ++q) {
        const fycase::CasePatch& cp = c.patches[q];
        std::printf("%s{\"name\": \"%s\", \"type\": \"%s\", \"sides\": [", q ? ", " : "", cp.name.c_str(), cp.type.c_str());
        for (size_t s = 0; s < cp.sides.size(); ++s) std::printf("%s\"%s\"", s ? ", " : "", fycase::SIDES[cp.sides[s]]);
        std::printf("], \"start\": %d, \"nFaces\": %d, \"bcU\": %d, \"valueU\": [%.17g, %.17g, %.17g], \"bcP\": %d, \"valueP\": %.17g}", cp.start,
                    cp.nFaces, cp.bcU, cp.valueU[0], cp.valueU[1], cp.valueU[2], cp.bcP, cp.valueP);
    }
    std::printf("],\n");
    // hashes of the raw bytes: the Python twin (foamcase.py + mesh.py) must produce the same arrays bit for bit
    std::printf(" \"hash\": {\"U\": \"%016llx\", \"p\": \"%016llx\", \"C\": \"%016llx\", \"V\": \"%016llx\", \"owner\": \"%016llx\", \"neighbour\": \"%016llx\", "
                "\"Sf\": \"%016llx\", \"magSf\": \"%016llx\", \"deltaCoeffs\": \"%016llx\"",
                fnv(c.U.data(), c.U.size() * 8), fnv(c.p.data(), c.p.size() * 8), fnv(m.C.data(), m.C.size() * 8), fnv(m.V.data(), m.V.size() * 8),
                fnv(m.owner.data(), m.owner.size() * 4), fnv(m.neighbour.data(), m.neighbour.size() * 4), fnv(m.Sf.data(), m.Sf.size() * 8),
                fnv(m.magSf.data(), m.magSf.size() * 8), fnv(m.deltaCoeffs.data(), m.deltaCoeffs.size() * 8));
    for (size_t q = 0; q < m.pdata.size(); ++q)
        std::printf(", \"faceCells%zu\": \"%016llx\", \"bSf%zu\": \"%016llx\", \"bDeltaCoeffs%zu\": \"%016llx\"", q,
                    fnv(m.pdata[q].faceCells.data(), m.pdata[q].faceCells.size() * 4), q, fnv(m.pdata[q].Sf.data(), m.pdata[q].Sf.size() * 8), q,
                    fnv(m.pdata[q].deltaCoeffs.data(), m.pdata[q].deltaCoeffs.size() * 8));
    std::printf("}}\n");
}

static std::string g6(double x) { return fycase::fmtG(x, 6); }

int main(int argc, char** argv)
{
    std::string dir, solver = "icoFoamYade", particles, wtime;
    int steps = 0, device = 0;
    bool dumpOnly = false, gaussian = false, noWrite = false;
    for (int a = 1; a < argc; ++a) {
        const std::string k = argv[a];
        auto next = [&]() -> std::string { if (a + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", k.c_str()); std::exit(1); } return argv[++a]; };
        if (k == "-case") dir = next();
        else if (k == "-solver") solver = next();
        else if (k == "-steps") steps = std::atoi(next().c_str());
        else if (k == "-particles") particles = next();
        else if (k == "-device") device = std::atoi(next().c_str());
        else if (k == "-writeNow") wtime = next();
        else if (k == "-dump") dumpOnly = true;
        else if (k == "-gaussian") gaussian = true;
        else if (k == "-noWrite") noWrite = true;
        else { std::fprintf(stderr, "unknown option %s\n", k.c_str()); return 1; }
    }
    if (dir.empty()) { std::fprintf(stderr, "usage: foamYadeB200 -case <dir> [-solver icoFoamYade|pimpleFoamYade] [-steps N] [-particles f] [-dump]\n"); return 1; }
    const bool pimple = solver == "pimpleFoamYade";
    fycase::Case c;
    fycase::MeshData mesh;
    try {
        c = fycase::loadCase(dir, "0", solver);
        fycase::buildMesh(c, mesh);
        if (!wtime.empty()) {                       // round-trip check of the writer: the fields as read, written as time <t>
            fycase::writeTime(c, std::atof(wtime.c_str()), c.U.data(), c.p.data());
            return 0;
        }
    } catch (const fycase::Error& e) {
        std::fprintf(stderr, "foamYadeB200: %s\n", e.what());
        return 3;
    }
    if (dumpOnly) { dump(c, mesh); return 0; }

    fy_handle h = nullptr;
    int rc = fy_create(&mesh.desc, device, &h);
    if (rc != FY_OK) { std::fprintf(stderr, "fy_create failed (%d): %s\n", rc, fy_last_error(nullptr)); return 2; }
    if (!fy_fv_supported(h)) { std::fprintf(stderr, "%s\n", fy_last_error(h)); return 2; }
    ck(h, "fy_set_properties", fy_set_properties(h, c.rhoP, c.rhoF, c.nu, (pimple || gaussian) ? 1 : 0));
    ck(h, "fy_set_viscosity", fy_set_viscosity(h, c.nu));
    ck(h, "fy_set_piso_controls", fy_set_piso_controls(h, &c.piso));
    if (pimple) ck(h, "fy_set_pimple_controls", fy_set_pimple_controls(h, c.nOuterCorrectors, c.relaxU, c.relaxUFinal, c.relaxP, c.relaxPFinal));
    ck(h, "fy_upload_field U", fy_upload_field(h, FY_F_U, c.U.data()));
    ck(h, "fy_upload_field p", fy_upload_field(h, FY_F_P, c.p.data()));
    ck(h, "fy_create_phi", fy_create_phi(h));
    std::vector<double> pd;
    if (!particles.empty()) {
        const std::string raw = fycase::slurp(particles);
        if (raw.size() % 80) { std::fprintf(stderr, "%s: not a multiple of 80 bytes ([P][10] float64 records)\n", particles.c_str()); return 1; }
        pd.resize(raw.size() / 8);
        std::memcpy(pd.data(), raw.data(), raw.size());
    }
    const int P = (int)(pd.size() / 10);
    std::vector<int> found((size_t)std::max(P, 1));
    std::vector<double> force((size_t)std::max(P, 1) * 6);
    bool empty[6] = {false, false, false, false, false, false};
    for (const fycase::CasePatch& cp : c.patches) if (cp.bcU == FY_BC_EMPTY) for (int s : cp.sides) empty[s] = true;
    const int nSteps = steps > 0 ? steps : (int)std::lround((c.endTime - c.startTime) / c.deltaT);
    const int every = std::max(1, (int)std::lround(c.writeInterval));
    const char* pname = c.piso.preconditioner == FY_PRECOND_DIC ? "DICPCG" : (c.piso.preconditioner == FY_PRECOND_DIAGONAL ? "diagonalPCG" : "PCG");
    std::vector<double> U(c.U.size()), p(c.p.size());
    double t = c.startTime;
    std::printf("\nStarting time loop\n\n");
    for (int it = 1; it <= nSteps; ++it) {
        t += c.deltaT;
        ck(h, pimple ? "fy_pimple_pre" : "fy_ico_pre", pimple ? fy_pimple_pre(h, c.deltaT) : fy_ico_pre(h, c.deltaT));
        ck(h, "fy_set_particle_action", fy_set_particle_action(h, c.deltaT, pd.data(), P, found.data(), force.data()));
        ck(h, pimple ? "fy_pimple_solve" : "fy_ico_solve", pimple ? fy_pimple_solve(h, c.deltaT, c.g) : fy_ico_solve(h, c.deltaT));
        ck(h, "fy_set_source_zero", fy_set_source_zero(h));
        fy_ico_stats st;
        ck(h, "fy_get_ico_stats", fy_get_ico_stats(h, &st));
        std::printf("Time = %s\n\nCourant Number mean: %s max: %s\n", g6(t).c_str(), g6(st.meanCoNum).c_str(), g6(st.CoNum).c_str());
        for (int j = 0; j < 3; ++j) {
            if (empty[2 * j] && empty[2 * j + 1]) continue;          // the component of an empty direction is not solved
            std::printf("smoothSolver:  Solving for %s%c, Initial residual = %s, Final residual = %s, No Iterations %d\n", c.Uname.c_str(), "xyz"[j],
                        g6(st.U[j].initialResidual).c_str(), g6(st.U[j].finalResidual).c_str(), st.U[j].nIterations);
        }
        for (int q = 0; q < st.nPSolves && q < 8; ++q) {
            std::printf("%s:  Solving for p, Initial residual = %s, Final residual = %s, No Iterations %d\n", pname, g6(st.p[q].initialResidual).c_str(),
                        g6(st.p[q].finalResidual).c_str(), st.p[q].nIterations);
            if ((q + 1) % (c.piso.nNonOrthogonalCorrectors + 1) == 0) {
                const int k = (q + 1) / (c.piso.nNonOrthogonalCorrectors + 1) - 1;
                std::printf("time step continuity errors : sum local = %s, global = %s\n", g6(st.corrSumLocal[k]).c_str(), g6(st.corrGlobal[k]).c_str());
            }
        }
        std::printf("\n");
        const bool last = it == nSteps;
        if (!noWrite && (it % every == 0 || (last && steps > 0))) {
            ck(h, "fy_download_field U", fy_download_field(h, FY_F_U, U.data()));
            ck(h, "fy_download_field p", fy_download_field(h, FY_F_P, p.data()));
            try { fycase::writeTime(c, t, U.data(), p.data()); }
            catch (const fycase::Error& e) { std::fprintf(stderr, "foamYadeB200: %s\n", e.what()); return 3; }
        }
    }
    std::printf("End\n\n");
    fy_destroy(h);
    return 0;
}
