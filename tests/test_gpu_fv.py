"""GPU: the sm_100a finite-volume / PISO kernels, called through the C ABI (include/fycuda.h), against the
CPU oracle (oracle/fv_oracle.cc, pinned by the OpenFOAM cavity tutorial log in tests/test_fv_oracle.py).
Bar: per-cell operators bit-exact (same operation order, no FMA contraction); solves and whole time steps
within 1e-10 relative L2 (the only difference is the association of the global dot products / norms) with
identical iteration counts."""
import numpy as np
import pytest

from oracle import port
from tests import cases, cases_fv
from tests.test_fv_oracle import CAVITY_LOG, sig6

pytestmark = pytest.mark.gpu
TOL = cases.TOL


def _rand_fields(mo, seed=0):
    rng = np.random.default_rng(seed)
    N = mo["nCells"]
    nF = mo["nInternalFaces"] + sum(p["faceCells"].size for p in mo["patches"])
    return rng.standard_normal((N, 3)), rng.standard_normal(N), rng.standard_normal(nF)


@pytest.mark.parametrize("case", ["channel", "cavity3d", "cavity2d", "flat"])
def test_fvc_operators_bit_exact(pkg, case):
    if case == "channel":
        mo, mp = cases_fv.channel(pkg, (13, 9, 7))
    elif case == "cavity3d":
        mo, mp = cases_fv.cavity3d(pkg, (8, 10, 12), (0.1, 0.2, 0.3))
    elif case == "flat":
        mo, mp = cases_fv.channel(pkg, (1, 6, 5))
    else:
        mo, mp = cases_fv.cavity2d(pkg, 12)
    U, p, phi = _rand_fields(mo, 1)
    O = port.IcoOracle(mo)
    E = pkg.Engine(mp)
    assert E.fv_supported(), E.L.fy_last_error(E.h).decode()
    assert np.array_equal(E.grad_vector(U), O.grad_vector(U))
    assert np.array_equal(E.grad_scalar(p), O.grad_scalar(p))
    if case == "cavity2d":
        Fi = mo["nInternalFaces"]
        phi[Fi + 12 * 4:] = 0.0          # empty faces carry no flux
    assert np.array_equal(E.div_flux(phi), O.div_flux(phi))
    # pimpleFoamYade.C:73,75: fvc::div(phic, Uc) and fvc::laplacian(alphac, Uc)
    gamma = np.random.default_rng(5).uniform(0.3, 1.0, mo["nCells"])
    assert np.array_equal(E.div_phi_vector(phi, U), O.div_phi_vector(phi, U))
    assert np.array_equal(E.laplacian_gamma_vector(gamma, U), O.laplacian_gamma_vector(gamma, U))
    # face field round trip through the owner-slot layout
    E.upload("phi", phi)
    assert np.array_equal(E.download("phi"), phi)
    E.close()
    O.close()


@pytest.mark.parametrize("n", [(9, 7, 5), (32, 32, 32), (12, 70, 9), (40, 33, 19), (5, 100, 3)])
def test_dic_precondition_bit_exact(pkg, n):
    mo, mp = cases_fv.cavity3d(pkg, n)
    rng = np.random.default_rng(2)
    N, Fi = mo["nCells"], mo["nInternalFaces"]
    upper = -rng.uniform(0.5, 1.5, Fi)
    diag = np.zeros(N)
    np.subtract.at(diag, mo["owner"], upper)
    np.subtract.at(diag, mo["neighbour"], upper)
    diag += rng.uniform(0.01, 0.05, N)
    r = rng.standard_normal(N)
    O = port.IcoOracle(mo)
    E = pkg.Engine(mp)
    assert np.array_equal(E.dic(diag, upper, r), O.dic(diag, upper, r))
    E.close()
    O.close()


@pytest.mark.parametrize("n,cluster,warps,planes", [((8, 8, 150), "16", "8", "1"), ((6, 40, 70), "2", "4", "2"), ((10, 33, 37), "1", "8", "1"),
                                                    ((9, 70, 20), "4", "2", "4"), ((5, 5, 300), "8", "1", "4"), ((7, 45, 131), "16", "3", "2"),
                                                    ((40, 33, 19), "16", "8", "1"), ((3, 100, 9), "2", "1", "4")])
def test_sweeps_across_clusters_and_helpers(pkg, monkeypatch, n, cluster, warps, planes):
    """every hand-off route of the pencil pipeline: shared-memory channels, DSMEM between the CTAs of a cluster,
    the L2 z-helper between clusters (nz > planes per cluster) and the y-helpers between j-blocks (ny > 32)"""
    monkeypatch.setenv("FY_PENCIL_CLUSTER", cluster)
    monkeypatch.setenv("FY_PENCIL_W", warps)
    monkeypatch.setenv("FY_PEN2_W", warps)
    monkeypatch.setenv("FY_PEN2_Z", planes)
    monkeypatch.setenv("FY_PEN2_SMEM_KB", "200")
    mo, mp = cases_fv.cavity3d(pkg, n)
    rng = np.random.default_rng(7)
    N, Fi = mo["nCells"], mo["nInternalFaces"]
    upper = -rng.uniform(0.5, 1.5, Fi)
    lower = -rng.uniform(0.5, 1.5, Fi)
    diag = np.zeros(N)
    np.subtract.at(diag, mo["owner"], upper)
    np.subtract.at(diag, mo["neighbour"], upper)
    diag += rng.uniform(0.01, 0.05, N)
    r = rng.standard_normal(N)
    O = port.IcoOracle(mo)
    E = pkg.Engine(mp)
    assert np.array_equal(E.dic(diag, upper, r), O.dic(diag, upper, r))
    dg = np.zeros(N)
    np.subtract.at(dg, mo["owner"], lower)
    np.subtract.at(dg, mo["neighbour"], upper)
    dg += rng.uniform(0.5, 1.0, N)
    xo, po = O.smooth(dg, lower, upper, r, np.zeros(N), tol=1e-9)
    xe, pe = E.smooth(dg, lower, upper, r, np.zeros(N), tol=1e-9)
    assert pe["iters"] == po["iters"] and cases.rel_l2(xe, xo) <= 1e-13
    E.close()
    O.close()


@pytest.mark.parametrize("pre,n", [("DIC", (24, 20, 16)), ("diagonal", (24, 20, 16)), ("none", (24, 20, 16)),
                                   ("DIC", (33, 40, 6)), ("none", (65, 33, 3)), ("diagonal", (1, 70, 5))])
def test_pcg_matches_oracle(pkg, pre, n):
    """(33, ..), (65, ..): nx + 31 is a multiple of 32, so a slab's last row of the pencil layout holds real cells and the
    row-blocked Amul's x-neighbours cross into the neighbouring slabs' pads."""
    mo, mp = cases_fv.cavity3d(pkg, n)
    rng = np.random.default_rng(4)
    N, Fi = mo["nCells"], mo["nInternalFaces"]
    upper = rng.uniform(0.5, 1.5, Fi)              # negative-definite Laplacian-like matrix, as pEqn's
    diag = np.zeros(N)
    np.subtract.at(diag, mo["owner"], upper)
    np.subtract.at(diag, mo["neighbour"], upper)
    diag -= rng.uniform(0.001, 0.01, N)
    b = rng.standard_normal(N)
    O = port.IcoOracle(mo)
    E = pkg.Engine(mp)
    for tol, rel in ((1e-6, 0.05), (1e-10, 0.0)):
        xo, po = O.pcg(diag, upper, b, np.zeros(N), tol=tol, relTol=rel, preconditioner=pre)
        xe, pe = E.pcg(diag, upper, b, np.zeros(N), tol=tol, relTol=rel, preconditioner=pre)
        assert pe["iters"] == po["iters"] and po["iters"] > 3
        np.testing.assert_allclose(pe["initial"], po["initial"], rtol=1e-12)
        np.testing.assert_allclose(pe["final"], po["final"], rtol=0.05)   # rounding of the dot products, amplified over the iterations
        assert cases.rel_l2(xe, xo) <= TOL
    # maxIter cap and an already-converged start behave like PCG.C
    xo, po = O.pcg(diag, upper, b, np.zeros(N), tol=1e-14, relTol=0.0, maxIter=5, preconditioner=pre)
    xe, pe = E.pcg(diag, upper, b, np.zeros(N), tol=1e-14, relTol=0.0, maxIter=5, preconditioner=pre)
    assert pe["iters"] == po["iters"] == 6 and cases.rel_l2(xe, xo) <= TOL
    xs, _ = O.pcg(diag, upper, b, np.zeros(N), tol=1e-13, relTol=0.0, preconditioner=pre)
    xo, po = O.pcg(diag, upper, b, xs, tol=1e-6, relTol=0.0, preconditioner=pre)
    xe, pe = E.pcg(diag, upper, b, xs, tol=1e-6, relTol=0.0, preconditioner=pre)
    assert pe["iters"] == po["iters"] == 0 and np.array_equal(xe, xs)
    E.close()
    O.close()


def test_smooth_solver_matches_oracle(pkg):
    mo, mp = cases_fv.channel(pkg, (20, 14, 10))
    rng = np.random.default_rng(5)
    N, Fi = mo["nCells"], mo["nInternalFaces"]
    upper = -rng.uniform(0.5, 1.5, Fi)
    lower = -rng.uniform(0.5, 1.5, Fi)
    diag = np.zeros(N)
    np.subtract.at(diag, mo["owner"], lower)
    np.subtract.at(diag, mo["neighbour"], upper)
    diag += rng.uniform(0.5, 1.0, N)
    b = rng.standard_normal(N)
    O = port.IcoOracle(mo)
    E = pkg.Engine(mp)
    xo, po = O.smooth(diag, lower, upper, b, np.zeros(N), tol=1e-9)
    xe, pe = E.smooth(diag, lower, upper, b, np.zeros(N), tol=1e-9)
    assert pe["iters"] == po["iters"] and po["iters"] > 2
    assert cases.rel_l2(xe, xo) <= 1e-13            # sweeps are bit-exact; only the stopping norm is summed differently
    E.close()
    O.close()


def test_openfoam_cavity_tutorial_log_on_gpu(pkg):
    """the device solver prints the stock cavity tutorial's log (see tests/test_fv_oracle.py for its provenance)"""
    mo, mp = cases_fv.cavity2d(pkg, 20)
    E = pkg.Engine(mp)
    assert E.fv_supported(), E.L.fy_last_error(E.h).decode()
    E.set_piso_controls(nu=0.01)
    E.create_phi()
    for ref in CAVITY_LOG:
        E.ico_pre(0.005)
        E.ico_solve(0.005)
        st = E.ico_stats()
        assert (sig6(st["meanCoNum"]), sig6(st["CoNum"])) == ref["Co"]
        for j, k in ((0, "Ux"), (1, "Uy")):
            u = st["U"][j]
            assert (sig6(u["initial"]), sig6(u["final"]), u["iters"]) == ref[k], k
        for j, k in ((0, "p1"), (1, "p2")):
            q = st["p"][j]
            assert (sig6(q["initial"]), sig6(q["final"]), q["iters"]) == ref[k], k
        if ref["c1"] is not None:
            assert sig6(st["corrSumLocal"][0]) == ref["c1"]
        assert sig6(st["corrSumLocal"][1]) == ref["c2"]
    E.close()


@pytest.mark.parametrize("case", ["cavity3d", "channel", "cavity2d"])
def test_ico_steps_match_oracle(pkg, case):
    if case == "cavity3d":
        mo, mp = cases_fv.cavity3d(pkg, (20, 20, 20))
        U, p, dt, nu, ctl = np.zeros((mo["nCells"], 3)), np.zeros(mo["nCells"]), 0.004, 0.01, None
    elif case == "channel":
        mo, mp = cases_fv.channel(pkg, (28, 14, 12))
        U, p = cases_fv.channel_init(mo["C"])
        dt, nu, ctl = 0.02, 0.005, dict(nCorrectors=3, nNonOrthogonalCorrectors=1)
    else:
        mo, mp = cases_fv.cavity2d(pkg, 24)
        U, p, dt, nu, ctl = np.zeros((mo["nCells"], 3)), np.zeros(mo["nCells"]), 0.004, 0.01, dict(preconditioner="diagonal")

    def src(it, Ucur, vGrad):        # a momentum source that depends on the fields, like the particle reaction
        s = np.zeros_like(Ucur)
        s[:, 0] = 0.5 * np.sin(40.0 * mo["C"][:, 1]) - 0.2 * Ucur[:, 0]
        s[:, 1] = 0.1 * vGrad[:, 3] * 1e-2
        return s * (it + 1)

    o = cases_fv.run_oracle_steps(mo, U, p, dt, 3, nu, ctl, src)
    e = cases_fv.run_engine_steps(pkg, mp, U, p, dt, 3, nu, ctl, src)
    cases_fv.compare_fluid(o, e)
    assert abs(e["stats"][-1]["globalContErr"]) < 1e-8
    assert e["stats"][-1]["sumLocalContErr"] < 1e-6


@pytest.mark.parametrize("config", ["small_point", "C1_point", "C1_gaussian"])
def test_coupled_icoFoamYade_step(pkg, config):
    """vGrad = grad(U) -> setParticleAction -> UEqn/PISO -> setSourceZero, engine vs (unmodified reference coupling +
    oracle fluid step).  C1 = BASELINE.json configs[0]: lid-driven cavity, 32^3 cells, 1k particles (seed 42); the
    point-force branch is the one icoFoamYade hard-codes (icoFoamYade.C:53), the Gaussian one is the operator's other
    constructor setting (FoamYade.H:117)."""
    from oracle import ref
    if config == "small_point":
        n, P, seed, gaussian, steps, dt, nu = 16, 3000, 11, False, 2, 2e-3, 1e-3
        mo, mp = cases_fv.cavity3d(pkg, (n, n, n), (1.0, 1.0, 1.0))
        U0 = 0.2 * cases.fields_for(mo["C"])["U"]
    else:
        n, P, seed, gaussian, steps, dt, nu = 32, 1000, 42, config == "C1_gaussian", 4, 5e-3, 0.01
        mo, mp = cases_fv.cavity3d(pkg, (n, n, n), (1.0, 1.0, 1.0))
        U0 = np.zeros((mo["nCells"], 3))
    pd = cases.particles(P, seed, radius=0.1 / n, moving=True)
    f = cases.fields_for(mo["C"])
    # oracle side
    O = port.IcoOracle(mo, nu=nu)
    O.field("U")[:] = U0
    O.create_phi()
    R = ref.RefFoamYade(mo, gaussian)
    R.set_properties(cases.RHOP, cases.RHOF, nu)
    # engine side
    E = pkg.Engine(mp)
    E.set_properties(cases.RHOP, cases.RHOF, nu, gaussian)
    E.set_piso_controls(nu=nu)
    E.upload("U", U0)
    E.create_phi()
    if gaussian:                      # gradP / divT are solver-owned inputs of the Gaussian branch: same values both sides
        for k in ("gradP", "divT"):
            R.field(k)[:] = 1e-2 * f[k]
            E.upload(k, 1e-2 * f[k])
    for step in range(steps):
        O.pre(dt)
        R.field("U")[:] = O.field("U")
        R.field("vGrad")[:] = O.field("vGrad")
        fo, Fo = R.step(dt, pd, pieces=True)
        O.field("uSource")[:] = R.field("uSource")
        O.solve(dt)
        R.set_source_zero()
        E.ico_pre(dt)
        fe, Fe = E.set_particle_action(dt, pd)
        E.ico_solve(dt)
        E.set_source_zero()
        assert np.array_equal(fo, fe)
        assert cases.rel_l2(Fe, Fo) <= TOL
        assert cases.rel_l2(E.download("U"), O.field("U")) <= TOL
        assert cases.rel_l2(E.download("p"), O.field("p")) <= TOL
        assert not np.any(E.download("uSource"))
        so, se = O.stats(), E.ico_stats()
        assert [q["iters"] for q in so["p"]] == [q["iters"] for q in se["p"]]
    E.close()
    R.close()
    O.close()


def _pimple_drive(mo, it):
    """synthetic void fraction / implicit drag / momentum source / gravity of one step (what a coupling pass leaves)"""
    C = mo["C"]
    lo, hi = C.min(0), C.max(0)
    ctr = np.array([0.4, 0.55, 0.5])
    x = np.where(hi > lo, (C - lo) / np.where(hi > lo, hi - lo, 1.0), ctr)
    blob = np.exp(-(((x - ctr) ** 2).sum(1)) / 0.04)
    alpha = 1.0 - (0.35 + 0.05 * it) * blob
    drag = -(40.0 + 10.0 * it) * (1.0 - alpha)
    src = np.stack([0.3 * (1.0 - alpha) * np.sin(5.0 * x[:, 1]), -0.8 * (1.0 - alpha), 0.1 * blob * x[:, 0]], 1)
    return alpha, drag, src


@pytest.mark.parametrize("case", ["cavity3d", "channel", "cavity2d", "no_predictor"])
def test_pimpleFoamYade_steps_match_oracle(pkg, case):
    """UcEqn.H + pEqn.H + continuityErrs.H (pimpleFoamYade.C:82-104) on the device against oracle/fv_oracle.cc's
    pimpleSolve, with a non-uniform void fraction, implicit drag, a momentum source and gravity.  What is assembled
    before the first linear solve (explicit stress term, UcEqn diagonal and source, 1/A, phicForces) is bit-exact;
    the fields after each step agree to 1e-10 with identical iteration counts."""
    ctl, g = {}, (0.0, -0.2, 0.05)
    if case in ("cavity3d", "no_predictor"):
        mo, mp = cases_fv.cavity3d(pkg, (18, 16, 14), (1.0, 0.9, 0.8))
        U, p, dt, nu = 0.2 * cases.fields_for(mo["C"])["U"], np.zeros(mo["nCells"]), 2e-3, 1e-2
        if case == "no_predictor":
            ctl = dict(momentumPredictor=0, nCorrectors=3)
    elif case == "channel":
        mo, mp = cases_fv.channel(pkg, (28, 14, 12))
        U, p = cases_fv.channel_init(mo["C"])
        dt, nu, ctl = 0.02, 0.005, dict(nCorrectors=2, nNonOrthogonalCorrectors=1)
    else:
        mo, mp = cases_fv.cavity2d(pkg, 24)
        U, p, dt, nu, g = np.zeros((mo["nCells"], 3)), np.zeros(mo["nCells"]), 0.004, 0.01, (0.0, -0.2, 0.0)
    N = mo["nCells"]
    O = port.IcoOracle(mo, nu=nu, **ctl)
    O.field("U")[:] = U
    O.field("p")[:] = p
    O.create_phi()
    E = pkg.Engine(mp)
    assert E.fv_supported(), E.L.fy_last_error(E.h).decode()
    E.set_piso_controls(nu=nu, **ctl)
    E.upload("U", U)
    E.upload("p", p)
    E.create_phi()
    for it in range(3):
        alpha, drag, src = _pimple_drive(mo, it)
        O.field("uSource")[:] = src
        O.pimple_solve(dt, alpha, drag, g)
        E.upload("alpha", alpha)
        E.upload("uSourceDrag", drag)
        E.upload("uSource", src)
        E.pimple_solve(dt, g)
        if it == 0:
            assert np.array_equal(E.fv_get("divDev"), O.pimple_field("divDev"))
            assert np.array_equal(E.fv_get("diagU"), O.field("diagU"))
            assert np.array_equal(E.fv_get("sourceU"), O.field("sourceU"))
            assert np.array_equal(E.fv_get("rAU"), O.field("rAU"))
            assert np.array_equal(E.fv_get("phicForces"), O.pimple_field("phicForces"))
            assert np.any(O.pimple_field("divDev")) and np.any(O.pimple_field("phicForces"))
        so, se = O.stats(), E.ico_stats()
        assert [q["iters"] for q in so["p"]] == [q["iters"] for q in se["p"]], it
        assert [q["iters"] for q in so["U"]] == [q["iters"] for q in se["U"]], it
        for k in ("U", "p", "phi"):
            assert cases.rel_l2(E.download(k), O.field(k)) <= TOL, (k, it)
        assert cases.rel_l2(E.fv_get("HbyA"), O.field("HbyA")) <= TOL
        assert cases.rel_l2(E.fv_get("phiHbyA"), O.field("phiHbyA")) <= TOL
        assert abs(se["sumLocalContErr"] - so["sumLocalContErr"]) <= 1e-6 * so["sumLocalContErr"] + 1e-18
        assert se["sumLocalContErr"] < 1e-6
    E.close()
    O.close()


@pytest.mark.parametrize("mode", ["outer2", "outer3_relaxed", "relaxU_only"])
def test_pimple_outer_correctors_and_relaxation_on_gpu(pkg, mode):
    """`while (pimple.loop())` (pimpleFoamYade.C:91-105) with nOuterCorrectors > 1, UcEqn.relax() (UcEqn.H:13) and p.relax()
    (pEqn.H:41) on the device against the oracle's restatement of OpenFOAM-6's pimpleControl / fvMatrix::relax /
    GeometricField::relax: the relaxed UcEqn (diagonal, source, 1/A) of the last outer corrector bit-exact, fields within
    1e-10, identical iteration counts in every pressure solve of every outer corrector."""
    pc = {"outer2": dict(nOuterCorrectors=2),
          "outer3_relaxed": dict(nOuterCorrectors=3, relaxU=0.7, relaxUFinal=1.0, relaxP=0.3, relaxPFinal=1.0),
          "relaxU_only": dict(nOuterCorrectors=1, relaxU=0.8)}[mode]
    mo, mp = cases_fv.channel(pkg, (28, 14, 12))
    U, p = cases_fv.channel_init(mo["C"])
    dt, nu, ctl, g = 0.02, 0.005, dict(nCorrectors=2), (0.0, -0.2, 0.05)
    O = port.IcoOracle(mo, nu=nu, **ctl)
    O.set_pimple_controls(**pc)
    O.field("U")[:] = U
    O.field("p")[:] = p
    O.create_phi()
    E = pkg.Engine(mp)
    assert E.fv_supported(), E.L.fy_last_error(E.h).decode()
    E.set_piso_controls(nu=nu, **ctl)
    E.set_pimple_controls(**pc)
    E.upload("U", U)
    E.upload("p", p)
    E.create_phi()
    for it in range(3):
        alpha, drag, src = _pimple_drive(mo, it)
        O.field("uSource")[:] = src
        O.pimple_solve(dt, alpha, drag, g)
        E.upload("alpha", alpha)
        E.upload("uSourceDrag", drag)
        E.upload("uSource", src)
        E.pimple_solve(dt, g)
        if it == 0 and mode != "outer3_relaxed":
            # (the last outer corrector starts from the same U on both sides only up to solver round-off once a linear
            # solve lies in between: bit-exact assembly is asserted where no solve precedes it or U enters nowhere)
            assert np.array_equal(E.fv_get("diagU"), O.field("diagU"))
            assert np.array_equal(E.fv_get("rAU"), O.field("rAU"))
        if it == 0 and mode == "relaxU_only":
            assert np.array_equal(E.fv_get("sourceU"), O.field("sourceU"))
        if it == 0 and mode == "outer3_relaxed":
            assert np.array_equal(E.fv_get("diagU"), O.field("diagU"))      # the diagonal does not depend on U
        so, se = O.stats(), E.ico_stats()
        assert so["nPSolves"] == se["nPSolves"] == pc["nOuterCorrectors"] * 2
        assert [q["iters"] for q in so["p"]] == [q["iters"] for q in se["p"]], it
        assert [q["iters"] for q in so["U"]] == [q["iters"] for q in se["U"]], it
        for k in ("U", "p", "phi"):
            assert cases.rel_l2(E.download(k), O.field(k)) <= TOL, (k, it)
        assert abs(se["cumulativeContErr"] - so["cumulativeContErr"]) <= 1e-9 * abs(so["cumulativeContErr"]) + 1e-16
    with pytest.raises(pkg.FyError):
        E.set_pimple_controls(nOuterCorrectors=1, relaxP=0.5)
        E.pimple_solve(dt, g)                              # p.relax() without a stored prevIter
    with pytest.raises(pkg.FyError):
        E.set_pimple_controls(nOuterCorrectors=5)          # 5 x 2 correctors > the 8 statistics slots
        E.pimple_solve(dt, g)
    E.close()
    O.close()


@pytest.mark.parametrize("case", ["cavity3d", "channel"])
def test_coupled_pimpleFoamYade_step(pkg, case):
    """The whole pimpleFoamYade time step with every field resident on the device (pimpleFoamYade.C:71-108): CourantNo,
    ddtU_f / gradP / divT / vGrad, Gaussian setParticleAction, UcEqn + PISO with the void fraction, implicit drag and
    momentum source the coupling pass just wrote, setSourceZero -- against oracle operators + the UNMODIFIED reference
    operator (oracle/_ref).  divT is evaluated with the alphac setSourceZero reset to 1 at the end of the previous step,
    as in the reference's loop."""
    from oracle import ref
    if case == "cavity3d":
        mo, mp = cases_fv.cavity3d(pkg, (24, 24, 24), (1.0, 1.0, 1.0))
        U0, p0, dt, nu = 0.3 * cases.fields_for(mo["C"])["U"], None, 2e-3, 1e-3
    else:
        mo, mp = cases_fv.channel(pkg, (28, 14, 12))
        U0, p0 = cases_fv.channel_init(mo["C"])
        dt, nu = 5e-3, 5e-3
    N = mo["nCells"]
    lo, hi = mo["C"].min(0), mo["C"].max(0)
    pd = cases.particles(1500, 9, radius=0.25 * float(np.cbrt(mo["V"][0])), moving=True)
    pd[:, 0:3] = lo + (0.1 + 0.8 * (pd[:, 0:3] - pd[:, 0:3].min(0)) / np.ptp(pd[:, 0:3], axis=0)) * (hi - lo)
    O = port.IcoOracle(mo, nu=nu)
    O.field("U")[:] = U0
    if p0 is not None:
        O.field("p")[:] = p0
    O.create_phi()
    R = ref.RefFoamYade(mo, True)
    R.set_properties(cases.RHOP, cases.RHOF, nu)
    E = pkg.Engine(mp)
    E.set_properties(cases.RHOP, cases.RHOF, nu, True)
    E.set_piso_controls(nu=nu)
    E.upload("U", U0)
    if p0 is not None:
        E.upload("p", p0)
    E.create_phi()
    for step in range(3):
        ddtU, gradP, divT, vGrad = O.pimple_pre(dt, R.field("alpha").reshape(N))
        for k, v in (("U", O.field("U")), ("ddtU", ddtU), ("gradP", gradP), ("divT", divT), ("vGrad", vGrad)):
            R.field(k)[:] = v.reshape(R.field(k).shape)
        fo, Fo = R.step(dt, pd, pieces=True)
        O.field("uSource")[:] = R.field("uSource")
        alphaO = R.field("alpha").reshape(N).copy()
        O.pimple_solve(dt, alphaO, R.field("uSourceDrag").reshape(N))
        R.set_source_zero()

        E.pimple_pre(dt)
        got = {k: E.download(k) for k in ("ddtU", "gradP", "divT", "vGrad")}
        fe, Fe = E.set_particle_action(dt, pd)
        alphaE = E.download("alpha")
        E.pimple_solve(dt)
        E.set_source_zero()

        for k, want in (("ddtU", ddtU), ("gradP", gradP), ("divT", divT), ("vGrad", vGrad)):
            if step == 0:
                assert np.array_equal(got[k].reshape(want.shape), want), k      # identical inputs: bit-exact operators
            else:
                assert cases.rel_l2(got[k].reshape(want.shape), want) <= TOL, (k, step)
        assert np.any(divT) and np.any(vGrad)
        assert np.array_equal(fo, fe)
        assert cases.rel_l2(Fe, Fo) <= TOL
        assert cases.rel_l2(alphaE.reshape(N), alphaO) <= TOL
        assert alphaO.min() < 1.0                     # the void fraction really enters UcEqn / pEqn
        assert cases.rel_l2(E.download("U"), O.field("U")) <= TOL
        assert cases.rel_l2(E.download("p"), O.field("p")) <= TOL
        assert cases.rel_l2(E.download("phi"), O.field("phi")) <= TOL
        assert not np.any(E.download("uSource")) and np.all(E.download("alpha") == 1.0)
        so, se = O.stats(), E.ico_stats()
        assert [q["iters"] for q in so["p"]] == [q["iters"] for q in se["p"]]
        assert abs(se["CoNum"] - so["CoNum"]) <= 1e-12 * max(so["CoNum"], 1e-300)
    E.close()
    R.close()
    O.close()


def test_pimple_closed_box_with_gravity_fixedFluxPressure_on_gpu(pkg):
    """constrainPressure on fixedFluxPressure walls (pimpleFoamYade/pEqn.H:21) on the device: a closed box under gravity --
    the case SURVEY.md 8(d) prescribes for C3 / C5 -- with a void-fraction blob, implicit drag and a momentum source that
    set the fluid in motion.  Assembly bit-exact, fields within 1e-10, identical iteration counts against the oracle; and
    with no forcing the box stays at rest with p = g.x (the hydrostatic balance of tests/test_fv_oracle.py on the GPU).
    grad(p) of the NEXT step's pre-coupling block sees the patch's gradient too (fvc::grad(p), pimpleFoamYade.C:74)."""
    from oracle import meshgen
    n, L = (10, 12, 8), (0.8, 1.2, 0.6)
    patches = [("walls", ["xmin", "xmax", "ymin", "ymax", "zmin", "zmax"])]
    mo = meshgen.hex_box_ldu(*n, *L, patches=patches)
    meshgen.set_bc(mo, "walls", bcP=meshgen.BC_FIXED_FLUX_PRESSURE)
    mp = pkg.box_mesh(*n, *L, patches=patches)
    pkg.set_bc(mp, "walls", bcP=pkg.BC_FIXED_FLUX_PRESSURE)
    N, Fi, C = mo["nCells"], mo["nInternalFaces"], mo["C"]
    g = (0.0, -9.81, 0.0)
    nu, dt = 0.01, 1e-3
    for forced in (False, True):
        O = port.IcoOracle(mo, nu=nu)
        O.create_phi()
        E = pkg.Engine(mp)
        assert E.fv_supported(), E.L.fy_last_error(E.h).decode()
        E.set_properties(cases.RHOP, cases.RHOF, nu, True)
        E.set_piso_controls(nu=nu)
        E.create_phi()
        with pytest.raises(pkg.FyError):
            E.ico_solve(dt)                               # icoFoamYade's step has no fixedFluxPressure path
        for it in range(4):
            if forced:
                alpha, drag, src = _pimple_drive(mo, it)
            else:
                alpha = np.ones(N) if it < 2 else 1 - 0.4 * np.exp(-((C - C.mean(0)) ** 2).sum(1) / 0.02)
                drag, src = -20.0 * (1 - alpha), np.zeros((N, 3))
            O.field("uSource")[:] = src
            O.pimple_solve(dt, alpha, drag, g)
            E.upload("alpha", alpha)
            E.upload("uSourceDrag", drag)
            E.upload("uSource", src)
            E.pimple_solve(dt, g)
            if it == 0:
                for k in ("diagU", "sourceU", "rAU"):
                    assert np.array_equal(E.fv_get(k), O.field(k)), k
                assert np.array_equal(E.fv_get("phicForces"), O.pimple_field("phicForces"))
            so, se = O.stats(), E.ico_stats()
            assert [q["iters"] for q in so["p"]] == [q["iters"] for q in se["p"]], (forced, it)
            # (at rest U and phi are solver-tolerance noise around zero: the scale of the comparison is then the velocity
            # / flux gravity alone would produce in one time step)
            floor = dict(U=np.sqrt(N) * 9.81 * dt, p=0.0, phi=np.sqrt(Fi) * mo["magSf"].max() * 9.81 * dt)
            for k in ("U", "p", "phi"):
                a, b = E.download(k), O.field(k)
                assert np.linalg.norm(a - b) <= TOL * max(np.linalg.norm(b), floor[k]), (k, forced, it)
            assert np.abs(E.download("phi")[Fi:]).max() < 1e-14               # no flux through the walls
            if not forced:
                p = E.download("p")
                assert np.abs(E.download("U")).max() < 1e-6                   # solver tolerance, not a flow
                assert np.abs(p - p[0] + 9.81 * (C[:, 1] - C[0, 1])).max() < 1e-3
            # the next step's pre-coupling block: gradP with the patch gradient in the boundary cells
            E.upload("alpha", np.ones(N))
            E.pimple_pre(dt)
            want = O.pimple_pre(dt, np.ones(N))[1]
            assert cases.rel_l2(E.download("gradP").reshape(want.shape), want) <= TOL
        E.close()
        O.close()


def test_unsupported_mesh_is_refused(pkg):
    mp = pkg.box_mesh(6, 6, 6)
    mp["V"] = mp["V"].copy()
    mp["V"][5] *= 1.5
    E = pkg.Engine(mp)
    assert not E.fv_supported()
    with pytest.raises(pkg.FyError):
        E.ico_solve(1e-3)
    E.close()
