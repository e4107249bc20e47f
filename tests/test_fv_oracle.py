"""CPU: pins the fluid-half oracle (oracle/fv_oracle.cc, this repo's restatement of the OpenFOAM-6 operators
behind icoFoamYade.C:65-149).

The reference ships no tests for this half and OpenFOAM is not installed here, so the pin is the solver log
OpenFOAM itself prints for its stock `cavity` tutorial (icoFoam, blockMesh 20x20x1 on 0.1 x 0.1 x 0.01 m,
lid U = (1 0 0), nu = 0.01, deltaT = 0.005, PISO nCorrectors 2, p: PCG/DIC 1e-06 relTol 0.05, pFinal relTol 0,
U: smoothSolver symGaussSeidel 1e-05) -- the same operator sequence icoFoamYade runs with uSource = 0.  The
numbers below are that log's first three time steps as OpenFOAM prints them (6 significant digits).  They were
written down from the public tutorial log, not produced by this code; the oracle reproduces every one of them,
which pins: Euler ddt, Gauss-linear convection / laplacian / gradient, boundary coefficients, the segregated
solve, symGaussSeidel sweeps and its residual normalisation, A()/H(), ddtCorr, pEqn assembly + setReference,
DIC-PCG with OpenFOAM's stopping rule, the flux update, continuityErrs and CourantNo.
"""
import numpy as np
import pytest

from oracle import meshgen, port

pytestmark = pytest.mark.skipif(not port.available(), reason="oracle/_build/liboracle.so not built")


def sig6(x):
    return float("%.6g" % x)


def cavity_mesh(n=20, nz=1):
    m = meshgen.hex_box_ldu(n, n, nz, 0.1, 0.1, 0.01,
                            patches=[("movingWall", ["ymax"]), ("fixedWalls", ["xmin", "xmax", "ymin"]),
                                     ("frontAndBack", ["zmin", "zmax"])])
    meshgen.set_bc(m, "movingWall", valueU=(1, 0, 0))
    meshgen.set_bc(m, "frontAndBack", bcU=meshgen.BC_EMPTY, bcP=meshgen.BC_EMPTY)
    return m


# log.icoFoam of the stock cavity tutorial: (Courant mean, max), Ux (init, final, iters), Uy, p corrector 1,
# sum local after corrector 1, p corrector 2, sum local after corrector 2
CAVITY_LOG = [
    dict(Co=(0.0, 0.0), Ux=(1.0, 8.90511e-06, 19), Uy=(0.0, 0.0, 0), p1=(1.0, 0.0492854, 12), c1=0.000466513,
         p2=(0.590864, 2.65225e-07, 35), c2=2.74685e-09),
    dict(Co=(0.0976825, 0.585607), Ux=(0.160686, 6.83031e-06, 19), Uy=(0.260828, 9.65939e-06, 18),
         p1=(0.428925, 0.0103739, 22), c1=None, p2=(0.30209, 5.26569e-07, 33), c2=6.61987e-09),
    dict(Co=(0.144686, 0.758934), Ux=(0.0447632, 9.12473e-06, 15), Uy=(0.0817804, 6.96014e-06, 17),
         p1=(0.131584, 0.00445878, 11), c1=None, p2=(0.0973392, 9.15339e-07, 31), c2=7.57271e-09),
]


def test_openfoam_cavity_tutorial_log():
    O = port.IcoOracle(cavity_mesh(), nu=0.01)
    O.create_phi()
    for ref in CAVITY_LOG:
        O.pre(0.005)
        O.solve(0.005)
        st = O.stats()
        assert (sig6(st["meanCoNum"]), sig6(st["CoNum"])) == ref["Co"]
        for j, k in ((0, "Ux"), (1, "Uy")):
            u = st["U"][j]
            assert (sig6(u["initial"]), sig6(u["final"]), u["iters"]) == ref[k], k
        assert st["U"][2]["iters"] == 0          # empty direction is not solved
        for j, k in ((0, "p1"), (1, "p2")):
            p = st["p"][j]
            assert (sig6(p["initial"]), sig6(p["final"]), p["iters"]) == ref[k], k
        if ref["c1"] is not None:
            assert sig6(st["corrSumLocal"][0]) == ref["c1"]
        assert sig6(st["corrSumLocal"][1]) == ref["c2"]
        assert abs(st["globalContErr"]) < 1e-17
    O.close()


def _mesh3d(n=(6, 5, 4)):
    m = meshgen.hex_box_ldu(*n, lx=1.0, ly=0.5, lz=2.0)
    meshgen.set_bc(m, "xmin", valueU=(0.3, 0.0, 0.0))
    meshgen.set_bc(m, "xmax", bcU=meshgen.BC_ZERO_GRADIENT, bcP=meshgen.BC_FIXED_VALUE, valueP=0.25)
    return m


def test_gradients_are_exact_for_linear_fields_in_the_interior():
    m = _mesh3d((8, 7, 6))
    O = port.IcoOracle(m)
    C = m["C"]
    A = np.array([[0.3, -1.0, 2.0], [0.7, 0.2, -0.4], [1.5, 0.0, 0.9]])      # U_j = sum_i C_i A_ij
    U = C @ A
    g = O.grad_vector(U).reshape(-1, 3, 3)
    nx, ny, nz = m["n"]
    idx = np.arange(m["nCells"]).reshape(nz, ny, nx)[1:-1, 1:-1, 1:-1].reshape(-1)
    np.testing.assert_allclose(g[idx], np.broadcast_to(A, (idx.size, 3, 3)), rtol=0, atol=1e-12)
    p = C @ np.array([2.0, -3.0, 0.5])
    gp = O.grad_scalar(p)
    np.testing.assert_allclose(gp[idx], np.broadcast_to([2.0, -3.0, 0.5], (idx.size, 3)), rtol=0, atol=1e-12)
    O.close()


def test_pimple_pre_operators_are_exact_on_polynomials():
    """fvc::div(phi, U) and fvc::laplacian(gamma, U) (pimpleFoamYade.C:73,75) against closed forms the Gauss-linear
    discretisation reproduces exactly in the interior, and fvo_pimple_pre's composition of them."""
    m = _mesh3d((9, 7, 6))
    O = port.IcoOracle(m, nu=0.02)
    C = m["C"]
    nx, ny, nz = m["n"]
    idx = np.arange(m["nCells"]).reshape(nz, ny, nx)[1:-1, 1:-1, 1:-1].reshape(-1)
    # laplacian: gamma = 1 + x/2, U = (x^2, y^2 + x, 0)  ->  d/dx(gamma 2x) = 2 + 2x ; gamma*2 + 1/2
    gamma = 1.0 + 0.5 * C[:, 0]
    U = np.stack([C[:, 0] ** 2, C[:, 1] ** 2 + C[:, 0], np.zeros(len(C))], 1)
    L = O.laplacian_gamma_vector(gamma, U)
    want = np.stack([2.0 + 2.0 * C[:, 0], 2.0 * gamma + 0.5, np.zeros(len(C))], 1)
    np.testing.assert_allclose(L[idx], want[idx], rtol=0, atol=1e-10)
    # convection: uniform flux field a, linear U  ->  a . grad(U)
    a = np.array([0.3, -0.2, 0.5])
    A = np.array([[0.3, -1.0, 2.0], [0.7, 0.2, -0.4], [1.5, 0.0, 0.9]])
    O.field("U")[:] = np.broadcast_to(a, (len(C), 3))
    O.create_phi()
    phi = np.asarray(O.field("phi")).copy()
    D = O.div_phi_vector(phi, C @ A)
    np.testing.assert_allclose(D[idx], np.broadcast_to(a @ A, (idx.size, 3)), rtol=0, atol=1e-11)
    # div(phi, 1) is div(phi) (away from the fixedValue patches, whose face value is the patch's)
    assert np.array_equal(O.div_phi_vector(phi, np.ones((len(C), 3)))[idx, 1], O.div_flux(phi)[idx])
    # the composition: ddtU_f == div(phic, Uc) (the Euler term vanishes), divT == 2 nu laplacian(alphac, Uc)
    O.field("U")[:] = U
    O.create_phi()
    O.field("p")[:] = C @ np.array([2.0, -3.0, 0.5])
    ddtU, gradP, divT, vGrad = O.pimple_pre(0.01, gamma)
    phi = np.asarray(O.field("phi")).copy()
    assert np.array_equal(ddtU, O.div_phi_vector(phi, U))
    assert np.array_equal(divT, (2 * 0.02) * O.laplacian_gamma_vector(gamma, U))
    assert np.array_equal(gradP, O.grad_scalar(np.asarray(O.field("p"))))
    assert np.array_equal(vGrad.reshape(-1, 9), O.grad_vector(U).reshape(-1, 9))
    O.close()


def test_pimple_explicit_stress_and_reconstruct_closed_forms():
    """fvc::div((alpha nu) dev2(T(grad U))) and fvc::reconstruct on fields the discretisation reproduces exactly two cells
    away from the walls (the patch values of grad(U) carry the wall's fixedValue)."""
    m = _mesh3d((9, 8, 7))
    O = port.IcoOracle(m, nu=0.02)
    C = m["C"]
    N = len(C)
    nx, ny, nz = m["n"]
    idx = np.arange(N).reshape(nz, ny, nx)[2:-2, 2:-2, 2:-2].reshape(-1)
    A = np.array([[0.3, -1.0, 2.0], [0.7, 0.2, -0.4], [1.5, 0.0, 0.9]])
    assert np.abs(O.div_dev(np.ones(N), C @ A, 0.02)[idx]).max() < 1e-12          # constant stress: no divergence
    U = np.stack([C[:, 0] * C[:, 1], np.zeros(N), np.zeros(N)], 1)             # T(grad U): xx = y, xy = x ; tr = y
    d = O.div_dev(np.ones(N), U, 0.02)[idx]                                   # div_y = d_x(x) + d_y(-2y/3) = 1/3
    np.testing.assert_allclose(d, np.broadcast_to([0.0, 0.02 / 3.0, 0.0], d.shape), rtol=0, atol=1e-12)
    a = np.array([0.3, -0.2, 0.5])
    O.field("U")[:] = a
    O.create_phi()
    r = O.reconstruct(np.asarray(O.field("phi")).copy())                      # reconstruct(a . Sf) = a
    idx1 = np.arange(N).reshape(nz, ny, nx)[1:-1, 1:-1, 1:-1].reshape(-1)
    np.testing.assert_allclose(r[idx1], np.broadcast_to(a, (idx1.size, 3)), rtol=0, atol=1e-14)
    O.close()


def test_pimple_step_reduces_to_icoFoam_and_conserves_mass():
    """pimpleSolve (pim/UcEqn.H, pEqn.H) with alphac = 1 and no particle sources is icoFoam's equation in another
    pressure-gradient form: from rest the first step is icoFoam's to round-off (the stock cavity log's numbers), later
    steps agree at first order in h.  With a non-uniform void fraction, implicit drag and a source,
    fvc::div(alphacf*phic) closes to the solver tolerance."""
    one = zero = None
    err = []
    for n in (10, 20):
        mc = cavity_mesh(n)
        Oi, Op = port.IcoOracle(mc, nu=0.01), port.IcoOracle(mc, nu=0.01)
        one, zero = np.ones(mc["nCells"]), np.zeros(mc["nCells"])
        dt = 0.005 * 20 / n
        for O_ in (Oi, Op):
            O_.create_phi()
        for it in range(int(round(0.05 / dt))):
            Oi.pre(dt)
            Oi.solve(dt)
            Op.pre(dt)
            Op.pimple_solve(dt, one, zero)
            if it == 0:
                assert np.linalg.norm(Oi.field("U") - Op.field("U")) <= 1e-12 * np.linalg.norm(Oi.field("U"))
                assert np.linalg.norm(Oi.field("p") - Op.field("p")) <= 1e-11 * np.linalg.norm(Oi.field("p"))
                if n == 20:
                    so = Op.stats()
                    assert (sig6(so["p"][0]["final"]), so["p"][0]["iters"]) == CAVITY_LOG[0]["p1"][1:]
                    assert (sig6(so["p"][1]["final"]), so["p"][1]["iters"]) == CAVITY_LOG[0]["p2"][1:]
        err.append(np.linalg.norm(Oi.field("U") - Op.field("U")) / np.linalg.norm(Oi.field("U")))
        assert not np.any(Op.field("U")[:, 2])                                  # the empty direction stays empty
        Oi.close()
        Op.close()
    assert err[1] < 0.04 and err[1] < 0.65 * err[0]
    m = meshgen.hex_box_ldu(10, 8, 6, 1.0, 0.8, 0.6, patches=[("walls", ["xmin", "xmax", "ymin", "ymax", "zmin", "zmax"])])
    O = port.IcoOracle(m, nu=0.01)
    C = m["C"]
    O.field("U")[:] = 0.1 * np.stack([np.sin(3 * C[:, 1]), np.cos(2 * C[:, 0]), 0 * C[:, 0]], 1)
    O.create_phi()
    alpha = 1 - 0.4 * np.exp(-((C - 0.4) ** 2).sum(1) / 0.05)
    O.field("uSource")[:] = np.stack([0 * alpha, -0.8 * (1 - alpha), 0.2 * (1 - alpha)], 1)
    for it in range(3):
        O.pimple_solve(2e-3, alpha, -50.0 * (1 - alpha))
        st = O.stats()
        assert st["sumLocalContErr"] < 5e-9 and abs(st["globalContErr"]) < 5e-9
        assert np.all(np.isfinite(O.field("U"))) and np.abs(O.field("U")).max() < 1.0
    O.close()


def test_pimple_outer_correctors_and_relaxation():
    """PIMPLE outer correctors (pimpleFoamYade.C:91-105) and relaxationFactors (UcEqn.relax() UcEqn.H:13, p.relax() pEqn.H:41)
    of the restatement: (a) the relaxed matrix equals OpenFOAM-6's fvMatrix::relax(alpha) definition written independently
    in numpy from the unrelaxed coefficients; (b) factor 1 on the (diagonally dominant) Euler matrix changes nothing;
    (c) the outer loop converges: successive outer correctors change the fields less and less, and the result does not
    depend on relaxing the intermediate iterations when the final one is unrelaxed and the loop has converged;
    (d) a p factor < 1 with one outer corrector is refused (OpenFOAM: prevIter not stored)."""
    m = meshgen.hex_box_ldu(10, 8, 6, 1.0, 0.8, 0.6, patches=[("inlet", ["xmin"]), ("outlet", ["xmax"]),
                                                                ("walls", ["ymin", "ymax", "zmin", "zmax"])])
    meshgen.set_bc(m, "inlet", bcU=meshgen.BC_FIXED_VALUE, valueU=(0.3, 0, 0), bcP=meshgen.BC_ZERO_GRADIENT)
    meshgen.set_bc(m, "outlet", bcU=meshgen.BC_ZERO_GRADIENT, bcP=meshgen.BC_FIXED_VALUE, valueP=0.0)
    C, N = m["C"], m["nCells"]
    alpha = 1 - 0.4 * np.exp(-((C - 0.4) ** 2).sum(1) / 0.05)
    drag = -50.0 * (1 - alpha)
    src = np.stack([0 * alpha, -0.8 * (1 - alpha), 0.2 * (1 - alpha)], 1)
    U0 = np.tile([0.3, 0.0, 0.0], (N, 1)) + 0.05 * np.stack([np.sin(3 * C[:, 1]), np.cos(2 * C[:, 0]), np.sin(4 * C[:, 2])], 1)
    dt = 5e-3

    def run(steps=1, tight=False, **pc):
        kw = dict(pTol=1e-13, pFinalTol=1e-13, pRelTol=0.0, UTol=1e-13) if tight else {}
        O = port.IcoOracle(m, nu=0.01, **kw)
        O.field("U")[:] = U0
        O.create_phi()
        O.set_pimple_controls(**pc)
        O.field("uSource")[:] = src
        for _ in range(steps):
            O.pimple_solve(dt, alpha, drag, (0.0, -0.5, 0.0))
        out = {k: O.field(k).copy() for k in ("U", "p", "phi", "diagU", "lowerU", "upperU", "sourceU", "icU", "rAU")}
        out["stats"] = O.stats()
        O.close()
        return out

    base = run()
    # (a) fvMatrix::relax(0.7), independently: D = max(|D0 + sum_b max|ic||, sumMagOffDiag)/alpha - sum_b min(ic); S += (D - D0) psi
    rel = run(relaxU=0.7)
    lo, up = m["owner"], m["neighbour"]
    bC = m["bCell"] if "bCell" in m else np.concatenate([np.asarray(pp["faceCells"]) for pp in m["patches"]])
    ic = base["icU"].reshape(-1, 3)
    D0 = base["diagU"]
    sumOff = np.zeros(N)
    np.add.at(sumOff, up, np.abs(base["lowerU"]))
    np.add.at(sumOff, lo, np.abs(base["upperU"]))
    D = D0.copy()
    np.add.at(D, bC, np.abs(ic).max(1))
    D = np.maximum(np.abs(D), sumOff) / 0.7
    np.subtract.at(D, bC, ic.min(1))
    assert np.allclose(rel["diagU"], D, rtol=1e-14, atol=0)
    assert np.allclose(rel["sourceU"], base["sourceU"] + (D - D0)[:, None] * U0, rtol=1e-13, atol=1e-18)
    assert np.array_equal(rel["lowerU"], base["lowerU"]) and np.array_equal(rel["upperU"], base["upperU"])
    assert np.linalg.norm(rel["U"] - base["U"]) > 1e-6 * np.linalg.norm(base["U"])       # (it does something)
    # (b) alpha = 1 on a dominant matrix: D + max|ic| - min(ic) = D for the fixedValue / zeroGradient coefficients
    one = run(relaxU=1.0)
    for k in ("U", "p", "phi"):
        assert np.linalg.norm(one[k] - base[k]) <= 1e-11 * np.linalg.norm(base[k]), k
    # (c) convergence of the outer loop
    seq = [run(tight=True, nOuterCorrectors=n) for n in (1, 2, 3, 6, 7)]
    d = [np.linalg.norm(seq[i + 1]["U"] - seq[i]["U"]) / np.linalg.norm(seq[i]["U"]) for i in range(4)]
    assert d[0] > d[1] and d[3] < 0.05 * d[0], d
    relaxed = run(tight=True, nOuterCorrectors=12, relaxU=0.8, relaxUFinal=1.0, relaxP=0.6, relaxPFinal=1.0)
    plain = run(tight=True, nOuterCorrectors=12)
    for k in ("U", "p"):
        assert np.linalg.norm(relaxed[k] - plain[k]) <= 2e-3 * np.linalg.norm(plain[k]), k
    assert relaxed["stats"]["nPSolves"] == 12 * 2
    # (d)
    with pytest.raises(RuntimeError):
        run(relaxP=0.5)


def test_pimple_UcEqn_matrix_is_the_sum_of_its_explicit_operators():
    """The assembled UcEqn (pim/UcEqn.H:3-11; diag / lower / upper / source of the restatement) applied to an arbitrary
    field W must equal, away from the patches, V times the same terms evaluated with the separately written fvc
    operators:  alpha (W - U0)/dt + div(alphaPhic, W) - (ddt(alpha) + div(alphaPhic)) W - laplacian(alpha nu, W)
                - div((alpha nu) dev2(T(grad U0))) - uSourceDrag W.
    Catches sign / weighting slips in the matrix assembly (Sp terms, the negated explicit stress term, alphacf)."""
    m = _mesh3d((9, 8, 7))
    nu, dt = 0.02, 0.01
    O = port.IcoOracle(m, nu=nu, momentumPredictor=0, nCorrectors=1)
    C = m["C"]
    N, Fi = m["nCells"], m["nInternalFaces"]
    nx, ny, nz = m["n"]
    rng = np.random.default_rng(3)
    U0 = 0.3 * np.stack([np.sin(3 * C[:, 1]) + C[:, 0] ** 2, np.cos(2 * C[:, 0]) * C[:, 2], C[:, 0] * C[:, 1]], 1)
    O.field("U")[:] = U0
    O.create_phi()
    phi0 = np.asarray(O.field("phi")).copy()
    alpha = 1 - 0.4 * np.exp(-((C - C.mean(0)) ** 2).sum(1) / (0.05 * np.ptp(C[:, 0]) ** 2))
    drag = -30.0 * (1 - alpha)
    O.field("uSource")[:] = 0.0
    O.pimple_solve(dt, alpha, drag)
    diag, lower, upper = (np.asarray(O.field(k)).copy() for k in ("diagU", "lowerU", "upperU"))
    source = np.asarray(O.field("sourceU")).copy()
    alphaf, spDiv, divDev = (np.asarray(O.pimple_field(k)).copy() for k in ("alphaf", "spDiv", "divDev"))
    W = rng.standard_normal((N, 3))
    AW = diag[:, None] * W
    np.add.at(AW, m["owner"], upper[:, None] * W[m["neighbour"]])
    np.add.at(AW, m["neighbour"], lower[:, None] * W[m["owner"]])
    V = m["V"][:, None]
    want = V * (alpha[:, None] * (W - U0) / dt + O.div_phi_vector(alphaf * phi0, W) - spDiv[:, None] * W
                - O.laplacian_gamma_vector(alpha * nu, W, gammaB=nu) - divDev - drag[:, None] * W)
    idx = np.arange(N).reshape(nz, ny, nx)[1:-1, 1:-1, 1:-1].reshape(-1)
    got = AW - source
    assert np.abs(got[idx] - want[idx]).max() <= 1e-11 * np.abs(want[idx]).max()
    assert np.abs(spDiv).max() > 1e-3 and np.abs(divDev[idx]).max() > 1e-6      # the terms under test are really there
    O.close()


def test_pimple_hydrostatic_balance_with_fixedFluxPressure():
    """constrainPressure on fixedFluxPressure walls (pim/pEqn.H:21; oracle only so far, the device path still refuses
    the patch type): a closed box under gravity stays at rest, p = g.x, no flux through the walls -- with and without a
    void-fraction blob."""
    m = meshgen.hex_box_ldu(8, 12, 6, 0.8, 1.2, 0.6, patches=[("walls", ["xmin", "xmax", "ymin", "ymax", "zmin", "zmax"])])
    meshgen.set_bc(m, "walls", bcP=meshgen.BC_FIXED_FLUX_PRESSURE)
    O = port.IcoOracle(m, nu=0.01)
    N, Fi, C = m["nCells"], m["nInternalFaces"], m["C"]
    O.create_phi()
    g = (0.0, -9.81, 0.0)
    alpha = np.ones(N)
    for it in range(4):
        if it == 2:
            alpha = 1 - 0.4 * np.exp(-((C - C.mean(0)) ** 2).sum(1) / 0.02)
        O.pimple_solve(1e-3, alpha, -20.0 * (1 - alpha), g)
        p = np.asarray(O.field("p"))
        assert np.abs(O.field("U")).max() < 1e-6                                  # solver tolerance, not a flow
        assert abs(np.polyfit(C[:, 1], p, 1)[0] + 9.81) < 1e-3
        assert np.abs(p - p[0] + 9.81 * (C[:, 1] - C[0, 1])).max() < 1e-3
        assert np.abs(np.asarray(O.field("phi"))[Fi:]).max() < 1e-15
        assert O.stats()["sumLocalContErr"] < 1e-9
    O.close()


@pytest.mark.parametrize("solver", ["ico", "pimple"])
def test_fluid_restatement_reproduces_its_fixture(solver):
    """tests/golden/fluid_*.npz (tests/golden/gen_golden_fluid.py): regression pins of the two fluid-step restatements,
    so that an edit of the oracle cannot silently move the target the CUDA path is compared with."""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("gen_golden_fluid", os.path.join(here, "gen_golden_fluid.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    want = np.load(os.path.join(here, "fluid_%s_channel.npz" % solver))
    got = gen.run(solver)
    assert np.array_equal(got["iters"], want["iters"])
    for k in ("U", "p", "phi"):
        assert np.linalg.norm(got[k] - want[k]) <= 1e-12 * np.linalg.norm(want[k]), k
    np.testing.assert_allclose(got["contErr"], want["contErr"], rtol=1e-6, atol=1e-18)


def _ldu_dense(m, diag, lower, upper):
    N = m["nCells"]
    A = np.zeros((N, N))
    A[np.arange(N), np.arange(N)] = diag
    A[m["neighbour"], m["owner"]] = lower
    A[m["owner"], m["neighbour"]] = upper
    return A


def test_ldu_kernels_against_dense_algebra():
    m = _mesh3d()
    O = port.IcoOracle(m)
    rng = np.random.default_rng(3)
    N, Fi = m["nCells"], m["nInternalFaces"]
    upper = -rng.uniform(0.5, 1.5, Fi)
    lower = -rng.uniform(0.5, 1.5, Fi)
    diag = np.zeros(N)
    np.subtract.at(diag, m["owner"], lower)
    np.subtract.at(diag, m["neighbour"], upper)
    diag += rng.uniform(0.1, 0.2, N)
    psi = rng.standard_normal(N)
    np.testing.assert_allclose(O.amul(diag, lower, upper, psi), _ldu_dense(m, diag, lower, upper) @ psi, rtol=1e-13)
    # symmetric negative-definite (Laplacian-like) system: PCG to tight tolerance == dense solve
    A = _ldu_dense(m, -diag, -upper, -upper)
    b = rng.standard_normal(N)
    for pre in ("DIC", "diagonal", "none"):
        x, perf = O.pcg(-diag, -upper, b, np.zeros(N), tol=1e-14, relTol=0.0, preconditioner=pre)
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-9, atol=1e-11)
        assert 0 < perf["iters"] < 200 and perf["final"] < 1e-14
    # DIC: M = (L + D) D^-1 (D + L^T) with the recurrence D_u = a_uu - sum a_ul^2 / D_l
    Asym = _ldu_dense(m, diag, upper, upper)
    D = diag.copy()
    for f in range(Fi):
        D[m["neighbour"][f]] -= upper[f] ** 2 / D[m["owner"][f]]
    Ls = np.tril(Asym, -1)
    M = (Ls + np.diag(D)) @ np.diag(1.0 / D) @ (np.diag(D) + Ls.T)
    r = rng.standard_normal(N)
    np.testing.assert_allclose(O.dic(diag, upper, r), np.linalg.solve(M, r), rtol=1e-11)
    # symGaussSeidel: converges to the dense solution of the asymmetric, diagonally dominant system
    Aas = _ldu_dense(m, diag, lower, upper)
    x, perf = O.smooth(diag, lower, upper, b, np.zeros(N), tol=1e-13)
    np.testing.assert_allclose(x, np.linalg.solve(Aas, b), rtol=1e-8, atol=1e-10)
    O.close()


def test_channel_step_is_divergence_free_and_bounded():
    """inlet (fixedValue U) / outlet (zeroGradient U, fixedValue p) box: two PISO steps leave a discretely
    divergence-free flux field whose outflow equals the inflow."""
    m = _mesh3d((12, 6, 5))
    O = port.IcoOracle(m, nu=0.01)
    O.field("U")[:] = (0.3, 0.0, 0.0)
    O.create_phi()
    for _ in range(2):
        O.pre(0.01)
        O.solve(0.01)
    st = O.stats()
    assert st["sumLocalContErr"] < 1e-8
    phi = O.field("phi")
    Fi = m["nInternalFaces"]
    sizes = [p["faceCells"].size for p in m["patches"]]
    inflow = phi[Fi:Fi + sizes[0]].sum()
    outflow = phi[Fi + sizes[0]:Fi + sizes[0] + sizes[1]].sum()
    assert inflow < 0 and abs(inflow + outflow) < 1e-7 * abs(inflow)
    assert np.all(np.isfinite(O.field("U"))) and np.abs(O.field("U")).max() < 1.0
    O.close()


def _bcells(m):
    return np.concatenate([p["faceCells"] for p in m["patches"]])


def test_ico_channel_UEqn_and_pEqn_against_independent_routes():
    """An independent pin of the icoFoamYade restatement on a 3-D case with an inlet, an outlet and walls -- what the
    2-D cavity log cannot reach (no reference test exists for the fluid half).
    (1) UEqn (icoFoamYade.C:79-84): the assembled LDU matrix INCLUDING its boundary coefficients, applied to an arbitrary
        field W, equals V ((W - U0)/dt + div(phi, W) - nu laplacian(W)) evaluated with the separately written explicit
        fvc operators, in EVERY cell (boundary cells too: a fixedValue patch contributes its value through the
        boundary coefficients on one side and through the patch value of the fvc operator on the other).
    (2) pEqn (icoFoamYade.C:118-125): the assembled matrix solved by dense LU gives the p the PCG returned (to solver
        tolerance), and div(phi) after the corrector is the pEqn residual (continuity closes to tolerance).
    (3) adjustPhi does not fire with a fixed-value outlet pressure; mass flux in == mass flux out after the step."""
    from tests import cases_fv
    mo, _ = cases_fv.channel(None, (10, 7, 6))
    nu, dt = 0.01, 5e-3
    O = port.IcoOracle(mo, nu=nu, momentumPredictor=1, nCorrectors=2, pTol=1e-12, pRelTol=0.0, pFinalTol=1e-12)
    N, Fi = mo["nCells"], mo["nInternalFaces"]
    U0, p0 = cases_fv.channel_init(mo["C"])
    O.field("U")[:] = U0
    O.field("p")[:] = p0
    O.create_phi()
    phi0 = np.asarray(O.field("phi")).copy()
    O.pre(dt)
    O.solve(dt)
    diag, lower, upper = (np.asarray(O.field(k)).copy() for k in ("diagU", "lowerU", "upperU"))
    source, ic, bc = (np.asarray(O.field(k)).copy() for k in ("sourceU", "icU", "bcU"))
    rng = np.random.default_rng(11)
    W = rng.standard_normal((N, 3))
    AW = diag[:, None] * W
    np.add.at(AW, mo["owner"], upper[:, None] * W[mo["neighbour"]])
    np.add.at(AW, mo["neighbour"], lower[:, None] * W[mo["owner"]])
    bcell = _bcells(mo)
    np.add.at(AW, bcell, ic * W[bcell])
    rhs = source.copy()
    np.add.at(rhs, bcell, bc)
    V = mo["V"][:, None]
    want = V * ((W - U0) / dt + O.div_phi_vector(phi0, W) - O.laplacian_gamma_vector(np.full(N, nu), W, gammaB=nu))
    got = AW - rhs
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    assert len(np.unique(bcell)) > N // 3                                   # the boundary layer is a large part of this mesh
    # (2) pEqn by dense algebra
    dP, uP, sP = (np.asarray(O.field(k)).copy() for k in ("diagP", "upperP", "sourceP"))
    A = _ldu_dense(mo, dP, uP, uP)
    bP = sP.copy()
    # the fixed-value outlet enters through the boundary coefficients: rebuild them from the definition
    rAU = np.asarray(O.field("rAU"))
    off = 0
    for pt in mo["patches"]:
        nf = pt["faceCells"].size
        if pt["bcP"] == meshgen.BC_FIXED_VALUE:
            g = rAU[pt["faceCells"]] * pt["magSf"]
            A[pt["faceCells"], pt["faceCells"]] += g * (-1.0 * pt["deltaCoeffs"])
            np.add.at(bP, pt["faceCells"], -g * (pt["deltaCoeffs"] * pt["valueP"]))
        off += nf
    p_dense = np.linalg.solve(A, bP)
    p = np.asarray(O.field("p"))
    assert np.linalg.norm(p - p_dense) <= 1e-8 * np.linalg.norm(p_dense)
    st = O.stats()
    assert st["sumLocalContErr"] < 1e-12 and abs(st["globalContErr"]) < 1e-13
    # (3) mass balance over the patches
    phi = np.asarray(O.field("phi"))
    assert abs(phi[Fi:].sum()) < 1e-12 * np.abs(phi[Fi:]).sum()
    O.close()


def test_pimple_UcEqn_matrix_with_boundary_cells():
    """The term-by-term check of the assembled UcEqn (above) extended to EVERY cell: the boundary coefficients of the
    restatement's matrix (fixedValue walls: value through the implicit laplacian and the convection; alphac's patches
    hold 1) against the explicit operators with the patch values."""
    m = _mesh3d((9, 8, 7))
    nu, dt = 0.02, 0.01
    O = port.IcoOracle(m, nu=nu, momentumPredictor=0, nCorrectors=1)
    C = m["C"]
    N = m["nCells"]
    rng = np.random.default_rng(5)
    U0 = 0.3 * np.stack([np.sin(3 * C[:, 1]) + C[:, 0] ** 2, np.cos(2 * C[:, 0]) * C[:, 2], C[:, 0] * C[:, 1]], 1)
    O.field("U")[:] = U0
    O.create_phi()
    phi0 = np.asarray(O.field("phi")).copy()
    alpha = 1 - 0.4 * np.exp(-((C - C.mean(0)) ** 2).sum(1) / (0.05 * np.ptp(C[:, 0]) ** 2))
    drag = -30.0 * (1 - alpha)
    O.field("uSource")[:] = 0.0
    O.pimple_solve(dt, alpha, drag)
    diag, lower, upper = (np.asarray(O.field(k)).copy() for k in ("diagU", "lowerU", "upperU"))
    source, ic, bc = (np.asarray(O.field(k)).copy() for k in ("sourceU", "icU", "bcU"))
    alphaf, spDiv, divDev = (np.asarray(O.pimple_field(k)).copy() for k in ("alphaf", "spDiv", "divDev"))
    W = rng.standard_normal((N, 3))
    AW = diag[:, None] * W
    np.add.at(AW, m["owner"], upper[:, None] * W[m["neighbour"]])
    np.add.at(AW, m["neighbour"], lower[:, None] * W[m["owner"]])
    bcell = _bcells(m)
    np.add.at(AW, bcell, ic * W[bcell])
    rhs = source.copy()
    np.add.at(rhs, bcell, bc)
    V = m["V"][:, None]
    want = V * (alpha[:, None] * (W - U0) / dt + O.div_phi_vector(alphaf * phi0, W) - spDiv[:, None] * W
                - O.laplacian_gamma_vector(alpha * nu, W, gammaB=nu) - divDev - drag[:, None] * W)
    got = AW - rhs
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()


@pytest.mark.parametrize("solver", ["ico", "pimple"])
def test_restatements_reproduce_plane_poiseuille_flow_at_second_order(solver):
    """A pin that does not pass through OpenFOAM at all: the developed flow between two plates has the closed form
    u(y) = 6 Um y (1 - y/H)/H, dp/dx = -12 nu Um / H^2.  Both restatements (icoFoamYade's PISO step; pimpleFoamYade's with
    alphac = 1 and no particle sources, whose pressure-gradient form and face fluxes are assembled differently) march a
    uniform inlet flow to the steady state on two grids: profile and pressure gradient converge to the analytic values
    at second order (errors 1.4e-2 -> 3.7e-3 and 2 % -> 0.5 %)."""
    nu, L, H = 0.1, 3.0, 1.0
    errs = []
    for nx, ny in ((30, 10), (60, 20)):
        m = meshgen.hex_box_ldu(nx, ny, 1, L, H, 0.1, patches=[("inlet", ["xmin"]), ("outlet", ["xmax"]), ("walls", ["ymin", "ymax"]),
                                                                 ("frontAndBack", ["zmin", "zmax"])])
        meshgen.set_bc(m, "inlet", bcU=meshgen.BC_FIXED_VALUE, valueU=(1, 0, 0), bcP=meshgen.BC_ZERO_GRADIENT)
        meshgen.set_bc(m, "outlet", bcU=meshgen.BC_ZERO_GRADIENT, bcP=meshgen.BC_FIXED_VALUE, valueP=0.0)
        meshgen.set_bc(m, "frontAndBack", bcU=meshgen.BC_EMPTY, bcP=meshgen.BC_EMPTY)
        O = port.IcoOracle(m, nu=nu)
        O.field("U")[:] = [1.0, 0.0, 0.0]
        O.create_phi()
        h = L / nx
        dt = 0.2 * h / 1.5
        one, zero = np.ones(m["nCells"]), np.zeros(m["nCells"])
        for it in range(int(3.0 / dt)):
            if solver == "ico":
                O.pre(dt)
                O.solve(dt)
            else:
                O.pimple_solve(dt, one, zero)
        U, p = O.field("U").reshape(ny, nx, 3), O.field("p").reshape(ny, nx)
        y = (np.arange(ny) + 0.5) * H / ny
        e_prof = np.abs(U[:, -2, 0] - 6 * y * (1 - y)).max()
        dpdx = (p[:, -3].mean() - p[:, -8].mean()) / (5 * h)
        errs.append((e_prof, abs(dpdx + 12 * nu) / (12 * nu)))
        assert np.abs(U[:, -2, 1]).max() < 2e-3 and not np.any(U[:, :, 2])
        O.close()
    (e0, g0), (e1, g1) = errs
    assert e0 < 2e-2 and e1 < 5e-3 and e1 < 0.35 * e0, errs
    assert g0 < 3e-2 and g1 < 8e-3 and g1 < 0.4 * g0, errs


@pytest.mark.parametrize("solver", ["ico", "pimple"])
def test_restatements_reproduce_the_lid_driven_cavity_benchmark(solver):
    """A second pin outside OpenFOAM, with the convective term at work: the steady lid-driven cavity at Re = 100 against the
    tabulated centre-line velocities of Ghia, Ghia & Shin (J. Comput. Phys. 48, 1982, table I, 129 x 129 multigrid
    solution).  Both restatements, marched to the steady state on 32 x 32 cells, meet all fifteen interior points within
    6e-3 of the lid speed (measured: 3.2e-3; 64 x 64 gives 3.4e-3, at the level of the table's own digits and of the
    linear interpolation between cell centres) and the minimum of u within 1.5 %."""
    y_g = np.array([0.9766, 0.9688, 0.9609, 0.9531, 0.8516, 0.7344, 0.6172, 0.5, 0.4531, 0.2813, 0.1719, 0.1016, 0.0703, 0.0625, 0.0547])
    u_g = np.array([0.84123, 0.78871, 0.73722, 0.68717, 0.23151, 0.00332, -0.13641, -0.20581, -0.21090, -0.15662, -0.10150, -0.06434,
                    -0.04775, -0.04192, -0.03717])
    n = 32
    m = meshgen.hex_box_ldu(n, n, 1, 1.0, 1.0, 0.1, patches=[("movingWall", ["ymax"]), ("fixedWalls", ["xmin", "xmax", "ymin"]),
                                                            ("frontAndBack", ["zmin", "zmax"])])
    meshgen.set_bc(m, "movingWall", valueU=(1, 0, 0))
    meshgen.set_bc(m, "frontAndBack", bcU=meshgen.BC_EMPTY, bcP=meshgen.BC_EMPTY)
    O = port.IcoOracle(m, nu=0.01)
    O.create_phi()
    h = 1.0 / n
    dt = 0.4 * h
    one, zero = np.ones(m["nCells"]), np.zeros(m["nCells"])
    for it in range(int(20.0 / dt)):
        if solver == "ico":
            O.pre(dt)
            O.solve(dt)
        else:
            O.pimple_solve(dt, one, zero)
    U = O.field("U").reshape(n, n, 3)
    O.close()
    y = (np.arange(n) + 0.5) * h
    uc = 0.5 * (U[:, n // 2 - 1, 0] + U[:, n // 2, 0])              # x = 0.5 lies on a face: mean of the two cell columns
    assert np.abs(np.interp(y_g, y, uc) - u_g).max() < 6e-3
    assert abs(uc.min() + 0.2109) < 0.015 * 0.2109 + 2e-3
