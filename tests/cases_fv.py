"""Synthetic fluid cases shared by the FV tests, __graft_entry__.smoke() and bench.py: the same case built
twice -- as the oracle's mesh dict (oracle.meshgen) and as the product's (pkg.box_mesh) -- plus the
oracle-vs-engine comparison of icoFoamYade time steps."""
import numpy as np

from tests import cases

CAVITY_PATCHES = [("movingWall", ["ymax"]), ("fixedWalls", ["xmin", "xmax", "ymin"]), ("frontAndBack", ["zmin", "zmax"])]


def _apply(set_bc, mesh, spec):
    for name, kw in spec.items():
        set_bc(mesh, name, **kw)


def cavity2d(pkg=None, n=20, oracle=True):
    """stock OpenFOAM cavity tutorial: n x n x 1 cells on 0.1 x 0.1 x 0.01, lid U = (1 0 0), empty front/back"""
    spec = dict(movingWall=dict(valueU=(1, 0, 0)), frontAndBack=dict(bcU=2, bcP=2))
    mo = None
    if oracle:
        from oracle import meshgen
        mo = meshgen.hex_box_ldu(n, n, 1, 0.1, 0.1, 0.01, patches=CAVITY_PATCHES)
        _apply(meshgen.set_bc, mo, spec)
    mp = None
    if pkg is not None:
        mp = pkg.box_mesh(n, n, 1, 0.1, 0.1, 0.01, patches=CAVITY_PATCHES)
        _apply(pkg.set_bc, mp, spec)
    return mo, mp


def cavity3d(pkg=None, n=(16, 16, 16), L=(0.1, 0.1, 0.1), oracle=True):
    """lid-driven cavity box (BASELINE config C1 flow): lid = ymax moving in +x, all other walls no-slip.
    oracle=False builds the product's mesh only (bench.py's engine arm must not touch oracle/)."""
    spec = dict(ymax=dict(valueU=(1, 0, 0)))
    mo = None
    if oracle:
        from oracle import meshgen
        mo = meshgen.hex_box_ldu(*n, *L)
        _apply(meshgen.set_bc, mo, spec)
    mp = None
    if pkg is not None:
        mp = pkg.box_mesh(*n, *L)
        _apply(pkg.set_bc, mp, spec)
    return mo, mp


def channel(pkg=None, n=(24, 12, 10), L=(2.0, 1.0, 1.0), Uin=0.3, oracle=True):
    """channel (BASELINE config C2 flow): xmin inlet U = (Uin 0 0), xmax outlet (zeroGradient U, p = 0), walls no-slip"""
    spec = dict(xmin=dict(valueU=(Uin, 0, 0)), xmax=dict(bcU=1, bcP=0, valueP=0.0))
    mo = None
    if oracle:
        from oracle import meshgen
        mo = meshgen.hex_box_ldu(*n, *L)
        _apply(meshgen.set_bc, mo, spec)
    mp = None
    if pkg is not None:
        mp = pkg.box_mesh(*n, *L)
        _apply(pkg.set_bc, mp, spec)
    return mo, mp


def closed_box(pkg=None, n=(16, 16, 16), L=(1.0, 1.0, 1.0), oracle=True):
    """closed box at rest under gravity (the C3 / C5 flow): no-slip walls all round, p fixedFluxPressure
    (constrainPressure, pimpleFoamYade/pEqn.H:21) -- the suspension's drag and buoyancy set the fluid in motion"""
    patches = [("walls", ["xmin", "xmax", "ymin", "ymax", "zmin", "zmax"])]
    mo = None
    if oracle:
        from oracle import meshgen
        mo = meshgen.hex_box_ldu(*n, *L, patches=patches)
        meshgen.set_bc(mo, "walls", bcP=meshgen.BC_FIXED_FLUX_PRESSURE)
    mp = None
    if pkg is not None:
        mp = pkg.box_mesh(*n, *L, patches=patches)
        pkg.set_bc(mp, "walls", bcP=pkg.BC_FIXED_FLUX_PRESSURE)
    return mo, mp


def channel_init(C, Uin=0.3):
    """smooth, divergence-bearing start field so that every operator has work to do"""
    N = C.shape[0]
    U = np.zeros((N, 3))
    U[:, 0] = Uin * (1.0 + 0.2 * np.sin(3.0 * C[:, 1]) * np.cos(2.0 * C[:, 2]))
    U[:, 1] = 0.05 * np.sin(2.0 * C[:, 0])
    U[:, 2] = -0.03 * np.cos(4.0 * C[:, 1])
    p = 0.1 * np.cos(1.5 * C[:, 0]) * np.sin(C[:, 1] + 0.3)
    return U, p


def run_oracle_steps(mo, U, p, dt, nsteps, nu, ctl=None, source_fn=None):
    from oracle import port
    O = port.IcoOracle(mo, nu=nu, **(ctl or {}))
    O.field("U")[:] = U
    O.field("p")[:] = p
    O.create_phi()
    hist = []
    for it in range(nsteps):
        O.pre(dt)
        if source_fn is not None:
            O.field("uSource")[:] = source_fn(it, O.field("U").copy(), O.field("vGrad").copy())
        O.solve(dt)
        hist.append(O.stats())
    out = dict(U=O.field("U").copy(), p=O.field("p").copy(), phi=O.field("phi").copy(), vGrad=O.field("vGrad").copy(),
               rAU=O.field("rAU").copy(), HbyA=O.field("HbyA").copy(), phiHbyA=O.field("phiHbyA").copy(), stats=hist)
    O.close()
    return out


def run_engine_steps(pkg, mp, U, p, dt, nsteps, nu, ctl=None, source_fn=None, engine=None):
    E = engine or pkg.Engine(mp)
    assert E.fv_supported(), E.L.fy_last_error(E.h).decode()
    E.set_piso_controls(nu=nu, **(ctl or {}))
    E.upload("U", U)
    E.upload("p", p)
    E.create_phi()
    hist = []
    for it in range(nsteps):
        E.ico_pre(dt)
        if source_fn is not None:
            E.upload("uSource", source_fn(it, E.download("U"), E.download("vGrad")))
        E.ico_solve(dt)
        hist.append(E.ico_stats())
    out = dict(U=E.download("U"), p=E.download("p"), phi=E.download("phi"), vGrad=E.download("vGrad"),
               rAU=E.fv_get("rAU"), HbyA=E.fv_get("HbyA"), phiHbyA=E.fv_get("phiHbyA"), stats=hist)
    if engine is None:
        E.close()
    return out


def compare_fluid(o, e, tol=cases.TOL, iters_exact=True):
    errs = {}
    for k in ("vGrad", "rAU", "HbyA", "phiHbyA", "U", "p", "phi"):
        errs[k] = cases.rel_l2(e[k], o[k])
        assert errs[k] <= tol, "%s: relative L2 %.3e > %.1e" % (k, errs[k], tol)
    for so, se in zip(o["stats"], e["stats"]):
        for j in range(3):
            if iters_exact:
                assert so["U"][j]["iters"] == se["U"][j]["iters"], "U%d iterations %s vs %s" % (j, so["U"][j], se["U"][j])
            np.testing.assert_allclose(se["U"][j]["initial"], so["U"][j]["initial"], rtol=1e-9, atol=1e-300)
        assert so["nPSolves"] == se["nPSolves"]
        for po, pe in zip(so["p"], se["p"]):
            if iters_exact:
                assert po["iters"] == pe["iters"], "p iterations %s vs %s" % (po, pe)
            np.testing.assert_allclose(pe["initial"], po["initial"], rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(se["CoNum"], so["CoNum"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(se["meanCoNum"], so["meanCoNum"], rtol=1e-12, atol=1e-300)
    return errs


def smoke_check_fv(pkg, verbose=False):
    """one small icoFoamYade fluid step (16^3 lid-driven cavity) on cuda:0 against the oracle"""
    mo, mp = cavity3d(pkg, (16, 16, 16))
    N = mo["nCells"]
    U, p = np.zeros((N, 3)), np.zeros(N)
    o = run_oracle_steps(mo, U, p, 0.005, 2, 0.01)
    e = run_engine_steps(pkg, mp, U, p, 0.005, 2, 0.01)
    errs = compare_fluid(o, e)
    if verbose:
        print("smoke fluid (16^3 cavity, 2 PISO steps): p iterations %s, rel-L2 %s" % (
            [q["iters"] for q in e["stats"][-1]["p"]], {k: "%.1e" % v for k, v in errs.items()}))
