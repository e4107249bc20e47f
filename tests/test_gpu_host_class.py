"""GPU: the OpenFOAM-side host class (yade-openfoam-coupling_b200/host/FoamYadeB200.H -> C ABI -> sm_100a kernels)
against the unmodified reference class, both driven by the SAME fake Yade peer (oracle/ref_harness.cpp):
identical message sequence on the wire, identical found flags, forces and fields within 1e-10."""
import numpy as np
import pytest

from oracle import meshgen, ref
from tests import cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref.available() and ref.host_available()), reason="harness libraries not built")]


def _run(host, gaussian, n_yade, mo, pd, fields, steps=2, realloc=False):
    L = ref.host_lib() if host else ref.lib()
    L.ref_clear_trace()
    L.ref_set_logging(1)
    R = ref.RefFoamYade(mo, gaussian, n_yade, host=host)
    R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
    for k, v in fields.items():
        R.field(k)[:] = v
    out = []
    for it in range(steps):
        if realloc and it > 0:
            R.realloc_fields()              # the solver's `field = tmp` assignments: new addresses every step
        found, force = R.step(1e-3, pd, yade_dt=2.5e-4)
        out.append(dict(found=found.copy(), force=force.copy(), uSource=R.field("uSource").copy(),
                        alpha=R.field("alpha").copy(), uSourceDrag=R.field("uSourceDrag").copy(),
                        uParticle=R.field("uParticle").copy(), dts=R.dts()))
        if realloc:
            R.realloc_fields()
        R.set_source_zero()
    tr = R.trace()
    L.ref_set_logging(0)
    R.close()
    return tr, out


@pytest.mark.parametrize("gaussian", [True, False])
@pytest.mark.parametrize("n_yade", [1, 3])
def test_host_class_is_a_drop_in(pkg, gaussian, n_yade):
    n = 16
    mo = meshgen.hex_box(n, n, n)
    pd = cases.particles(400, 5, radius=0.1 / n, moving=True)
    pd[:7, :3] += 2.0                         # a few particles outside the mesh: reported as not found
    f = cases.fields_for(mo["C"])
    fields = dict(U=f["U"], gradP=f["gradP"], divT=f["divT"], vGrad=f["vGrad"])
    tr_ref, o_ref = _run(False, gaussian, n_yade, mo, pd, fields)
    tr_host, o_host = _run(True, gaussian, n_yade, mo, pd, fields)
    assert tr_host == tr_ref                   # same collectives, same order, same sizes, same tags
    for a, b in zip(o_ref, o_host):
        assert np.array_equal(a["found"], b["found"])
        assert cases.rel_l2(b["force"], a["force"]) <= cases.TOL
        for k in ("uSource", "alpha", "uSourceDrag", "uParticle"):
            assert cases.rel_l2(b[k], a[k]) <= cases.TOL, k
        assert a["dts"] == b["dts"]


@pytest.mark.parametrize("gaussian", [True, False])
def test_host_class_follows_reallocated_fields(pkg, gaussian):
    """OpenFOAM solvers reassign their fields from tmps every time step, which moves the internal field's storage;
    the host class must hand the engine the current addresses on every call (it once cached them at construction)."""
    n = 12
    mo = meshgen.hex_box(n, n, n)
    pd = cases.particles(300, 9, radius=0.1 / n, moving=True)
    f = cases.fields_for(mo["C"])
    fields = dict(U=f["U"], gradP=f["gradP"], divT=f["divT"], vGrad=f["vGrad"])
    _, o_ref = _run(False, gaussian, 1, mo, pd, fields, steps=3, realloc=True)
    _, o_host = _run(True, gaussian, 1, mo, pd, fields, steps=3, realloc=True)
    for a, b in zip(o_ref, o_host):
        assert np.array_equal(a["found"], b["found"])
        assert cases.rel_l2(b["force"], a["force"]) <= cases.TOL
        for k in ("uSource", "alpha", "uSourceDrag", "uParticle"):
            assert cases.rel_l2(b[k], a[k]) <= cases.TOL, k


@pytest.mark.parametrize("solver,gaussian", [("ico", False), ("ico", True), ("pimple", True)])
def test_cpp_loop_drivers_run_the_fluid_step_on_the_device(pkg, solver, gaussian):
    """host/icoFoamYadeB200.H: the loop bodies of icoFoamYade.C:65-149 / pimpleFoamYade.C:65-110 in C++ over the C ABI,
    driven through the host class (which hands the engine the fvMesh's LDU addressing, face geometry, patches and the
    patch types of U and p, so that fy_fv_supported() holds) with the fake Yade peer on the wire -- three time steps
    against (unmodified reference operator + oracle fluid step)."""
    from oracle import port
    from tests import cases_fv
    n = 16
    mo, _ = cases_fv.cavity3d(None, (n, n, n), (1.0, 1.0, 1.0))
    nu, dt = (0.01, 5e-3) if solver == "ico" else (1e-3, 2e-3)
    N = mo["nCells"]
    U0 = 0.2 * cases.fields_for(mo["C"])["U"]
    pd = cases.particles(800, 3, radius=0.1 / n, moving=True)
    # oracle side
    O = port.IcoOracle(mo, nu=nu)
    O.field("U")[:] = U0
    O.create_phi()
    R = ref.RefFoamYade(mo, gaussian)
    R.set_properties(cases.RHOP, cases.RHOF, nu)
    # product side: host class + C++ loop body
    H = ref.RefFoamYade(mo, gaussian, host=True)
    H.field("U")[:] = U0
    H.set_fv_mesh(mo)
    H.set_properties(cases.RHOP, cases.RHOF, nu)
    for step in range(3):
        if solver == "ico":
            O.pre(dt)
            R.field("U")[:] = O.field("U")
            R.field("vGrad")[:] = O.field("vGrad")
            if gaussian:
                R.field("gradP")[:] = 0.0
            fo, Fo = R.step(dt, pd, pieces=True)
            O.field("uSource")[:] = R.field("uSource")
            O.solve(dt)
        else:
            ddtU, gradP, divT, vGrad = O.pimple_pre(dt, R.field("alpha").reshape(N))
            for k, v in (("U", O.field("U")), ("ddtU", ddtU), ("gradP", gradP), ("divT", divT), ("vGrad", vGrad)):
                R.field(k)[:] = v.reshape(R.field(k).shape)
            fo, Fo = R.step(dt, pd, pieces=True)
            O.field("uSource")[:] = R.field("uSource")
            O.pimple_solve(dt, R.field("alpha").reshape(N).copy(), R.field("uSourceDrag").reshape(N))
        R.set_source_zero()
        fe, Fe, lg = H.fluid_step(solver, dt, pd)
        H.download_fluid()
        assert np.array_equal(fo, fe)
        assert cases.rel_l2(Fe, Fo) <= cases.TOL
        assert lg["p_iters"] == [q["iters"] for q in O.stats()["p"]]
        assert cases.rel_l2(H.field("U"), O.field("U")) <= cases.TOL
        assert cases.rel_l2(H.field("p"), O.field("p")) <= cases.TOL
    H.close()
    R.close()
    O.close()


def test_host_class_options_beyond_the_reference_loop(pkg):
    """FoamYadeB200::setGaussianOptions / setPimpleControls through the C++ class and the pimpleFoamYade loop body: full-support
    Gaussian cell sets + addedMassForce + Gaussian torque on the wire, two outer PIMPLE correctors with relaxation -- three
    time steps against (the unmodified reference's own functions fed with the full cell sets + the oracle's pimpleSolve)."""
    from oracle import port
    from tests import cases_fv
    n = 16
    mo, _ = cases_fv.cavity3d(None, (n, n, n), (1.0, 1.0, 1.0))
    nu, dt = 1e-3, 2e-3
    N = mo["nCells"]
    U0 = 0.2 * cases.fields_for(mo["C"])["U"]
    pd = cases.particles(300, 3, radius=0.1 / n, moving=True)
    pd[:, 0:3] = 0.1 + 0.8 * pd[:, 0:3]
    pim = dict(nOuterCorrectors=2, relaxU=0.8, relaxUFinal=1.0, relaxP=0.5, relaxPFinal=1.0)
    O = port.IcoOracle(mo, nu=nu)
    O.set_pimple_controls(**pim)
    O.field("U")[:] = U0
    O.create_phi()
    R = ref.RefFoamYade(mo, True)
    R.set_properties(cases.RHOP, cases.RHOF, nu)
    R.set_gaussian_options(True, True, True)
    H = ref.RefFoamYade(mo, True, host=True)
    H.field("U")[:] = U0
    H.set_fv_mesh(mo)
    H.host_gaussian_options(True, True, True)          # before setScalarProperties: applied when the engine is created
    H.host_pimple_controls(**pim)
    H.set_properties(cases.RHOP, cases.RHOF, nu)
    for step in range(3):
        ddtU, gradP, divT, vGrad = O.pimple_pre(dt, R.field("alpha").reshape(N))
        for k, v in (("U", O.field("U")), ("ddtU", ddtU), ("gradP", gradP), ("divT", divT), ("vGrad", vGrad)):
            R.field(k)[:] = v.reshape(R.field(k).shape)
        fo, Fo = R.step(dt, pd, pieces=True)
        O.field("uSource")[:] = R.field("uSource")
        O.pimple_solve(dt, R.field("alpha").reshape(N).copy(), R.field("uSourceDrag").reshape(N))
        R.set_source_zero()
        fe, Fe, lg = H.fluid_step("pimple", dt, pd)
        H.download_fluid()
        assert np.array_equal(fo, fe)
        assert np.any(Fo[:, 3:6]) and cases.rel_l2(Fe, Fo) <= cases.TOL
        assert len(lg["p_iters"]) == 4 and lg["p_iters"] == [q["iters"] for q in O.stats()["p"]]
        assert cases.rel_l2(H.field("U"), O.field("U")) <= cases.TOL
        assert cases.rel_l2(H.field("p"), O.field("p")) <= cases.TOL
    H.close()
    R.close()
    O.close()


@pytest.mark.parametrize("gaussian", [True, False])
def test_batched_wire_mode(pkg, gaussian):
    """F1: with batchedWire the serial-Yade exchange is one message per direction and step -- Bcast n, Bcast records,
    ONE Allreduce(i32[n], MAX), ONE Allreduce(f64[6n], SUM) (Gaussian) or ONE Send(f64[6n], tag 1005) (point force), dt
    -- instead of the reference's n + 6n one-element collectives; the replies are those of the legacy protocol."""
    n, P = 12, 250
    mo = meshgen.hex_box(n, n, n)
    pd = cases.particles(P, 17, radius=0.1 / n, moving=True)
    pd[:5, :3] += 2.0
    f = cases.fields_for(mo["C"])
    fields = dict(U=f["U"], gradP=f["gradP"], divT=f["divT"], vGrad=f["vGrad"])
    tr_legacy, o_legacy = _run(True, gaussian, 1, mo, pd, fields, steps=1)
    L = ref.host_lib()
    L.ref_clear_trace()
    L.ref_set_logging(1)
    R = ref.RefFoamYade(mo, gaussian, 1, host=True)
    R.set_batched_wire(True)
    R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
    for k, v in fields.items():
        R.field(k)[:] = v
    found, force = R.step(1e-3, pd, yade_dt=2.5e-4)
    tr = R.trace()
    L.ref_set_logging(0)
    R.close()
    assert np.array_equal(found, o_legacy[0]["found"])
    assert np.array_equal(force, o_legacy[0]["force"])
    step = [l for l in tr if not l.startswith("Wait")]
    allred = [l for l in step if l.startswith("Allreduce")]
    sends = [l for l in step if l.startswith("Send") and "tag=1005" in l]
    assert ("Allreduce i32[%d]" % P) in allred[0]
    if gaussian:
        assert len(allred) == 2 and ("Allreduce f64[%d]" % (6 * P)) in allred[1] and not sends
    else:
        assert len(allred) == 1 and len(sends) == 1 and ("Send f64[%d]" % (6 * P)) in sends[0]
    n_legacy = len([l for l in tr_legacy if l.startswith(("Allreduce", "Send"))])
    n_batched = len([l for l in step if l.startswith(("Allreduce", "Send"))])
    assert n_batched <= 4 and n_legacy >= P
