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
