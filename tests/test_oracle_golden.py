"""CPU: the oracle (the unmodified reference coupling code behind oracle/ref_harness.cpp) against the
committed fixtures and the known answers recorded in SURVEY.md section 8(c)."""
import os

import numpy as np
import pytest

from oracle import meshgen, ref
from tests import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (make -C oracle ref)")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _check_against_fixture(out, g, gaussian):
    assert np.array_equal(out["found"], g["found"].astype(np.int32))
    assert np.array_equal(out["force"], g["force"])
    t = g["touched"]
    for k in ("uSource", "uSourceDrag", "alpha", "uParticle"):
        assert np.array_equal(out[k][t], g[k]), k
    rest = np.ones(out["alpha"].shape[0], dtype=bool)
    rest[t] = False
    assert np.all(out["alpha"][rest] == 1.0) and not np.any(out["uSource"][rest])
    if gaussian:
        assert np.array_equal(out["cnt"], g["cnt"].astype(np.int32))
        m = np.arange(12)[None, :] < out["cnt"][:, None]
        assert np.array_equal(np.where(m, out["ids"], -1), g["ids"])
        assert cases.list_hash(out["cnt"], np.where(m, out["ids"], -1)) == int(g["list_hash"])


@pytest.mark.parametrize("name", ["c1_gauss_moving", "c1_point_moving", "c1_gauss_parallel3", "c1_point_parallel3",
                                  "n16_gauss_dense"])
def test_ref_reproduces_fixture(name):
    g = _load(name)
    n, P, seed = int(g["n"]), int(g["P"]), int(g["seed"])
    gaussian = bool(g["gaussian"])
    mo = meshgen.hex_box(n, n, n)
    pd = cases.particles(P, seed, radius=0.1 / n, moving=bool(g["moving"]))
    out = cases.run_reference_step(mo, cases.fields_for(mo["C"]), pd, gaussian, n_yade=int(g["n_yade"]))
    _check_against_fixture(out, g, gaussian)


def test_unmodified_driver_equals_pieces_with_dense_accumulate():
    """FoamYade::setParticleAction untouched (quadratic scan) == the public pieces with the order-preserving
    dense accumulate, bit for bit, on a case with heavy cell overlap."""
    n, P = 16, 4000
    mo = meshgen.hex_box(n, n, n)
    flds = cases.fields_for(mo["C"])
    pd = cases.particles(P, 5, radius=0.1 / n, moving=True)
    outs = []
    for pieces, dense in ((False, False), (True, False), (True, True)):
        R = ref.RefFoamYade(mo, True)
        R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
        for k in ("U", "gradP", "divT"):
            R.field(k)[:] = flds[k]
        found, force = R.step(1e-3, pd, pieces=pieces, truncate12=False, dense=dense)
        outs.append((found.copy(), force.copy(), R.field("alpha").copy(), R.field("uParticle").copy(),
                     R.field("uSource").copy(), R.field("uSourceDrag").copy()))
        R.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    assert outs[0][2].min() < 0.999      # the case does overlap


def test_survey_known_answers_c1():
    """32^3 / 1000 particles / seed 42 / static particles, Gaussian branch."""
    g = _load("c1_gauss_static")
    n, P = 32, 1000
    mo = meshgen.hex_box(n, n, n)
    R = ref.RefFoamYade(mo, True)
    R.set_properties(2500.0, 1000.0, 1e-6)
    C = mo["C"]
    U = R.field("U")
    U[:, 0] = np.sin(6.28 * C[:, 1])
    U[:, 1] = 0.1 * np.cos(6.28 * C[:, 0])
    U[:, 2] = 0.05
    R.field("gradP")[:] = (0.1, 0.0, 0.0)
    pd = cases.particles(P, 42, radius=0.1 / n)
    assert pd[0, 0] == 0.75515553295453897 and pd[0, 1] == 0.63903139385469743 and pd[0, 2] == 0.7521452007480266
    R.L.ref_set_logging(1)
    found, F = R.step(1e-3, pd)
    tr = R.trace()
    R.L.ref_set_logging(0)
    cnt, ids = R.locate(pd[:, :3])
    assert int((found == 1).sum()) == 1000 and int(cnt.sum()) == 5231
    assert ids[0, :3].tolist() == [25240, 25272, 25241] and cnt[0] == 3
    assert ids[1, :2].tolist() == [3972, 4967] and cnt[1] == 2
    assert ids[2, :4].tolist() == [8562, 8594, 8595, 7507] and cnt[2] == 4
    # the whole cell-id stream as one number (FNV-1a-64 over the ids as u32 words, 0xffffffff after each particle).  The
    # survey quotes 3010d9a434ff91c7 for its own hashing code, which it does not give and the stated recipe does not
    # reproduce (every other known answer of SURVEY.md 8(c) -- pair count, lists, weights, forces, field sums -- does);
    # the value below is this recipe on the reference's lists and pins them all at once.
    assert ref.fnv1a_lists(cnt, ids) == 0xcb6dab77c8744071
    np.testing.assert_allclose(F[0, :3], [-2.7627465786243748e-05, 5.1855587194777599e-07, 1.7114769700056499e-06], rtol=1e-14)
    np.testing.assert_allclose(F[1, :3], [-1.8041716856012765e-05, 1.6363090206990569e-06, 1.3900794444619142e-06], rtol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(R.field("uSource"), axis=1).sum(), 0.00041887902047863922, rtol=1e-12)
    np.testing.assert_allclose(R.field("uSourceDrag").sum(), -0.36068139976006941, rtol=1e-12)
    np.testing.assert_allclose(R.field("alpha").min(), 0.98934923000575281, rtol=1e-14)
    np.testing.assert_allclose(np.abs(F[:, :3]).sum(), 0.024338500117913289, rtol=1e-12)
    # this repo's divT/vGrad fields are zero here, so the fixture made with fields_for() differs only there
    assert np.array_equal(g["found"].astype(np.int32), found)
    # per-step message counts of the serial protocol
    c = R.counts()
    assert (c["bcast"], c["allreduce"], c["send"]) == (3, 7 * P, 1)
    assert tr[0].startswith("Bcast i32[1] peer=0") and tr[1].startswith("Bcast f64[10000] peer=0")
    R.close()


def test_interp_constants():
    """initFields (FoamYade.C:69-72): interpRange = 4*cbrt(V0) is NOT exactly 4h at 128^3."""
    mo = meshgen.hex_box(128, 128, 128)
    # constants do not need the tree; use a tiny mesh with the same V0 to keep this fast
    small = meshgen.hex_box(2, 2, 2, lx=2 / 128, ly=2 / 128, lz=2 / 128)
    assert small["V"][0] == mo["V"][0]
    R = ref.RefFoamYade(small, True)
    c = R.constants()
    R.close()
    assert c["interpRange"] == 0.031250000000000007
    assert c["sigmaInterp"] == c["interpRange"] * 0.42460


def test_analytic_stokes_drag_and_torque():
    """Point-force branch: F = 3 pi d nu rho_f (U - u_p); T = pi d^3 (omega_f - omega_p) nu rho_f
    (FoamYade.C:437-453) in a simple shear U = (G y, 0, 0)."""
    n = 8
    mo = meshgen.hex_box(n, n, n)
    R = ref.RefFoamYade(mo, False)
    rhoF, nu, G = 1000.0, 1e-3, 2.5
    R.set_properties(2500.0, rhoF, nu)
    R.field("U")[:, 0] = G * mo["C"][:, 1]
    R.field("vGrad")[:, 3] = G           # yx = dUx/dy
    pd = np.zeros((1, 10))
    pd[0, :3] = (0.33, 0.71, 0.52)
    pd[0, 3:6] = (0.1, -0.2, 0.05)
    pd[0, 6:9] = (0.0, 0.3, -1.0)
    r = 0.01
    pd[0, 9] = r
    found, F = R.step(1e-3, pd)
    cell = R.L.ref_find_cell(R.h, pd[0, :3].ctypes.data_as(ref._dp))
    d = 2 * r
    Uc = np.array([G * mo["C"][cell, 1], 0.0, 0.0])
    np.testing.assert_allclose(F[0, :3], 3 * np.pi * d * nu * rhoF * (Uc - pd[0, 3:6]), rtol=1e-13)
    omega_f = np.array([0.0, 0.0, G])    # (zy - yz, zx - xz, yx - xy)
    np.testing.assert_allclose(F[0, 3:], np.pi * d ** 3 * (omega_f - pd[0, 6:9]) * nu * rhoF, rtol=1e-13, atol=1e-18)
    # reaction on the fluid: uSource[cell] = -F / (V rho_f)
    np.testing.assert_allclose(R.field("uSource")[cell], -F[0, :3] / (mo["V"][cell] * rhoF), rtol=1e-13)
    R.close()


def test_void_fraction_conservation():
    """sum_cells (1 - alpha) V == sum_p V_p when nothing is clamped (weights normalised per particle)."""
    n, P = 16, 500
    mo = meshgen.hex_box(n, n, n)
    pd = cases.particles(P, 11, radius=0.05 / n)
    out = cases.run_reference_step(mo, cases.fields_for(mo["C"]), pd, True)
    vp = np.pi * (2 * pd[:, 9]) ** 3 / 6.0
    lhs = ((1.0 - out["alpha"]) * mo["V"]).sum()
    np.testing.assert_allclose(lhs, vp[out["found"] == 1].sum(), rtol=1e-9)


def test_hex_mesh_first_cell_is_containing_cell():
    n, P = 16, 3000
    mo = meshgen.hex_box(n, n, n)
    R = ref.RefFoamYade(mo, True)
    xyz = cases.particles(P, 3, radius=0.01)[:, :3]
    cnt, ids = R.locate(xyz)
    cell = np.floor(xyz * n).astype(np.int64)
    cid = cell[:, 0] + n * (cell[:, 1] + n * cell[:, 2])
    ok = cnt > 0
    assert np.array_equal(ids[ok, 0], cid[ok])
    assert (~ok).sum() < 10           # particles whose nearest centre is the tree root are lost (H4)
    R.close()


def test_full_support_lists_and_dormant_forces_of_the_harness():
    """SURVEY 8(f)3 oracle: the harness feeds the reference's OWN weight / force functions with every cell inside the k-d
    search bound (ref_set_gaussian_options).  Checked here without a GPU: (a) an interior particle's support is the
    4/3 pi (sqrt(1.25) 4 h)^3 = 374.6-cell ball and contains the trail list; (b) the void fraction conserves the particle
    volume (the normalised weights sum to one); (c) addedMassForce / the Gaussian torque equal their closed forms
    (FoamYade.C:392-413, 467-478) on uniform fields, where every weighted average is the field value itself."""
    from oracle import meshgen, ref
    from tests import cases
    n, P = 24, 400
    mo = meshgen.hex_box(n, n, n)
    C, V = mo["C"], mo["V"]
    pd = cases.particles(P, 5, radius=0.1 / n, moving=True)
    pd[:, 0:3] = 0.25 + 0.5 * pd[:, 0:3]                  # supports fully inside the box
    N = C.shape[0]
    dt, nu = 1e-3, cases.NU
    ddtU0, vg = np.array([0.3, -0.2, 0.5]), np.array([0.0, 0.1, 0.2, 0.3, 0.0, 0.4, 0.5, 0.6, 0.0])
    out = {}
    for full in (False, True):
        R = ref.RefFoamYade(mo, True)
        R.set_properties(cases.RHOP, cases.RHOF, nu)
        R.set_gaussian_options(full, True, True)
        R.field("U")[:] = np.tile([0.1, 0.0, 0.0], (N, 1))
        R.field("ddtU")[:] = np.tile(ddtU0, (N, 1))
        R.field("vGrad")[:] = np.tile(vg, (N, 1)).reshape(R.field("vGrad").shape)
        f, F = R.step(dt, pd, pieces=True)
        cnt, ids = R.lists(P)
        alpha = R.field("alpha").reshape(N).copy()
        # the same step without the dormant forces: their contribution is the difference
        R.set_source_zero()
        R.set_gaussian_options(full, False, False)
        f0, F0 = R.step(dt, pd, pieces=True)
        out[full] = dict(cnt=cnt, ids=ids, F=F, F0=F0, alpha=alpha)
        R.close()
        vol = np.pi * (2 * pd[:, 9]) ** 3 / 6
        assert abs(((1 - alpha) * V).sum() - vol.sum()) <= 1e-12 * vol.sum()            # (b)
        # (c) f = (vol/k) (ddtU - u_p/dt) rhoP ; T = pi d^3 (w_f - w_p) nu rhoF, w_f = (yz-zy, zx-xz, yx-xy)
        k = cnt.astype(float)
        am = (vol / k)[:, None] * (ddtU0[None, :] - pd[:, 3:6] / dt) * cases.RHOP
        assert np.allclose(F[:, 0:3] - F0[:, 0:3], am, rtol=1e-9, atol=1e-9 * np.abs(am).max())
        wf = np.array([vg[5] - vg[7], vg[6] - vg[2], vg[3] - vg[1]])
        T = np.pi * (2 * pd[:, 9:10]) ** 3 * (wf[None, :] - pd[:, 6:9]) * nu * cases.RHOF
        assert np.allclose(F[:, 3:6], T, rtol=1e-10, atol=0) and not np.any(F0[:, 3:6])
    cf = out[True]["cnt"]
    assert cf.min() >= 340 and cf.max() <= 410 and abs(cf.mean() - 374.6) < 6                       # (a)
    assert np.all(out[False]["cnt"] <= 12)


@pytest.mark.parametrize("name", ["n16_gauss_full_support", "n16_gauss_full_support_dormant", "n16_gauss_trail_dormant"])
def test_f3_fixtures_are_reproduced_by_the_reference_harness(name):
    """tests/golden/gen_golden.py f3: the committed full-support / dormant-force vectors are what the reference's own
    functions give today, bit for bit (they travel to the GPU box, /root/reference does not)."""
    import os
    from oracle import meshgen, ref
    from tests import cases
    from tests.golden.gen_golden import ddtU_of
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    n, P = int(g["n"]), int(g["P"])
    mo = meshgen.hex_box(n, n, n)
    flds = cases.fields_for(mo["C"])
    flds["ddtU"] = ddtU_of(mo["C"])
    pd = cases.particles(P, int(g["seed"]), radius=0.1 / n, moving=True)
    pd[:, 0:3] = 0.05 + 0.9 * pd[:, 0:3]
    R = ref.RefFoamYade(mo, True)
    R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
    R.set_gaussian_options(bool(g["full"]), bool(g["added_mass"]), bool(g["torque"]))
    for k in ("U", "gradP", "divT", "vGrad", "ddtU"):
        R.field(k)[:] = flds[k].reshape(R.field(k).shape)
    found, force = R.step(1e-3, pd, yade_dt=5e-4, pieces=True, truncate12=True, dense=True)
    cnt, _ = R.lists(P)
    assert np.array_equal(found, g["found"]) and np.array_equal(cnt, g["cnt"]) and np.array_equal(force, g["force"])
    for k in ("uSource", "uSourceDrag", "alpha", "uParticle"):
        assert np.array_equal(R.field(k), g[k].reshape(R.field(k).shape)), k
    R.close()
