"""GPU: the sm_100a coupling kernels, called through the C ABI (include/fycuda.h), against
 (a) the committed fixtures made from the unmodified reference, and
 (b) the unmodified reference run live (oracle/_ref/libfoamyade_ref.so travels to the GPU box).
Bar: cell lists / found flags bit-exact; forces and source fields within 1e-10 relative L2 (fp64;
the only difference is the order of the atomic per-cell additions)."""
import os

import ctypes as C

import numpy as np
import pytest

from oracle import meshgen, ref
from tests import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = cases.TOL


def _scatter(g, N, k, w):
    a = np.zeros((N, w)) if w > 1 else np.zeros(N)
    if k == "alpha":
        a[:] = 1.0
    a[g["touched"]] = g[k]
    return a


@pytest.mark.parametrize("name", ["c1_gauss_static", "c1_gauss_moving", "c1_point_moving", "c1_gauss_parallel3",
                                  "c1_point_parallel3", "n16_gauss_dense"])
def test_engine_matches_fixture(pkg, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    n, P, seed, gaussian, n_yade = int(g["n"]), int(g["P"]), int(g["seed"]), bool(g["gaussian"]), int(g["n_yade"])
    mp = pkg.box_mesh(n, n, n, faces=False)
    flds = cases.fields_for(mp["C"])
    pd = cases.particles(P, seed, radius=0.1 / n, moving=bool(g["moving"]))
    E = pkg.Engine(mp)
    E.set_properties(cases.RHOP, cases.RHOF, cases.NU, gaussian)
    for k in ("U", "gradP", "divT", "vGrad"):
        E.upload(k, flds[k])
    if n_yade == 1:
        found, force = E.set_particle_action(1e-3, pd)
        cnt, ids = (E.last_lists(P)[:2]) if gaussian else (None, None)
    else:
        # parallel-Yade: one buffer per worker rank, processed in rank order (FoamYade.C:612-628)
        E.coupling_begin(1e-3)
        found, force, cnt, ids = [], [], [], []
        W = n_yade - 1
        for w in range(W):
            lo, hi = P * w // W, P * (w + 1) // W
            f, F = E.coupling_proc(pd[lo:hi])
            found.append(f)
            force.append(F)
            if gaussian:
                c, i, _ = E.last_lists(hi - lo)
                cnt.append(c)
                ids.append(i)
        E.coupling_end()
        found, force = np.concatenate(found), np.concatenate(force)
        if gaussian:
            cnt, ids = np.concatenate(cnt), np.concatenate(ids)
    assert np.array_equal(found, g["found"].astype(np.int32))
    if gaussian:
        assert np.array_equal(cnt, g["cnt"].astype(np.int32))
        assert np.array_equal(ids, g["ids"])
        assert cases.list_hash(cnt, ids) == int(g["list_hash"])
    N = mp["nCells"]
    assert cases.rel_l2(force, g["force"]) <= TOL
    for k, w in (("uSource", 3), ("uSourceDrag", 1), ("alpha", 1), ("uParticle", 3)):
        assert cases.rel_l2(E.download(k), _scatter(g, N, k, w)) <= TOL, k
    E.close()


@pytest.mark.parametrize("n,P,seed,gaussian", [(32, 1000, 42, True), (32, 1000, 42, False), (24, 20000, 9, True),
                                               (64, 32000, 7, True), (64, 32000, 7, False)])
def test_engine_matches_live_reference(pkg, n, P, seed, gaussian):
    mo = meshgen.hex_box(n, n, n)
    mp = pkg.box_mesh(n, n, n, faces=False)
    flds = cases.fields_for(mo["C"])
    pd = cases.particles(P, seed, radius=0.1 / n, moving=True)
    r = cases.run_reference_step(mo, flds, pd, gaussian)
    e = cases.run_engine_step(pkg, mp, flds, pd, gaussian)
    cases.compare_steps(r, e, gaussian)


def test_locate_bit_exact_noncubic_box(pkg):
    """fy_locate == meshTree::nnearestCellsRange (truncated to its 12 nearest) on an anisotropic box,
    including points outside the mesh."""
    n = (20, 12, 28)
    mo = meshgen.hex_box(*n, lx=1.0, ly=0.5, lz=2.0, origin=(-0.3, 0.1, 0.0))
    mp = pkg.box_mesh(*n, lx=1.0, ly=0.5, lz=2.0, origin=(-0.3, 0.1, 0.0), faces=False)
    R = ref.RefFoamYade(mo, True)
    xyz = cases.particles(20000, 77, box=(1.2, 0.7, 2.2), origin=(-0.4, 0.0, -0.1), radius=0.01)[:, :3].copy()
    rc, ri = R.locate(xyz)
    R.close()
    E = pkg.Engine(mp)
    c, i = E.locate(xyz)
    fc = E.find_cell(xyz)
    E.close()
    rc12 = np.minimum(rc, 12)
    assert np.array_equal(c, rc12)
    m = np.arange(12)[None, :] < rc12[:, None]
    assert np.array_equal(i, np.where(m, ri[:, :12], -1))
    # findCell: index arithmetic, -1 outside
    h = np.array(mo["h"])
    o = np.array([-0.3, 0.1, 0.0])
    ijk = np.floor((xyz - o) / h)
    inside = np.all((ijk >= 0) & (ijk < np.array(n)), axis=1)
    cid = (ijk[:, 0] + n[0] * (ijk[:, 1] + n[1] * ijk[:, 2])).astype(np.int64)
    assert np.array_equal(fc, np.where(inside, cid, -1).astype(np.int32))


def test_full_size_c2_lists_and_forces(pkg):
    """BASELINE config C2 (128^3 cells, 1 M particles, seed 7), both branches, against the reference's own
    code at full size (its quadratic buildCellPartList replaced by the order-preserving dense accumulate,
    SURVEY.md H6; everything else is the reference's public methods)."""
    n, P = 128, 1000000
    mo = meshgen.hex_box(n, n, n)
    mp = pkg.box_mesh(n, n, n, faces=False)
    flds = cases.fields_for(mo["C"])
    pd = cases.particles(P, 7, radius=0.1 / n, moving=True)
    E = pkg.Engine(mp)
    for gaussian in (True, False):
        r = cases.run_reference_step(mo, flds, pd, gaussian)
        e = cases.run_engine_step(pkg, mp, flds, pd, gaussian, engine=E)
        cases.compare_steps(r, e, gaussian)
        if gaussian:
            # size-independent properties: weights sum to 1; void fraction conserved
            cnt, ids, w = e["cnt"], e["ids"], e["w"]
            s = w.sum(axis=1)
            assert np.all(np.abs(s[cnt > 0] - 1.0) < 1e-14)
            vp = np.pi * (2 * pd[:, 9]) ** 3 / 6.0
            np.testing.assert_allclose(((1.0 - e["alpha"]) * mo["V"]).sum(), vp[e["found"] == 1].sum(), rtol=1e-9)
            assert np.array_equal(np.bincount(cnt, minlength=13)[:2], np.bincount(r["cnt"], minlength=13)[:2])
    E.close()


def test_device_resident_path_and_repeatability(pkg):
    """fy_coupling_proc_device with torch-owned device buffers gives the same answer as the host path."""
    import torch
    n, P = 32, 5000
    mp = pkg.box_mesh(n, n, n, faces=False)
    flds = cases.fields_for(mp["C"])
    pd = cases.particles(P, 5, radius=0.1 / n, moving=True)
    E = pkg.Engine(mp)
    for gaussian in (True, False):
        E.set_properties(cases.RHOP, cases.RHOF, cases.NU, gaussian)
        for k in ("U", "gradP", "divT", "vGrad"):
            E.upload(k, flds[k])
        f0, F0 = E.set_particle_action(1e-3, pd)
        src0 = E.download("uSource")
        E.set_source_zero()
        d_pd = torch.from_numpy(pd).cuda()
        d_found = torch.empty(P, dtype=torch.int32, device="cuda")
        d_force = torch.empty(P, 6, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        E.coupling_begin(1e-3)
        E.coupling_proc_device(d_pd.data_ptr(), P, d_found.data_ptr(), d_force.data_ptr())
        E.synchronize()
        assert np.array_equal(d_found.cpu().numpy(), f0)
        assert cases.rel_l2(d_force.cpu().numpy(), F0) <= 1e-14
        assert cases.rel_l2(E.download("uSource"), src0) <= TOL
        E.set_source_zero()
    E.close()


def test_overlapped_wire_transfers_equal_the_blocking_call(pkg):
    """fy_particles_upload_async + fy_coupling_proc_staged + fy_results_wait (copies on the engine's second stream) give
    what fy_set_particle_action gives, step after step, with pinned host buffers that are refilled in between."""
    import torch
    n, P = 24, 4000
    mp = pkg.box_mesh(n, n, n, faces=False)
    flds = cases.fields_for(mp["C"])
    E = pkg.Engine(mp)
    L = E.L
    for gaussian in (True, False):
        E.set_properties(cases.RHOP, cases.RHOF, cases.NU, gaussian)
        for k in ("U", "gradP", "divT", "vGrad"):
            E.upload(k, flds[k])
        h_pd = torch.empty(P, 10, dtype=torch.float64).pin_memory()
        h_found = torch.zeros(P, dtype=torch.int32).pin_memory()
        h_force = torch.zeros(P, 6, dtype=torch.float64).pin_memory()
        for it in range(3):
            pd = cases.particles(P, 50 + it, radius=0.1 / n, moving=True)
            f0, F0 = E.set_particle_action(1e-3, pd)
            src0 = E.download("uSource")
            E.set_source_zero()
            h_pd.copy_(torch.from_numpy(pd))
            E._ck(L.fy_particles_upload_async(E.h, C.c_void_p(h_pd.data_ptr()), P))
            E.coupling_begin(1e-3)
            E._ck(L.fy_coupling_proc_staged(E.h, C.c_void_p(h_found.data_ptr()), C.c_void_p(h_force.data_ptr())))
            src1 = E.download("uSource")
            E.set_source_zero()
            E._ck(L.fy_results_wait(E.h))
            assert np.array_equal(h_found.numpy(), f0)
            assert cases.rel_l2(h_force.numpy(), F0) <= 1e-14
            assert cases.rel_l2(src1, src0) <= TOL
    E.close()


def _ddtU(C):
    return np.stack([0.4 * np.sin(2.0 * C[:, 2]), -0.3 * np.cos(3.0 * C[:, 1]), 0.2 + 0.1 * C[:, 0]], 1)


@pytest.mark.parametrize("n,P,seed,full,added_mass,torque,cap", [(32, 1000, 42, True, False, False, None), (32, 1000, 42, False, True, True, None),
                                                                 (24, 3000, 9, True, True, True, None), (48, 6000, 7, True, False, True, None),
                                                                 (24, 1500, 9, True, True, True, 300)])
def test_full_support_and_dormant_forces_match_reference(pkg, n, P, seed, full, added_mass, torque, cap, monkeypatch):
    """the full-support Gaussian mode (every cell within the k-d search bound, one warp per particle with
    warp-shuffle reductions) and the forces the reference defines but never calls (addedMassForce F.C:392-413, Gaussian
    torque F.C:467-478) against the UNMODIFIED reference: its own calcInterpWeightGaussian / hydroDragForce /
    archimedesForce / addedMassForce / calcHydroTorque, fed with the full cell sets by the harness.  Cell counts
    bit-exact (the in-range test is evaluated in meshTree::distance's operation order), forces, torques and all four
    per-cell fields within 1e-10; particles near and outside the walls included.  cap: the kernels keep a particle's hits
    in shared memory up to a capacity and recompute beyond it -- FY_RANGE_CAP = 300 sends the interior particles down the
    recomputing path and the wall particles down the compacted one."""
    if cap is not None:
        monkeypatch.setenv("FY_RANGE_CAP", str(cap))
    mo = meshgen.hex_box(n, n, n)
    mp = pkg.box_mesh(n, n, n, faces=False)
    flds = cases.fields_for(mo["C"])
    flds["ddtU"] = _ddtU(mo["C"])
    pd = cases.particles(P, seed, radius=0.1 / n, moving=True)
    pd[:40, 0:3] = pd[:40, 0:3] * 1.3 - 0.15            # some particles outside the box, some within reach of it
    dt = 1e-3
    R = ref.RefFoamYade(mo, True)
    R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
    R.set_gaussian_options(full, added_mass, torque)
    if full:
        # meshTree's descent never reports the tree's ROOT cell (MT.C:156,192), so the reference does not "find" a particle
        # whose nearest cell is the root although it lies inside the mesh; the full-support mode finds every particle with
        # a cell inside the bound.  Such particles (about P/N of them) are moved off the root cell for this comparison.
        cnt0, ids0 = R.locate(pd[:, 0:3])
        maxDist = 1.25 * R.constants()["interpRange"] ** 2
        for q in np.nonzero(cnt0 == 0)[0]:
            if ((mo["C"] - pd[q, 0:3]) ** 2).sum(1).min() < maxDist and np.all((pd[q, 0:3] > 0) & (pd[q, 0:3] < 1)):
                pd[q, 0:3] = pd[(q + 57) % P, 0:3] + 1e-3
    E = pkg.Engine(mp)
    E.set_properties(cases.RHOP, cases.RHOF, cases.NU, True)
    E.set_gaussian_options(full, added_mass, torque)
    for k in ("U", "gradP", "divT", "vGrad", "ddtU"):
        R.field(k)[:] = flds[k].reshape(R.field(k).shape)
        E.upload(k, flds[k])
    for step in range(2):                                # (the second step starts from setSourceZero's state)
        fr, Fr = R.step(dt, pd, yade_dt=0.5 * dt, pieces=True, truncate12=True, dense=True)
        cr, _ = R.lists(P)
        fe, Fe = E.set_particle_action(dt, pd)
        ce = E.last_counts(P)
        assert np.array_equal(fr, fe)
        assert np.array_equal(cr, ce)
        if full:
            inside = np.all((pd[:, 0:3] > 0.2) & (pd[:, 0:3] < 0.8), axis=1)
            assert 340 <= ce[inside].min() and ce[inside].max() <= 410     # 4/3 pi (sqrt(1.25) 4)^3 = 374.6 cells
            with pytest.raises(pkg.FyError):
                E.last_lists(P)
        assert cases.rel_l2(Fe[:, 0:3], Fr[:, 0:3]) <= TOL
        if torque:
            assert np.any(Fr[:, 3:6]) and cases.rel_l2(Fe[:, 3:6], Fr[:, 3:6]) <= TOL
        else:
            assert not np.any(Fe[:, 3:6]) and not np.any(Fr[:, 3:6])
        for k in ("uSource", "uSourceDrag", "alpha", "uParticle"):
            assert cases.rel_l2(E.download(k), R.field(k).reshape(E.download(k).shape)) <= TOL, (k, step)
        R.set_source_zero()
        E.set_source_zero()
        pd[:, 0:3] += dt * pd[:, 3:6]
    R.close()
    E.close()


@pytest.mark.parametrize("name", ["n16_gauss_full_support", "n16_gauss_full_support_dormant", "n16_gauss_trail_dormant"])
def test_engine_matches_f3_fixture(pkg, name):
    """the committed vectors of the full-support mode and of the dormant forces (tests/golden/gen_golden.py f3: the
    reference's own functions) against the kernels: counts and found flags exact, forces / torques / fields 1e-10"""
    from tests.golden.gen_golden import ddtU_of
    g = np.load(os.path.join(GOLD, name + ".npz"))
    n, P = int(g["n"]), int(g["P"])
    mp = pkg.box_mesh(n, n, n, faces=False)
    flds = cases.fields_for(mp["C"])
    flds["ddtU"] = ddtU_of(mp["C"])
    pd = cases.particles(P, int(g["seed"]), radius=0.1 / n, moving=True)
    pd[:, 0:3] = 0.05 + 0.9 * pd[:, 0:3]
    E = pkg.Engine(mp)
    E.set_properties(cases.RHOP, cases.RHOF, cases.NU, True)
    E.set_gaussian_options(bool(g["full"]), bool(g["added_mass"]), bool(g["torque"]))
    for k in ("U", "gradP", "divT", "vGrad", "ddtU"):
        E.upload(k, flds[k])
    found, force = E.set_particle_action(1e-3, pd)
    assert np.array_equal(found, g["found"]) and np.array_equal(E.last_counts(P), g["cnt"])
    assert cases.rel_l2(force, g["force"]) <= TOL
    for k in ("uSource", "uSourceDrag", "alpha", "uParticle"):
        assert cases.rel_l2(E.download(k), g[k].reshape(E.download(k).shape)) <= TOL, k
    E.close()


def test_empty_and_all_outside(pkg):
    mp = pkg.box_mesh(8, 8, 8, faces=False)
    E = pkg.Engine(mp)
    for gaussian in (True, False):
        E.set_properties(cases.RHOP, cases.RHOF, cases.NU, gaussian)
        f, F = E.set_particle_action(1e-3, np.zeros((0, 10)))
        assert f.shape == (0,) and F.shape == (0, 6)
        pd = np.zeros((4, 10))
        pd[:, :3] = [[5, 5, 5], [-3, 0.5, 0.5], [0.5, 9, 0.5], [0.5, 0.5, -7]]
        pd[:, 9] = 0.01
        f, F = E.set_particle_action(1e-3, pd)
        assert np.all(f == -1) and not np.any(F)
        assert not np.any(E.download("uSource"))
        assert np.all(E.download("alpha") == 1.0)
    E.close()
