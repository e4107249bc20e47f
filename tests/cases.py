"""Synthetic cases shared by the tests, __graft_entry__.smoke() and bench.py.

Recipe of SURVEY.md section 8(c): unit cube, n^3 hex cells (x fastest), particles uniform in the box
from std::mt19937_64(seed) drawn x,y,z per particle, U = (sin(6.28 y), 0.1 cos(6.28 x), 0.05) at cell
centres, gradP = (0.1,0,0), rhoP 2500, rhoF 1000, nu 1e-6.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RHOP, RHOF, NU = 2500.0, 1000.0, 1e-6


def mt_uniform(seed, n):
    """std::mt19937_64(seed) + uniform_real_distribution<double>(0,1): numpy's MT19937 is the 32-bit
    engine, so this is a small pure-numpy 64-bit Mersenne twister (libstdc++ generate_canonical<double,53>
    on a 64-bit engine takes one draw: (x) * 2^-64, clamped below 1)."""
    NN, MM = 312, 156
    MATRIX_A, UM, LM = 0xB5026F5AA96619E9, 0xFFFFFFFF80000000, 0x7FFFFFFF
    M64 = (1 << 64) - 1
    mt = [0] * NN
    mt[0] = seed & M64
    for i in range(1, NN):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & M64
    state = np.array(mt, dtype=np.uint64)
    out = np.empty(n, dtype=np.float64)
    pos = 0
    A = np.uint64(MATRIX_A)
    while pos < n:
        # regenerate the block (sequential dependency only across the three segments)
        s = state
        new = np.empty_like(s)
        # segment 1: i in [0, NN-MM)
        x = (s[:NN - MM] & np.uint64(UM)) | (s[1:NN - MM + 1] & np.uint64(LM))
        new[:NN - MM] = s[MM:NN] ^ (x >> np.uint64(1)) ^ np.where(x & np.uint64(1), A, np.uint64(0))
        # segment 2: i in [NN-MM, NN-1): depends on new[i + MM - NN]
        for lo in range(NN - MM, NN - 1, NN - MM):
            hi = min(lo + (NN - MM), NN - 1)
            x = (s[lo:hi] & np.uint64(UM)) | (s[lo + 1:hi + 1] & np.uint64(LM))
            new[lo:hi] = new[lo + MM - NN:hi + MM - NN] ^ (x >> np.uint64(1)) ^ np.where(x & np.uint64(1), A, np.uint64(0))
        x = (s[NN - 1] & np.uint64(UM)) | (new[0] & np.uint64(LM))
        new[NN - 1] = new[MM - 1] ^ (x >> np.uint64(1)) ^ (A if (x & np.uint64(1)) else np.uint64(0))
        state = new
        y = new.copy()
        y ^= (y >> np.uint64(29)) & np.uint64(0x5555555555555555)
        y ^= (y << np.uint64(17)) & np.uint64(0x71D67FFFEDA60000)
        y ^= (y << np.uint64(37)) & np.uint64(0xFFF7EEE000000000)
        y ^= (y >> np.uint64(43))
        take = min(NN, n - pos)
        v = y[:take].astype(np.float64) * (1.0 / 18446744073709551616.0)
        v[v >= 1.0] = np.nextafter(1.0, 0.0)
        out[pos:pos + take] = v
        pos += take
    return out


def fields_for(C):
    """U, gradP, divT, vGrad at cell centres (analytic, smooth, non-trivial in every component)."""
    N = C.shape[0]
    U = np.empty((N, 3))
    U[:, 0] = np.sin(6.28 * C[:, 1])
    U[:, 1] = 0.1 * np.cos(6.28 * C[:, 0])
    U[:, 2] = 0.05
    gradP = np.tile(np.array([0.1, 0.0, 0.0]), (N, 1))
    divT = np.empty((N, 3))
    divT[:, 0] = 0.3 * np.cos(3.0 * C[:, 2])
    divT[:, 1] = -0.2 * np.sin(2.0 * C[:, 0])
    divT[:, 2] = 0.1 * C[:, 1]
    vGrad = np.zeros((N, 9))
    vGrad[:, 1] = -0.628 * np.sin(6.28 * C[:, 0])      # xy = d/dx U_y
    vGrad[:, 3] = 6.28 * np.cos(6.28 * C[:, 1])        # yx = d/dy U_x
    vGrad[:, 2] = 0.01 * C[:, 2]
    vGrad[:, 5] = 0.02 * C[:, 0]
    vGrad[:, 6] = -0.03 * C[:, 1]
    vGrad[:, 7] = 0.04
    return dict(U=U, gradP=gradP, divT=divT, vGrad=vGrad)


def particles(n, seed, box=(1.0, 1.0, 1.0), radius=None, moving=False, origin=(0.0, 0.0, 0.0)):
    """[n][10] wire records: x y z vx vy vz wx wy wz radius."""
    u = mt_uniform(seed, 3 * n).reshape(n, 3)
    pd = np.zeros((n, 10))
    pd[:, 0] = origin[0] + u[:, 0] * box[0]
    pd[:, 1] = origin[1] + u[:, 1] * box[1]
    pd[:, 2] = origin[2] + u[:, 2] * box[2]
    if moving:
        v = mt_uniform(seed + 7919, 6 * n).reshape(n, 6)
        pd[:, 3:6] = 0.2 * (v[:, :3] - 0.5)
        pd[:, 6:9] = 2.0 * (v[:, 3:] - 0.5)
    pd[:, 9] = radius
    return pd


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))


def list_hash(cnt, ids):
    """order-sensitive 64-bit hash of the cell-id lists (own definition, used only to pin fixtures)."""
    cnt = np.asarray(cnt, dtype=np.int64)
    w = ids.shape[1]
    mask = np.arange(w)[None, :] < cnt[:, None]
    vals = np.where(mask, ids.astype(np.int64) + 1, 0).astype(np.uint64)
    mult = (np.arange(w, dtype=np.uint64) * np.uint64(2) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    row = (vals * mult[None, :]).sum(axis=1, dtype=np.uint64) + cnt.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
    idx = (np.arange(len(cnt), dtype=np.uint64) * np.uint64(0xD6E8FEB86659FD93)) | np.uint64(1)
    return int((row * idx).sum(dtype=np.uint64))


# ----------------------------------------------------------------------------------------------
# reference-vs-engine comparison of one coupling step (used by the gpu tests and smoke())
# ----------------------------------------------------------------------------------------------
def run_reference_step(mesh_o, flds, pdata, gaussian, n_yade=1, dt=1e-3, split=None, dense=True):
    from oracle import ref
    R = ref.RefFoamYade(mesh_o, gaussian, n_yade)
    R.set_properties(RHOP, RHOF, NU)
    for k in ("U", "gradP", "divT", "vGrad"):
        R.field(k)[:] = flds[k]
    found, force = R.step(dt, pdata, yade_dt=0.5 * dt, split=split, pieces=True, truncate12=True, dense=dense)
    out = dict(found=found.copy(), force=force.copy(), uSource=R.field("uSource").copy(),
               uSourceDrag=R.field("uSourceDrag").copy(), alpha=R.field("alpha").copy(),
               uParticle=R.field("uParticle").copy(), const=R.constants())
    if gaussian:
        cnt, ids = R.lists(pdata.shape[0])
        out["cnt"], out["ids"] = cnt, ids[:, :12]
    R.set_source_zero()
    out["zero"] = dict(uSource=R.field("uSource").copy(), alpha=R.field("alpha").copy(),
                       uSourceDrag=R.field("uSourceDrag").copy(), uParticle=R.field("uParticle").copy())
    R.close()
    return out


def run_engine_step(pkg, mesh_p, flds, pdata, gaussian, dt=1e-3, engine=None):
    E = engine or pkg.Engine(mesh_p)
    E.set_properties(RHOP, RHOF, NU, gaussian)
    for k in ("U", "gradP", "divT", "vGrad"):
        E.upload(k, flds[k])
    found, force = E.set_particle_action(dt, pdata)
    out = dict(found=found, force=force, uSource=E.download("uSource"), uSourceDrag=E.download("uSourceDrag"),
               alpha=E.download("alpha"), uParticle=E.download("uParticle"), const=E.constants())
    if gaussian:
        cnt, ids, w = E.last_lists(pdata.shape[0])
        out["cnt"], out["ids"], out["w"] = cnt, ids, w
    E.set_source_zero()
    out["zero"] = dict(uSource=E.download("uSource"), alpha=E.download("alpha"),
                       uSourceDrag=E.download("uSourceDrag"), uParticle=E.download("uParticle"))
    if engine is None:
        E.close()
    return out


TOL = 1e-10     # north star: forces and fields within 1e-10 relative L2 in fp64


def compare_steps(ref_out, eng_out, gaussian, tol=TOL):
    """Raises AssertionError with a message naming the first mismatch."""
    for k in ("interpRange", "sigmaInterp", "interpRangeCu", "sigmaPi"):
        assert ref_out["const"][k] == eng_out["const"][k], "constant %s differs" % k
    assert np.array_equal(ref_out["found"], eng_out["found"]), "found flags differ"
    if gaussian:
        assert np.array_equal(ref_out["cnt"], eng_out["cnt"]), "cell-list lengths differ"
        m = np.arange(12)[None, :] < ref_out["cnt"][:, None]
        assert np.array_equal(np.where(m, ref_out["ids"], -1), eng_out["ids"]), "cell ids differ (must be bit-exact)"
    errs = {}
    for k in ("force", "uSource", "uSourceDrag", "alpha", "uParticle"):
        errs[k] = rel_l2(eng_out[k], ref_out[k])
        assert errs[k] <= tol, "%s: relative L2 %.3e > %.1e" % (k, errs[k], tol)
    for k in ("uSource", "alpha", "uSourceDrag", "uParticle"):
        assert np.array_equal(ref_out["zero"][k], eng_out["zero"][k]), "setSourceZero: %s differs" % k
    return errs


def smoke_check(verbose=False):
    import __graft_entry__ as g
    from oracle import meshgen
    pkg = g.load_package()
    if pkg.lib().fy_device_count() < 1:
        raise RuntimeError("smoke(): no CUDA device; the engine has no CPU path")
    n, P = 32, 1000
    mo = meshgen.hex_box(n, n, n)
    mp = pkg.box_mesh(n, n, n)
    flds = fields_for(mo["C"])
    pd = particles(P, 42, radius=0.1 / n, moving=True)
    for gaussian in (True, False):
        r = run_reference_step(mo, flds, pd, gaussian)
        e = run_engine_step(pkg, mp, flds, pd, gaussian)
        errs = compare_steps(r, e, gaussian)
        if verbose:
            print("smoke coupling gaussian=%d: found %d/%d, rel-L2 %s" % (gaussian, int((e["found"] == 1).sum()), P,
                  {k: "%.1e" % v for k, v in errs.items()}))
    try:
        from tests import cases_fv
    except ImportError:
        cases_fv = None
    if cases_fv is not None:
        cases_fv.smoke_check_fv(pkg, verbose=verbose)
