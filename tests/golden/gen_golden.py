"""Generates tests/golden/*.npz by running the UNMODIFIED reference coupling operator
(/root/reference/FoamYade/FoamYade.C + meshtree/meshTree.C, built as oracle/_ref/libfoamyade_ref.so by
`make -C oracle ref`) on the seeded synthetic cases of tests/cases.py.  Run from the repo root in the
container that has /root/reference:   python tests/golden/gen_golden.py
The fixtures pin (a) the oracle harness itself (CPU test: _ref still reproduces them) and (b) the CUDA
engine (gpu test: fy_* output == fixture), so the GPU box needs neither /root/reference nor a rebuild.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import meshgen, ref  # noqa: E402
from tests import cases  # noqa: E402


def one(name, n, P, seed, gaussian, moving, n_yade=1, unmodified=False):
    mo = meshgen.hex_box(n, n, n)
    flds = cases.fields_for(mo["C"])
    pd = cases.particles(P, seed, radius=0.1 / n, moving=moving)
    if unmodified:
        # FoamYade::setParticleAction untouched (quadratic list scan, no <=12 truncation)
        R = ref.RefFoamYade(mo, gaussian, n_yade)
        R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
        for k in ("U", "gradP", "divT", "vGrad"):
            R.field(k)[:] = flds[k]
        found, force = R.step(1e-3, pd, yade_dt=5e-4)
        out = dict(found=found.copy(), force=force.copy(), uSource=R.field("uSource").copy(),
                   uSourceDrag=R.field("uSourceDrag").copy(), alpha=R.field("alpha").copy(),
                   uParticle=R.field("uParticle").copy())
        cnt, ids = R.locate(pd[:, :3])
        out["cnt"], out["ids"] = np.minimum(cnt, 12), ids[:, :12]
        R.close()
    else:
        out = cases.run_reference_step(mo, flds, pd, gaussian, n_yade=n_yade)
    touched = np.flatnonzero((np.abs(out["uSource"]).sum(axis=1) > 0) | (out["alpha"] != 1.0))
    save = dict(n=n, P=P, seed=seed, gaussian=int(gaussian), moving=int(moving), n_yade=n_yade,
                found=out["found"].astype(np.int8), force=out["force"], touched=touched.astype(np.int32),
                uSource=out["uSource"][touched], uSourceDrag=out["uSourceDrag"][touched],
                alpha=out["alpha"][touched], uParticle=out["uParticle"][touched])
    if gaussian:
        save["cnt"] = out["cnt"].astype(np.int8)
        save["ids"] = out["ids"].astype(np.int32)
        save["list_hash"] = np.uint64(cases.list_hash(out["cnt"], out["ids"]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(name, "found", int((out["found"] == 1).sum()), "/", P, "touched cells", touched.size)


def ddtU_of(C):
    return np.stack([0.4 * np.sin(2.0 * C[:, 2]), -0.3 * np.cos(3.0 * C[:, 1]), 0.2 + 0.1 * C[:, 0]], 1)


def one_f3(name, n, P, seed, full, added_mass, torque):
    """full-support cell sets fed to the reference's own weight / force functions, addedMassForce
    (FoamYade.C:392-413) and the Gaussian torque (FoamYade.C:467-478) -- oracle/ref_harness.cpp ref_set_gaussian_options"""
    mo = meshgen.hex_box(n, n, n)
    flds = cases.fields_for(mo["C"])
    flds["ddtU"] = ddtU_of(mo["C"])
    pd = cases.particles(P, seed, radius=0.1 / n, moving=True)
    pd[:, 0:3] = 0.05 + 0.9 * pd[:, 0:3]
    R = ref.RefFoamYade(mo, True)
    R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
    R.set_gaussian_options(full, added_mass, torque)
    for k in ("U", "gradP", "divT", "vGrad", "ddtU"):
        R.field(k)[:] = flds[k].reshape(R.field(k).shape)
    found, force = R.step(1e-3, pd, yade_dt=5e-4, pieces=True, truncate12=True, dense=True)
    cnt, _ = R.lists(P)
    save = dict(n=n, P=P, seed=seed, full=int(full), added_mass=int(added_mass), torque=int(torque), found=found.astype(np.int8),
                force=force.copy(), cnt=cnt.astype(np.int16), uSource=R.field("uSource").copy(), uSourceDrag=R.field("uSourceDrag").copy(),
                alpha=R.field("alpha").copy(), uParticle=R.field("uParticle").copy())
    R.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(name, "found", int((found == 1).sum()), "/", P, "cells per particle", cnt.min(), "...", cnt.max())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "f3":           # only the fixtures added in round 2 (the others stay byte-identical)
        one_f3("n16_gauss_full_support", 16, 300, 21, True, False, False)
        one_f3("n16_gauss_full_support_dormant", 16, 300, 21, True, True, True)
        one_f3("n16_gauss_trail_dormant", 16, 300, 21, False, True, True)
        sys.exit(0)
    one("c1_gauss_static", 32, 1000, 42, True, False, unmodified=True)
    one("c1_gauss_moving", 32, 1000, 42, True, True)
    one("c1_point_moving", 32, 1000, 42, False, True)
    one("c1_gauss_parallel3", 32, 1000, 43, True, True, n_yade=3)
    one("c1_point_parallel3", 32, 1000, 43, False, True, n_yade=3)
    one("n16_gauss_dense", 16, 4000, 5, True, True)
