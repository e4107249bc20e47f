"""Generates tests/golden/fluid_*.npz: two time steps of the fluid-half restatements (oracle/fv_oracle.cc) on a small
channel -- icoFoamYade's PISO step with a momentum source, and pimpleFoamYade's UcEqn/pEqn step with a void-fraction
blob, implicit drag, a source and gravity.  These are REGRESSION pins of the restatement (it has no external golden
vectors for the fluid half), so that an edit of the oracle cannot silently move the target the
CUDA path is compared with.   python tests/golden/gen_golden_fluid.py   (from the repo root, after `make -C oracle`)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import port  # noqa: E402
from tests import cases_fv  # noqa: E402

N3, DT, NU, G = (14, 8, 6), 0.02, 0.005, (0.0, -0.2, 0.05)


def drive(C, it):
    lo, hi = C.min(0), C.max(0)
    x = (C - lo) / (hi - lo)
    blob = np.exp(-(((x - np.array([0.4, 0.55, 0.5])) ** 2).sum(1)) / 0.04)
    alpha = 1.0 - (0.35 + 0.05 * it) * blob
    drag = -(40.0 + 10.0 * it) * (1.0 - alpha)
    src = np.stack([0.3 * (1.0 - alpha) * np.sin(5.0 * x[:, 1]), -0.8 * (1.0 - alpha), 0.1 * blob * x[:, 0]], 1)
    return alpha, drag, src


def run(solver):
    mo, _ = cases_fv.channel(None, N3)
    U, p = cases_fv.channel_init(mo["C"])
    O = port.IcoOracle(mo, nu=NU)
    O.field("U")[:] = U
    O.field("p")[:] = p
    O.create_phi()
    iters = []
    for it in range(2):
        alpha, drag, src = drive(mo["C"], it)
        O.field("uSource")[:] = src
        if solver == "pimple":
            O.pimple_solve(DT, alpha, drag, G)
        else:
            O.pre(DT)
            O.solve(DT)
        st = O.stats()
        iters.append([q["iters"] for q in st["p"][:st["nPSolves"]]] + [q["iters"] for q in st["U"]])
    out = dict(U=O.field("U").copy(), p=O.field("p").copy(), phi=O.field("phi").copy(), iters=np.array(iters, dtype=np.int32),
               contErr=np.array([st["sumLocalContErr"], st["globalContErr"]]))
    O.close()
    return out


if __name__ == "__main__":
    for solver in ("ico", "pimple"):
        out = run(solver)
        np.savez_compressed(os.path.join(HERE, "fluid_%s_channel.npz" % solver), **out)
        print(solver, "iters", out["iters"].tolist(), "|U|max", float(np.abs(out["U"]).max()))
