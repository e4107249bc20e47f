"""Dev check (GPU): PCG/DIC on an n^3 (or given) box vs the CPU oracle, increasing iteration caps."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
from oracle import port
from tests import cases, cases_fv
dims = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (40, 40, 40)
pkg = g.load_package()
mo, mp = cases_fv.cavity3d(pkg, dims)
rng = np.random.default_rng(4)
N, Fi = mo["nCells"], mo["nInternalFaces"]
upper = rng.uniform(0.5, 1.5, Fi)
diag = np.zeros(N)
np.subtract.at(diag, mo["owner"], upper)
np.subtract.at(diag, mo["neighbour"], upper)
diag -= rng.uniform(0.001, 0.01, N)
b = rng.standard_normal(N)
O = port.IcoOracle(mo)
E = pkg.Engine(mp)
for mi in (0, 1, 2, 5, 30):
    xo, po = O.pcg(diag, upper, b, np.zeros(N), tol=1e-14, relTol=0.0, maxIter=mi, preconditioner="DIC")
    t0 = time.time()
    try:
        xe, pe = E.pcg(diag, upper, b, np.zeros(N), tol=1e-14, relTol=0.0, maxIter=mi, preconditioner="DIC")
    except Exception as ex:
        print("maxIter", mi, "ENGINE ERROR", ex)
        break
    print("maxIter %d: iters %d/%d relL2 %.2e final %.3e/%.3e  (%.1f ms)" % (
        mi, pe["iters"], po["iters"], cases.rel_l2(xe, xo), pe["final"], po["final"], (time.time() - t0) * 1e3))
