"""Dev probe (GPU): times fy_dic_precondition's pencil sweeps on an n^3 box and prints the per-pencil
time stamps of the last launch (FY_PENCIL_TRACE=1).  Not part of the product or the tests."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FY_PENCIL_TRACE", "1")
import __graft_entry__ as g  # noqa: E402
from tests import cases_fv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pkg = g.load_package()
mo, mp = cases_fv.cavity3d(pkg, (n, n, n))
rng = np.random.default_rng(2)
N, Fi = mo["nCells"], mo["nInternalFaces"]
upper = -rng.uniform(0.5, 1.5, Fi)
diag = np.zeros(N)
np.subtract.at(diag, mo["owner"], upper)
np.subtract.at(diag, mo["neighbour"], upper)
diag += rng.uniform(0.01, 0.05, N)
r = rng.standard_normal(N)
E = pkg.Engine(mp)
E.set_profiling(True)
E.kernel_ms(reset=True)
b = rng.standard_normal(N)
diag2 = -diag
up2 = -upper
x, perf = E.pcg(diag2, up2, b, np.zeros(N), tol=1e-12, relTol=0.0, maxIter=40, preconditioner="DIC")
km = E.kernel_ms(reset=True)
print("pcg iters %d; kernel ms per launch: %s" % (perf["iters"], {k: round(v, 4) for k, v in km.items()}))
w = E.dic(diag, upper, r)
G = E.L
nJB = (n + 31) // 32
W = int(os.environ.get("FY_PENCIL_W", "8"))
tr = np.empty((nJB * (n + 16 * 8 + 64) * 4 + 64) * 32)
E._ck(G.fy_fv_get(E.h, b"pencilTrace", tr.ctypes.data_as(C.POINTER(C.c_double))))
nKQ = (n + W - 1) // W
NW = 2 * W + 1
tr = tr[: nKQ * nJB * NW * 32].reshape(nKQ, nJB, NW, 32)
base = tr[..., 0][tr[..., 0] >= 0].min()
print("last launch (backward sweep): span %.1f us" % ((tr[..., 3].max() - base) / 1e3))
f = lambda a: " ".join("%5.1f" % ((x - base) / 1e3) for x in a)
for jb in range(nJB - 1, -1, -1):
    print("column jb=%d: [start of block 0, 1, 10, 19 | end] of warp 0 and warp W-1" % jb)
    for kq in (range(nKQ) if os.environ.get("FY_PROBE_FWD") else range(nKQ - 1, -1, -1)):
        c = tr[kq, jb]
        print("  kq %2d w0: %s | %s   w%d: %s | %s" % (kq, f(c[0, [8, 9, 18, 27]]), f(c[0, [3]]), W - 1, f(c[W - 1, [8, 9, 18, 27]]), f(c[W - 1, [3]])))
E.close()
