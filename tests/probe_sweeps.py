"""Dev probe (GPU, not a test): times the DIC sweeps of the PCG iteration on an n^3 box for a list of pencil
pipeline configurations (one subprocess per configuration, the knobs are environment variables read at engine
creation) and checks fy_dic_precondition bit for bit against the oracle in each.

    python -m tests.probe_sweeps 128 "FY_PEN2_Z=4 FY_PEN2_R=4" "FY_PENCIL_VER=1" ...
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(n):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    from tests import cases_fv
    pkg = g.load_package()
    check = os.environ.get("PROBE_CHECK", "1") == "1"
    dims = tuple(int(x) for x in os.environ["PROBE_DIMS"].split(",")) if os.environ.get("PROBE_DIMS") else (n, n, n)
    mo, mp = cases_fv.cavity3d(pkg, dims, oracle=check)
    rng = np.random.default_rng(2)
    N, Fi = mp["nCells"], mp["nInternalFaces"]
    upper = -rng.uniform(0.5, 1.5, Fi)
    diag = np.zeros(N)
    np.subtract.at(diag, mp["owner"], upper)
    np.subtract.at(diag, mp["neighbour"], upper)
    diag += rng.uniform(0.01, 0.05, N)
    r = rng.standard_normal(N)
    E = pkg.Engine(mp)
    out = dict(n=n)
    t0 = time.time()
    w = E.dic(diag, upper, r)
    if check:
        from oracle import port
        O = port.IcoOracle(mo)
        out["dic_bit_exact"] = bool(np.array_equal(w, O.dic(diag, upper, r)))
        O.close()
    E.set_profiling(True)
    b = rng.standard_normal(N)
    for rep in range(2):
        E.kernel_ms(reset=True)
        x, perf = E.pcg(-diag, -upper, b, np.zeros(N), tol=1e-30, relTol=0.0, maxIter=int(os.environ.get("PROBE_ITERS", "48")), preconditioner="DIC")
        km = E.kernel_ms(reset=True)
    out["iters"] = perf["iters"]
    out["final"] = perf["final"]
    out.update({k: round(v / max(1, km["samples"]), 5) if k not in ("samples", "pcg_iterations") else v for k, v in km.items()})
    # whole-solve wall time without profiling (graph replay)
    E.set_profiling(False)
    E.pcg(-diag, -upper, b, np.zeros(N), tol=1e-30, relTol=0.0, maxIter=96, preconditioner="DIC")
    E.synchronize()
    t0 = time.time()
    E.pcg(-diag, -upper, b, np.zeros(N), tol=1e-30, relTol=0.0, maxIter=96, preconditioner="DIC")
    E.synchronize()
    out["ms_per_iter_wall"] = round((time.time() - t0) * 1e3 / 97, 5)
    if os.environ.get("FY_PENCIL_TRACE") and os.environ.get("FY_PENCIL_VER", "2") != "1":
        import ctypes as C
        Z = int(os.environ.get("FY_PEN2_Z", "4")); W = int(os.environ.get("FY_PEN2_W", "1"))
        E.dic(diag, upper, r)
        nJB = (n + 31) // 32
        tr = np.empty((nJB * (n + 16 * 8 + 64) * 4 + 64) * 32)
        E._ck(E.L.fy_fv_get(E.h, b"pencilTrace", tr.ctypes.data_as(C.POINTER(C.c_double))))
        nKQ = (n + W * Z - 1) // (W * Z)
        raw = tr.copy()
        tr = tr[: nKQ * nJB * 32 * 8].reshape(nKQ, nJB, 32, 8)
        lines = []
        rev = not (int(os.environ.get("FY_PENCIL_DBG", "0")) & 64)
        for jb in range(nJB):
            jj = nJB - 1 - jb if rev else jb
            row = []
            for kq in range(nKQ):
                kk = kq
                c = tr[kk, jj, 0]
                row.append("%.1f-%.1f" % (c[0] / 1e3, c[1] / 1e3))
            lines.append("jb%d: " % jj + " ".join(row))
        out["trace_us(start-end of compute warp 0 per kq, sweep order)"] = lines
        if os.environ.get("PROBE_TSEC"):
            # PEN2_TIMING build: cycles per section of a step (A loads+flow control, C y/shuffle/pre, D z wait + finish + hand-over, E stores),
            # per block of R steps, for the head warp and a mid-chain consumer of the first column
            jj = nJB - 1 if rev else 0
            sec = {}
            raw = np.empty_like(raw)
            E._ck(E.L.fy_fv_get(E.h, b"pencilTraceRaw", raw.ctypes.data_as(C.POINTER(C.c_double))))
            for name, kq, wv in (("head", 0, 0), ("consumer_kq0_w3", 0, 3), ("consumer_kq5_w0", min(5, nKQ - 1), 0)):
                c = raw[: nKQ * nJB * 32 * 8].reshape(nKQ, nJB, 32, 8)[kq, jj, wv]
                # the stored values are offsets from the earliest stamp (fy_fv_get), counters start at 0: undo by differences
                nb = max(c[6], 1.0)
                sec[name] = dict(blocks=int(c[6]), cycles_per_step=[round(float(x) / nb / 4, 1) for x in c[2:6]])
            out["tsec_raw"] = sec
    E.close()
    print("PROBE " + json.dumps(out), flush=True)


def main():
    n = int(sys.argv[1])
    configs = sys.argv[2:] or [""]
    for cfg in configs:
        env = dict(os.environ)
        for kv in cfg.split():
            k, v = kv.split("=")
            env[k] = v
        r = subprocess.run([sys.executable, "-m", "tests.probe_sweeps", "--child", str(n)], cwd=ROOT, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
        print("[%s] n=%d rc=%d %s" % (cfg, n, r.returncode, line[0][6:] if line else r.stdout[-600:]), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        main()
