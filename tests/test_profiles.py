"""The evidence pipeline under profiles/: bench.py reads the newest ncu traffic file that names the dominant kernel, and
that file is what tools/ncu_summary.py derives from the committed `ncu --set full` summaries."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_json_is_derived_from_the_committed_ncu_summaries():
    csvs = ["profiles/r2h_ncu_full_pcg_kernels.csv", "profiles/r2h_ncu_full_coupling_kernels.csv"]
    r = subprocess.run([sys.executable, "tools/ncu_summary.py", "traffic"] + csvs, cwd=ROOT, capture_output=True, text=True, check=True)
    got = json.loads(r.stdout)["kernels"]
    want = json.load(open(os.path.join(ROOT, "profiles", "r2h_ncu_traffic.json")))["kernels"]
    assert got == want
    for k in ("k_pen2<Op2DicFwd>", "k_pen2<Op2DicBwd>", "k_pen_tail<8>", "k_locate_gauss", "k_force_gauss"):
        assert k in got and got[k]["dram_read_MB"] > 0


def test_bench_reads_the_newest_capture_that_names_the_kernel():
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.ncu_traffic("k_pen2<Op2DicBwd> (DIC backward sweep + wA.rA)")
    assert src == os.path.join("profiles", "r2h_ncu_traffic.json") and 1.0e8 < t < 1.6e8       # 117 MB algorithmic, 128 MB measured
    t2, src2 = bench.ncu_traffic("k_pen_tail<8> (direction + Amul + update)")
    assert src2 == src and 1.5e8 < t2 < 3.0e8
    assert bench.ncu_traffic("k_no_such_kernel") == (None, None)
