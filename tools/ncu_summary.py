#!/usr/bin/env python
"""Condense an ncu report into the per-kernel summaries kept under profiles/.

  python tools/ncu_summary.py full   gpurun_out/x.ncu-rep  > profiles/rNN_ncu_full_*.csv     (from `ncu --set full`)
  python tools/ncu_summary.py launch gpurun_out/launches.csv > profiles/rNN_launch_summary_*.csv
      (from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`)
Per-launch values are averaged per kernel name (template arguments kept, namespaces and parameter lists dropped)."""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size"]
SCALE = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3,
         "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}


def short(name):
    name = re.sub(r"(\(anonymous namespace\)|<unnamed>)::", "", name)
    name = re.sub(r"^void ", "", name)
    depth, out = 0, []
    for ch in name:                      # drop the parameter list, keep template arguments
        if ch == "(" and depth == 0:
            break
        depth += ch == "<"
        depth -= ch == ">"
        out.append(ch)
    return "".join(out).replace(",", ";").strip()


def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    acc, cnt = OrderedDict(), defaultdict(int)
    for r in data:
        k = short(r[col["Kernel Name"]])
        vals = []
        for m in FULL:
            v = float(r[col[m]].replace(",", "")) if r[col[m]] else 0.0
            vals.append(v * SCALE.get(units[col[m]], 1.0))
        a = acc.setdefault(k, [0.0] * len(FULL))
        for i, v in enumerate(vals):
            a[i] += v
        cnt[k] += 1
    print("kernel,launches," + ",".join(m + (" [ms]" if "time" in m else " [Gbyte]" if "bytes" in m else "") for m in FULL))
    for k, a in acc.items():
        print("%s,%d," % (k, cnt[k]) + ",".join("%.6g" % (v / cnt[k]) for v in a))


def launch(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot, cnt = OrderedDict(), defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", "")) * SCALE.get(r["Metric Unit"], 1.0)
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] += 1
    s = sum(tot.values())
    print("kernel,launches,total_ms,share")
    for k in sorted(tot, key=lambda x: -tot[x]):
        print("%s,%d,%.3f,%.4f" % (k, cnt[k], tot[k], tot[k] / s))


def traffic(*csvs):
    """profiles/rNN_ncu_traffic.json (what bench.py reads for roofline.traffic) from `full` summaries:
       python tools/ncu_summary.py traffic profiles/rNN_ncu_full_a.csv [profiles/rNN_ncu_full_b.csv ...] > profiles/rNN_ncu_traffic.json"""
    import json
    ks = OrderedDict()
    for path in csvs:
        for r in csv.DictReader(open(path)):
            name = re.sub(r"<(Op2?\w+);.*>", r"<\1>", r["kernel"])       # k_pen2<Op2DicBwd; 1; 1; 4> -> k_pen2<Op2DicBwd>
            ks[name] = {"dram_read_MB": float(r["dram__bytes_read.sum [Gbyte]"]) * 1e3,
                        "dram_write_MB": float(r["dram__bytes_write.sum [Gbyte]"]) * 1e3,
                        "duration_us_under_ncu": float(r["gpu__time_duration.sum [ms]"]) * 1e3}
    print(json.dumps({"source": "ncu --set full --clock-control none on `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (C2, 128^3 cells "
                                "/ 1M particles), per launch, mid-solve; summaries: " + ", ".join(csvs),
                      "workload": "C2", "kernels": ks}, indent=1))


if __name__ == "__main__":
    {"full": full, "launch": launch, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
