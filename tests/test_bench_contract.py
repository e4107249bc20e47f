"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm prints one JSON line with the keys
the driver reads, and the engine arm refuses to run without a CUDA device (there is no CPU fallback to time)."""
import json
import os
import subprocess
import sys

import pytest

from oracle import port, ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not (port.available() and ref.available()), reason="oracle libraries not built")


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          timeout=600)


@pytest.mark.parametrize("extra", [(), ("--solver", "pimple"), ("--coupling", "point")])
def test_reference_arm_line(extra):
    r = _run("--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0", *extra)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "coupled timesteps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6 * 1e3
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["config"]["workload"].startswith("C1") and line["dtype"] == "f64" and line["vs_baseline"] is None


def test_engine_arm_needs_a_gpu():
    r = _run("--workload", "C1", "--steps", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
