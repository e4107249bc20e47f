"""GPU: an OpenFOAM case directory run on the engine as it lies on disk (tools/foam_case_run.py over
yade-openfoam-coupling_b200/foamcase.py): the stock cavity tutorial's files -> the solver log OpenFOAM-6 prints for it
(the public log.icoFoam digits pinned in tests/test_fv_oracle.py) and the time directory runTime.write() leaves."""
import contextlib
import io
import os
import re
import sys

import numpy as np
import pytest

from tests.test_fv_oracle import CAVITY_LOG
from tests.test_foamcase import cavity_case, fc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cavity_case_directory_runs_on_the_gpu_and_prints_the_tutorial_log(pkg, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import foam_case_run
    case_dir = cavity_case(tmp_path)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        foam_case_run.main([case_dir, "--steps", "3"])
    log = buf.getvalue()
    steps = log.split("Time = ")[1:]
    assert len(steps) == 3 and steps[0].startswith("0.005") and steps[2].startswith("0.015")
    num = r"([-+0-9.eE]+)"
    for txt, ref in zip(steps, CAVITY_LOG):
        co = re.search(r"Courant Number mean: %s max: %s" % (num, num), txt)
        assert (float(co.group(1)), float(co.group(2))) == ref["Co"]
        for k in ("Ux", "Uy"):
            m = re.search(r"Solving for %s, Initial residual = %s, Final residual = %s, No Iterations (\d+)" % (k, num, num), txt)
            if ref[k][2] == 0 and m is None:
                continue
            assert (float(m.group(1)), float(m.group(2)), int(m.group(3))) == ref[k], k
        ps = re.findall(r"Solving for p, Initial residual = %s, Final residual = %s, No Iterations (\d+)" % (num, num), txt)
        assert [(float(a), float(b), int(c)) for a, b, c in ps] == [ref["p1"], ref["p2"]]
        ce = re.findall(r"sum local = %s" % num, txt)
        assert float(ce[1]) == ref["c2"]
    assert "Solving for Uz" not in log and log.rstrip().endswith("End")
    # the shortened run leaves its last state as a time directory in OpenFOAM's format
    back = fc.load_case(case_dir, time="0.015")
    assert np.abs(back["U"]).max() > 0.1 and back["patches"][0]["valueU"] == (1.0, 0.0, 0.0)
    txt = open(os.path.join(case_dir, "0.015", "U")).read()
    assert "class       volVectorField;" in txt and "internalField   nonuniform List<vector>" in txt and "type            empty;" in txt


def _check_log(log, case_dir):
    steps = log.split("Time = ")[1:]
    assert len(steps) == 3 and steps[0].startswith("0.005") and steps[2].startswith("0.015")
    num = r"([-+0-9.eE]+)"
    for txt, ref in zip(steps, CAVITY_LOG):
        co = re.search(r"Courant Number mean: %s max: %s" % (num, num), txt)
        assert (float(co.group(1)), float(co.group(2))) == ref["Co"]
        for k in ("Ux", "Uy"):
            m = re.search(r"Solving for %s, Initial residual = %s, Final residual = %s, No Iterations (\d+)" % (k, num, num), txt)
            assert (float(m.group(1)), float(m.group(2)), int(m.group(3))) == ref[k], k
        ps = re.findall(r"DICPCG:  Solving for p, Initial residual = %s, Final residual = %s, No Iterations (\d+)" % (num, num), txt)
        assert [(float(a), float(b), int(c)) for a, b, c in ps] == [ref["p1"], ref["p2"]]
        assert float(re.findall(r"sum local = %s" % num, txt)[1]) == ref["c2"]
    assert "Solving for Uz" not in log and log.rstrip().endswith("End")
    back = fc.load_case(case_dir, time="0.015")
    assert np.abs(back["U"]).max() > 0.1 and back["patches"][0]["valueU"] == (1.0, 0.0, 0.0)


def test_standalone_cpp_solver_runs_the_cavity_case(pkg, tmp_path):
    """yade-openfoam-coupling_b200/foamYadeB200 (host/foamYadeRun.cpp + host/foamCase.H over the C ABI, no Python in the
    loop): the stock cavity tutorial's directory in, OpenFOAM's solver log and time directory out."""
    import subprocess
    binary = os.path.join(ROOT, "yade-openfoam-coupling_b200", "foamYadeB200")
    assert os.path.exists(binary), "foamYadeB200 not built (make -C yade-openfoam-coupling_b200)"
    case_dir = cavity_case(tmp_path)
    r = subprocess.run([binary, "-case", case_dir, "-steps", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    _check_log(r.stdout, case_dir)
    # with particles standing in for the Yade side (point-force branch, as icoFoamYade.C:53 hard-codes it)
    from tests import cases
    pd = cases.particles(200, 11, radius=1e-4, moving=True)
    pd[:, 0:3] *= [0.1, 0.1, 0.01]
    pd[:, 3:9] = 0.0
    pd.tofile(os.path.join(case_dir, "particles.bin"))
    r2 = subprocess.run([binary, "-case", case_dir, "-steps", "2", "-particles", os.path.join(case_dir, "particles.bin"), "-noWrite"],
                        capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0 and r2.stdout.count("Time = ") == 2, r2.stderr
