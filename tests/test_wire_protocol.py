"""CPU: the Yade<->Foam message sequence of the reference (SURVEY.md section 4 traces), recorded from the
unmodified FoamYade.C through the stub MPI of oracle/shim/mpi.h."""
import numpy as np
import pytest

from oracle import meshgen, ref
from tests import cases

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def _trace(gaussian, n_yade, P=5):
    mo = meshgen.hex_box(8, 8, 8)
    R = ref.RefFoamYade(mo, gaussian, n_yade)
    R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
    pd = cases.particles(P, 1, radius=0.01)
    pd[:, :3] = 0.2 + 0.5 * pd[:, :3]
    R.L.ref_clear_trace()
    R.L.ref_set_logging(1)
    found, force = R.step(1e-3, pd, yade_dt=2.5e-4)
    tr = R.trace()
    R.L.ref_set_logging(0)
    dts = R.dts()
    R.close()
    return tr, found, dts


def test_serial_gaussian_sequence():
    tr, found, dts = _trace(True, 1)
    exp = (["Bcast i32[1] peer=0 tag=-1 WORLD", "Bcast f64[50] peer=0 tag=-1 WORLD"]
           + ["Allreduce i32[1] peer=-1 tag=-1 WORLD"] * 5 + ["Allreduce f64[1] peer=-1 tag=-1 WORLD"] * 30
           + ["Send f64[1] peer=0 tag=1050 WORLD", "Bcast f64[1] peer=0 tag=-1 WORLD"])
    assert tr == exp
    assert dts == (1e-3, 2.5e-4)


def test_serial_point_force_sequence():
    tr, found, dts = _trace(False, 1)
    exp = (["Bcast i32[1] peer=0 tag=-1 WORLD", "Bcast f64[50] peer=0 tag=-1 WORLD"]
           + ["Allreduce i32[1] peer=-1 tag=-1 WORLD"] * 5 + ["Send f64[6] peer=0 tag=1005 WORLD"] * 5
           + ["Send f64[1] peer=0 tag=1050 WORLD", "Bcast f64[1] peer=0 tag=-1 WORLD"])
    assert tr == exp


@pytest.mark.parametrize("gaussian", [True, False])
def test_parallel_sequence(gaussian):
    mo = meshgen.hex_box(8, 8, 8)
    ref.lib().ref_clear_trace()
    ref.lib().ref_set_logging(1)
    R = ref.RefFoamYade(mo, gaussian, 3)            # world = 4: Yade master 0 + workers 1,2; Foam rank = world 3
    ctor = R.trace()
    assert ctor == ["Isend f64[6] peer=0 tag=1001 WORLD", "Isend f64[6] peer=1 tag=1001 WORLD",
                    "Isend f64[6] peer=2 tag=1001 WORLD", "Wait -[0] peer=-1 tag=-1 -"] + ["Wait -[0] peer=-1 tag=-1 -"] * 2
    bb = np.empty(18)
    assert R.L.ref_get_bbox(bb.ctypes.data_as(ref._dp), 18) == 18
    assert bb[:6].tolist() == [0, 0, 0, 1, 1, 1]
    R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
    pd = cases.particles(10, 1, radius=0.01)
    pd[:, :3] = 0.2 + 0.5 * pd[:, :3]
    R.L.ref_clear_trace()
    found, force = R.step(1e-3, pd, yade_dt=2.5e-4)
    tr = R.trace()
    R.L.ref_set_logging(0)
    exp = ["Recv i32[1] peer=1 tag=1003 WORLD", "Recv i32[1] peer=2 tag=1003 WORLD",
           "Recv f64[50] peer=1 tag=1002 WORLD", "Recv f64[50] peer=2 tag=1002 WORLD",
           "Send i32[5] peer=1 tag=1004 WORLD", "Send i32[5] peer=2 tag=1004 WORLD",
           "Send f64[30] peer=1 tag=1005 WORLD", "Send f64[30] peer=2 tag=1005 WORLD",
           "Send f64[1] peer=0 tag=1050 WORLD", "Recv f64[1] peer=0 tag=1060 WORLD", "Bcast f64[1] peer=0 tag=-1 FOAM"]
    assert tr == exp
    assert np.all(found == 1)
    assert R.dts() == (1e-3, 2.5e-4)
    R.close()
