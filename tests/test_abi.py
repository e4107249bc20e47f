"""CPU: libfycuda.so loads, exports every symbol include/fycuda.h declares, and refuses to compute
without a CUDA device (there is no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(fy_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libfycuda.so does not export %s" % n


def test_version(pkg):
    assert pkg.lib().fy_version().decode().startswith("fycuda")


def test_no_cpu_fallback(pkg):
    L = pkg.lib()
    if L.fy_device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(pkg.FyError):
        pkg.Engine(pkg.box_mesh(4, 4, 4))
    assert b"no CPU path" in L.fy_last_error(None)


def test_invalid_arguments(pkg):
    L = pkg.lib()
    h = ctypes.c_void_p()
    assert L.fy_create(None, 0, ctypes.byref(h)) == -1
    assert L.fy_set_properties(None, 1.0, 1.0, 1.0, 0) == -1
    assert L.fy_set_source_zero(None) == -1
    assert L.fy_destroy(None) == 0


def test_box_mesh_matches_oracle_generator(pkg):
    from oracle import meshgen
    for n in ((4, 5, 6), (8, 8, 8)):
        a = pkg.box_mesh(*n, lx=1.0, ly=2.0, lz=0.5)
        b = meshgen.hex_box(*n, lx=1.0, ly=2.0, lz=0.5)
        assert np.array_equal(a["C"], b["C"]) and np.array_equal(a["V"], b["V"])
        Fi = a["nInternalFaces"]
        assert Fi == 3 * n[0] * n[1] * n[2] - n[0] * n[1] - n[1] * n[2] - n[0] * n[2]
        o, nb = a["owner"], a["neighbour"]
        assert np.all(o < nb)
        key = o.astype(np.int64) * a["nCells"] + nb
        assert np.all(np.diff(key) > 0)      # upper-triangular order
        assert sum(p["faceCells"].shape[0] for p in a["patches"]) == 2 * (n[0] * n[1] + n[1] * n[2] + n[0] * n[2])
