"""yade-openfoam-coupling_b200 -- B200-native FoamYade coupling engine.

The product is `libfycuda.so` (hand-written sm_100a CUDA behind the C ABI of
include/fycuda.h) plus the C++ host mirror of the reference's operator surface
in host/.  This Python package is the thin ctypes binding the tests, bench.py
and __graft_entry__ use; it has NO CPU fallback: importing works anywhere, but
creating an Engine without the built library or without a CUDA device raises.

The directory name contains '-', so load it with
    importlib.util.spec_from_file_location("yade_openfoam_coupling_b200", ".../__init__.py")
(see __graft_entry__.load_package()).
"""
from .binding import Engine, FyError, lib, lib_path, FIELD  # noqa: F401
from .mesh import box_mesh, set_bc, BC_FIXED_VALUE, BC_ZERO_GRADIENT, BC_EMPTY, BC_FIXED_FLUX_PRESSURE  # noqa: F401
from . import replicas  # noqa: F401,E402
from . import sharded  # noqa: F401,E402
from . import domain  # noqa: F401,E402
from . import foamcase  # noqa: F401,E402
