"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper of oracle/_ref/libfoamyade_ref.so: the UNMODIFIED reference
coupling operator (/root/reference/FoamYade/FoamYade.C + meshtree/meshTree.C)
behind the shim + harness of oracle/ref_harness.cpp.  Build with `make -C oracle ref`
(done by __graft_entry__.build() where /root/reference exists; the built .so
travels to the GPU box).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfoamyade_ref.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _ip, _dp, C.c_int, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_properties.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.ref_field.restype = _dp
        L.ref_field.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_get_constants.argtypes = [C.c_void_p, _dp]
        L.ref_set_logging.argtypes = [C.c_int]
        L.ref_get_trace.restype = C.c_int
        L.ref_get_trace.argtypes = [C.c_char_p, C.c_int]
        L.ref_get_counts.argtypes = [C.POINTER(C.c_long)]
        L.ref_get_dt.argtypes = [_dp, C.c_void_p]
        L.ref_get_bbox.restype = C.c_int
        L.ref_get_bbox.argtypes = [_dp, C.c_int]
        L.ref_locate.argtypes = [C.c_void_p, _dp, C.c_int, _ip, _ip, C.c_int]
        L.ref_find_cell.restype = C.c_int
        L.ref_find_cell.argtypes = [C.c_void_p, _dp]
        L.ref_step.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, C.c_int, _ip, _ip, _dp]
        L.ref_step_pieces.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, C.c_int, _ip, C.c_int, C.c_int, _ip, _dp]
        L.ref_set_gaussian_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_get_lists.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
        L.ref_get_times.argtypes = [C.c_void_p, _dp]
        L.ref_set_source_zero.argtypes = [C.c_void_p]
        L.ref_mt19937_64_uniform.argtypes = [C.c_ulonglong, C.c_long, _dp]
        _lib = L
    return _lib


HOST_LIB_PATH = os.path.join(_HERE, "_build", "libhost_harness.so")
_host = None


def host_available():
    return os.path.exists(HOST_LIB_PATH)


def host_lib():
    """oracle/_build/libhost_harness.so: the same fake Yade peer and fields around the PRODUCT's host class
    (yade-openfoam-coupling_b200/host/FoamYadeB200.H -> libfycuda.so).  GPU tests only."""
    global _host
    if _host is None:
        L = C.CDLL(HOST_LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _ip, _dp, C.c_int, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_properties.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.ref_field.restype = _dp
        L.ref_field.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_get_constants.argtypes = [C.c_void_p, _dp]
        L.ref_set_logging.argtypes = [C.c_int]
        L.ref_get_trace.restype = C.c_int
        L.ref_get_trace.argtypes = [C.c_char_p, C.c_int]
        L.ref_get_counts.argtypes = [C.POINTER(C.c_long)]
        L.ref_get_dt.argtypes = [_dp, C.c_void_p]
        L.ref_get_bbox.restype = C.c_int
        L.ref_get_bbox.argtypes = [_dp, C.c_int]
        L.ref_step.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, C.c_int, _ip, _ip, _dp]
        L.ref_set_source_zero.argtypes = [C.c_void_p]
        _host = L
    return _host


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def mt19937_64_uniform(seed, n):
    out = np.empty(n, dtype=np.float64)
    lib().ref_mt19937_64_uniform(seed, n, _d(out))
    return out


FIELD_WIDTH = dict(U=3, gradP=3, divT=3, ddtU=3, uSource=3, uParticle=3, vGrad=9, uSourceDrag=1, alpha=1, p=1)


class RefFoamYade:
    """The reference Foam::FoamYade object on a mesh given by oracle.meshgen.hex_box()."""

    def __init__(self, mesh, gaussian, n_yade=1, host=False):
        self.L = host_lib() if host else lib()
        self.mesh = mesh
        self.N = mesh["V"].shape[0]
        self.gaussian = bool(gaussian)
        self.n_yade = n_yade
        self.h = self.L.ref_create(self.N, _d(mesh["C"]), _d(mesh["V"]), mesh["points"].shape[0], _d(mesh["points"]),
                                   _i(mesh["boxN"]), _d(mesh["boxGeom"]), int(gaussian), n_yade)

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def set_properties(self, rhoP, rhoF, nu):
        self.L.ref_set_properties(self.h, rhoP, rhoF, nu)

    def field(self, name):
        """numpy VIEW of the reference's field storage."""
        w = FIELD_WIDTH[name]
        p = self.L.ref_field(self.h, name.encode())
        a = np.ctypeslib.as_array(p, shape=(self.N * w,))
        return a.reshape(self.N, w) if w > 1 else a

    def constants(self):
        out = np.empty(4)
        self.L.ref_get_constants(self.h, _d(out))
        return dict(interpRange=out[0], sigmaInterp=out[1], interpRangeCu=out[2], sigmaPi=out[3])

    def locate(self, xyz, stride=16):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        n = xyz.shape[0]
        ids = np.empty((n, stride), dtype=np.int32)
        cnt = np.empty(n, dtype=np.int32)
        self.L.ref_locate(self.h, _d(xyz), n, _i(ids), _i(cnt), stride)
        return cnt, ids

    def step(self, dt, pdata, yade_dt=0.0, split=None, pieces=False, truncate12=True, dense=True):
        pdata = np.ascontiguousarray(pdata, dtype=np.float64)
        n = pdata.shape[0]
        found = np.empty(max(n, 1), dtype=np.int32)
        force = np.empty((max(n, 1), 6), dtype=np.float64)
        sp = None if split is None else np.ascontiguousarray(split, dtype=np.int32)
        spp = _i(sp) if sp is not None else None
        if pieces:
            self.L.ref_step_pieces(self.h, dt, yade_dt, _d(pdata), n, spp, int(truncate12), int(dense), _i(found), _d(force))
        else:
            self.L.ref_step(self.h, dt, yade_dt, _d(pdata), n, spp, _i(found), _d(force))
        return found[:n], force[:n]

    def set_gaussian_options(self, support_full=False, added_mass=False, torque=False):
        """SURVEY 8(f)3 options for step(pieces=True): full-support cell lists (every cell within the search bound, fed to
        the reference's own weight / force code), addedMassForce (FoamYade.C:392-413), Gaussian torque (FoamYade.C:467-478)"""
        self.L.ref_set_gaussian_options(self.h, int(support_full), int(added_mass), int(torque))

    def lists(self, n):
        cnt = np.empty(n, dtype=np.int32)
        ids = np.empty((n, 16), dtype=np.int32)
        self.L.ref_get_lists(self.h, n, _i(cnt), _i(ids))
        return cnt, ids

    def times(self):
        out = np.empty(5)
        self.L.ref_get_times(self.h, _d(out))
        return dict(locate=out[0], weights=out[1], accum=out[2], force=out[3], send=out[4])

    def set_source_zero(self):
        self.L.ref_set_source_zero(self.h)

    # ---- host-class build only: the solver's loop body with the fluid step on the device (icoFoamYadeB200.H)
    def set_fv_mesh(self, mo):
        """LDU addressing, face geometry, patches and the patch types of U / p from an oracle.meshgen.hex_box_ldu() dict;
        call before set_properties (the engine is created there)."""
        L = self.L
        L.ref_set_fv_mesh.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_int, _ip, _ip, _dp, _dp, _dp, _ip, _dp, _ip, _dp]
        pl = mo["patches"]
        c = lambda a, dt=np.float64: np.ascontiguousarray(a, dtype=dt)
        sizes = c([p["faceCells"].size for p in pl], np.int32)
        keep = [c(mo["owner"], np.int32), c(mo["neighbour"], np.int32), c(mo["Sf"]), c(mo["magSf"]), c(mo["weights"]),
                c(mo["deltaCoeffs"]), sizes, c(np.concatenate([p["faceCells"] for p in pl]), np.int32),
                c(np.concatenate([p["Sf"] for p in pl])), c(np.concatenate([p["magSf"] for p in pl])),
                c(np.concatenate([p["deltaCoeffs"] for p in pl])), c([p["bcU"] for p in pl], np.int32),
                c([p["valueU"] for p in pl]).reshape(-1), c([p["bcP"] for p in pl], np.int32), c([p["valueP"] for p in pl])]
        L.ref_set_fv_mesh(self.h, int(mo["nInternalFaces"]), _i(keep[0]), _i(keep[1]), _d(keep[2]), _d(keep[3]), _d(keep[4]),
                          _d(keep[5]), len(pl), _i(keep[6]), _i(keep[7]), _d(keep[8]), _d(keep[9]), _d(keep[10]), _i(keep[11]),
                          _d(keep[12]), _i(keep[13]), _d(keep[14]))

    def set_piso(self, nCorrectors=2, nNonOrthogonalCorrectors=0, momentumPredictor=1):
        self.L.ref_set_piso.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.L.ref_set_piso(self.h, nCorrectors, nNonOrthogonalCorrectors, momentumPredictor)

    def fluid_step(self, solver, dt, pdata, g=(0.0, 0.0, 0.0), yade_dt=0.0):
        """one pass of the solver's loop body on the device: 'ico' (icoFoamYade.C:65-149) or 'pimple'"""
        L = self.L
        L.ref_fluid_step.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _dp, C.c_int, _ip, _dp, _ip, _dp, _dp]
        pdata = np.ascontiguousarray(pdata, dtype=np.float64)
        n = pdata.shape[0]
        found = np.empty(max(n, 1), dtype=np.int32)
        force = np.empty((max(n, 1), 6), dtype=np.float64)
        out = np.zeros(10)
        gv = np.ascontiguousarray(g, dtype=np.float64)
        L.ref_fluid_step(self.h, 0 if solver == "ico" else 1, dt, yade_dt, _d(pdata), n, None, _d(gv), _i(found), _d(force), _d(out))
        return found[:n], force[:n], dict(p_iters=[int(x) for x in out[:int(out[8])]], CoNum=out[9])

    def download_fluid(self):
        self.L.ref_download_fluid.argtypes = [C.c_void_p]
        self.L.ref_download_fluid(self.h)

    def host_gaussian_options(self, support_full=False, added_mass=False, torque=False):
        """host-class build only: FoamYadeB200::setGaussianOptions"""
        self.L.ref_host_gaussian_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.L.ref_host_gaussian_options(self.h, int(support_full), int(added_mass), int(torque))

    def host_pimple_controls(self, nOuterCorrectors=1, relaxU=0.0, relaxUFinal=0.0, relaxP=0.0, relaxPFinal=0.0):
        """host-class build only: FoamYadeB200::setPimpleControls"""
        self.L.ref_host_pimple_controls.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        self.L.ref_host_pimple_controls(self.h, int(nOuterCorrectors), relaxU, relaxUFinal, relaxP, relaxPFinal)

    def set_batched_wire(self, on=True):
        """host-class build only: one message per direction and step (F1)"""
        self.L.ref_set_batched_wire.argtypes = [C.c_void_p, C.c_int]
        self.L.ref_set_batched_wire(self.h, int(on))

    def realloc_fields(self):
        """moves every solver-owned field to newly allocated storage (views taken before are stale afterwards)"""
        self.L.ref_realloc_fields.argtypes = [C.c_void_p]
        self.L.ref_realloc_fields(self.h)

    def dts(self):
        out = np.empty(2)
        self.L.ref_get_dt(_d(out), self.h)
        return out[0], out[1]

    def counts(self):
        out = (C.c_long * 5)()
        self.L.ref_get_counts(out)
        return dict(bcast=out[0], allreduce=out[1], send=out[2], recv=out[3], isend=out[4])

    def trace(self):
        n = self.L.ref_get_trace(None, 0)
        buf = C.create_string_buffer(n + 1)
        self.L.ref_get_trace(buf, n + 1)
        return buf.value.decode().splitlines()


def fnv1a_lists(cnt, ids):
    """FNV-1a-64 over the stream of cell ids (u32 each, XOR-then-multiply on the whole word) with 0xffffffff
    appended after each particle."""
    h = 0xcbf29ce484222325
    M = (1 << 64) - 1
    for i in range(len(cnt)):
        for j in range(int(cnt[i])):
            h = ((h ^ (int(ids[i, j]) & 0xffffffff)) * 0x100000001b3) & M
        h = ((h ^ 0xffffffff) * 0x100000001b3) & M
    return h
