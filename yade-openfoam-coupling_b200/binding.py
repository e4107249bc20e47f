"""ctypes binding of libfycuda.so (include/fycuda.h).  No CPU fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

FIELD = dict(U=0, gradP=1, vGrad=2, divT=3, ddtU=4, uSourceDrag=5, alpha=6, uSource=7, uParticle=8, p=9, phi=10)
_WIDTH = dict(U=3, gradP=3, vGrad=9, divT=3, ddtU=3, uSourceDrag=1, alpha=1, uSource=3, uParticle=3, p=1)


class FyError(RuntimeError):
    pass


class _PatchDesc(C.Structure):
    _fields_ = [("nFaces", C.c_int), ("faceCells", _ip), ("Sf", _dp), ("magSf", _dp), ("deltaCoeffs", _dp),
                ("bcU", C.c_int), ("valueU", C.c_double * 3), ("bcP", C.c_int), ("valueP", C.c_double)]


class PisoControls(C.Structure):
    _fields_ = [("nCorrectors", C.c_int), ("nNonOrthogonalCorrectors", C.c_int), ("momentumPredictor", C.c_int),
                ("pRefCell", C.c_int), ("pRefValue", C.c_double), ("pTol", C.c_double), ("pRelTol", C.c_double),
                ("pFinalTol", C.c_double), ("pFinalRelTol", C.c_double), ("UTol", C.c_double), ("URelTol", C.c_double),
                ("maxIter", C.c_int), ("preconditioner", C.c_int)]


class _SolverPerf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double), ("nIterations", C.c_int),
                ("pad_", C.c_int)]


class _IcoStats(C.Structure):
    _fields_ = [("CoNum", C.c_double), ("meanCoNum", C.c_double), ("U", _SolverPerf * 3), ("p", _SolverPerf * 8),
                ("nPSolves", C.c_int), ("pad_", C.c_int), ("sumLocalContErr", C.c_double),
                ("globalContErr", C.c_double), ("cumulativeContErr", C.c_double), ("corrSumLocal", C.c_double * 8),
                ("corrGlobal", C.c_double * 8)]


PRECOND = dict(DIC=0, diagonal=1, none=2)


class _MeshDesc(C.Structure):
    _fields_ = [("nCells", C.c_int), ("C", _dp), ("V", _dp), ("nInternalFaces", C.c_int), ("owner", _ip),
                ("neighbour", _ip), ("Sf", _dp), ("magSf", _dp), ("weights", _dp), ("deltaCoeffs", _dp),
                ("nPatches", C.c_int), ("patches", C.POINTER(_PatchDesc)), ("boxN", C.c_int * 3),
                ("boxGeom", C.c_double * 6), ("bbox", C.c_double * 6)]


def lib_path():
    # FY_LIBFYCUDA: a differently-built libfycuda.so (dev A/B runs); the product path is the in-tree library
    return os.environ.get("FY_LIBFYCUDA") or os.path.join(_HERE, "libfycuda.so")


_lib = None


def lib():
    """Loads libfycuda.so; raises FyError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise FyError("libfycuda.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "or `make -C yade-openfoam-coupling_b200`); this engine has no CPU fallback")
    L = C.CDLL(p)
    H = C.c_void_p
    L.fy_device_count.restype = C.c_int
    L.fy_version.restype = C.c_char_p
    L.fy_last_error.restype = C.c_char_p
    L.fy_last_error.argtypes = [H]
    L.fy_create.argtypes = [C.POINTER(_MeshDesc), C.c_int, C.POINTER(H)]
    L.fy_destroy.argtypes = [H]
    L.fy_set_properties.argtypes = [H, C.c_double, C.c_double, C.c_double, C.c_int]
    L.fy_get_constants.argtypes = [H, _dp]
    L.fy_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.fy_host_free.argtypes = [C.c_void_p]
    L.fy_bind_host_fields.argtypes = [H] + [_dp] * 9
    L.fy_upload_field.argtypes = [H, C.c_int, _dp]
    L.fy_download_field.argtypes = [H, C.c_int, _dp]
    L.fy_device_field.argtypes = [H, C.c_int, C.POINTER(C.c_void_p)]
    L.fy_locate.argtypes = [H, _dp, C.c_int, _ip, _ip]
    L.fy_find_cell.argtypes = [H, _dp, C.c_int, _ip]
    L.fy_coupling_begin.argtypes = [H, C.c_double]
    L.fy_coupling_proc.argtypes = [H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.fy_coupling_end.argtypes = [H]
    L.fy_set_particle_action.argtypes = [H, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.fy_coupling_proc_device.argtypes = [H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.fy_set_source_zero.argtypes = [H]
    L.fy_coupling_pass_device.argtypes = [H, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.fy_device_accumulators.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.fy_stream.argtypes = [H, C.POINTER(C.c_void_p)]
    L.fy_get_last_lists.argtypes = [H, C.c_int, _ip, _ip, _dp]
    L.fy_synchronize.argtypes = [H]
    L.fy_set_profiling.argtypes = [H, C.c_int]
    L.fy_get_phase_ms.argtypes = [H, _dp]
    L.fy_timer_start.argtypes = [H]
    L.fy_timer_stop.argtypes = [H, _dp]
    L.fy_launch_count.restype = C.c_longlong
    L.fy_launch_count.argtypes = [H]
    L.fy_fv_supported.argtypes = [H]
    L.fy_piso_default_controls.argtypes = [C.POINTER(PisoControls)]
    L.fy_set_piso_controls.argtypes = [H, C.POINTER(PisoControls)]
    L.fy_set_viscosity.argtypes = [H, C.c_double]
    L.fy_set_gaussian_options.argtypes = [H, C.c_int, C.c_int, C.c_int]
    L.fy_set_pimple_controls.argtypes = [H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    L.fy_create_phi.argtypes = [H]
    L.fy_ico_pre.argtypes = [H, C.c_double]
    L.fy_ico_solve.argtypes = [H, C.c_double]
    L.fy_get_ico_stats.argtypes = [H, C.POINTER(_IcoStats)]
    L.fy_fvc_grad_vector.argtypes = [H, _dp, _dp]
    L.fy_fvc_grad_scalar.argtypes = [H, _dp, _dp]
    L.fy_fvc_div_flux.argtypes = [H, _dp, _dp]
    L.fy_fvc_div_phi_vector.argtypes = [H, _dp, _dp, _dp]
    L.fy_fvc_laplacian_gamma_vector.argtypes = [H, _dp, C.c_double, _dp, _dp]
    L.fy_pimple_pre.argtypes = [H, C.c_double]
    L.fy_pimple_solve.argtypes = [H, C.c_double, _dp]
    L.fy_pcg_solve.argtypes = [H, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp]
    L.fy_smooth_solve.argtypes = [H, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, _dp]
    L.fy_dic_precondition.argtypes = [H, _dp, _dp, _dp, _dp]
    L.fy_fv_get.argtypes = [H, C.c_char_p, _dp]
    L.fy_get_fluid_ms.argtypes = [H, _dp]
    L.fy_get_kernel_ms.argtypes = [H, _dp, C.c_int]
    _lib = L
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_mesh_desc(mesh):
    """Returns (desc, keepalive) for a dict from mesh.box_mesh()."""
    keep = []

    def d(a):
        a = _c64(a)
        keep.append(a)
        return _d(a)

    def i(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return _i(a)

    md = _MeshDesc()
    md.nCells = int(mesh["nCells"])
    md.C = d(mesh["C"])
    md.V = d(mesh["V"])
    md.nInternalFaces = int(mesh.get("nInternalFaces", 0))
    if md.nInternalFaces > 0:
        md.owner = i(mesh["owner"])
        md.neighbour = i(mesh["neighbour"])
        md.Sf = d(mesh["Sf"])
        md.magSf = d(mesh["magSf"])
        md.weights = d(mesh["weights"])
        md.deltaCoeffs = d(mesh["deltaCoeffs"])
    pl = mesh.get("patches", []) if md.nInternalFaces > 0 else []
    md.nPatches = len(pl)
    if pl:
        arr = (_PatchDesc * len(pl))()
        for n, p in enumerate(pl):
            arr[n].nFaces = int(p["faceCells"].shape[0])
            arr[n].faceCells = i(p["faceCells"])
            arr[n].Sf = d(p["Sf"])
            arr[n].magSf = d(p["magSf"])
            arr[n].deltaCoeffs = d(p["deltaCoeffs"])
            arr[n].bcU = int(p["bcU"])
            arr[n].valueU = (C.c_double * 3)(*p["valueU"])
            arr[n].bcP = int(p["bcP"])
            arr[n].valueP = float(p["valueP"])
        keep.append(arr)
        md.patches = arr
    md.boxN = (C.c_int * 3)(*[int(v) for v in mesh["boxN"]])
    md.boxGeom = (C.c_double * 6)(*[float(v) for v in mesh["boxGeom"]])
    md.bbox = (C.c_double * 6)(*[float(v) for v in mesh["bbox"]])
    return md, keep


class Engine:
    """One fy_handle.  Mirrors the reference's operator surface (FoamYade.H:106-155):
    set_properties ~ setScalarProperties, set_particle_action ~ setParticleAction,
    set_source_zero ~ setSourceZero."""

    def __init__(self, mesh, device=0):
        self.L = lib()
        self.mesh = mesh
        self.N = int(mesh["nCells"])
        md, self._keep = make_mesh_desc(mesh)
        h = C.c_void_p()
        rc = self.L.fy_create(C.byref(md), device, C.byref(h))
        if rc != 0:
            raise FyError("fy_create failed (%d): %s" % (rc, self.L.fy_last_error(None).decode()))
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise FyError("fycuda error %d: %s" % (rc, self.L.fy_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.fy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_properties(self, rhoP, rhoF, nu, gaussian):
        self._ck(self.L.fy_set_properties(self.h, rhoP, rhoF, nu, int(gaussian)))
        self.gaussian = bool(gaussian)

    def constants(self):
        out = np.empty(4)
        self._ck(self.L.fy_get_constants(self.h, _d(out)))
        return dict(interpRange=out[0], sigmaInterp=out[1], interpRangeCu=out[2], sigmaPi=out[3])

    def upload(self, name, arr):
        a = _c64(arr)
        self._ck(self.L.fy_upload_field(self.h, FIELD[name], _d(a)))

    def download(self, name, n=None):
        if name == "phi":
            if n is None:
                n = int(self.mesh["nInternalFaces"]) + sum(int(p["faceCells"].shape[0]) for p in self.mesh["patches"])
            out = np.empty(n, dtype=np.float64)
        else:
            w = _WIDTH[name]
            out = np.empty((self.N, w) if w > 1 else (self.N,), dtype=np.float64)
        self._ck(self.L.fy_download_field(self.h, FIELD[name], _d(out)))
        return out

    def device_field(self, name):
        p = C.c_void_p()
        self._ck(self.L.fy_device_field(self.h, FIELD[name], C.byref(p)))
        return p.value

    def locate(self, xyz):
        xyz = _c64(xyz)
        n = xyz.shape[0]
        ids = np.empty((n, 12), dtype=np.int32)
        cnt = np.empty(n, dtype=np.int32)
        self._ck(self.L.fy_locate(self.h, _d(xyz), n, _i(ids), _i(cnt)))
        return cnt, ids

    def find_cell(self, xyz):
        xyz = _c64(xyz)
        n = xyz.shape[0]
        cell = np.empty(n, dtype=np.int32)
        self._ck(self.L.fy_find_cell(self.h, _d(xyz), n, _i(cell)))
        return cell

    def set_particle_action(self, dt, pdata, found=None, force=None):
        pdata = _c64(pdata)
        n = pdata.shape[0]
        if found is None:
            found = np.empty(n, dtype=np.int32)
        if force is None:
            force = np.empty((n, 6), dtype=np.float64)
        self._ck(self.L.fy_set_particle_action(self.h, dt, pdata.ctypes.data, n, found.ctypes.data, force.ctypes.data))
        return found, force

    def coupling_begin(self, dt):
        self._ck(self.L.fy_coupling_begin(self.h, dt))

    def coupling_proc(self, pdata):
        pdata = _c64(pdata)
        n = pdata.shape[0]
        found = np.empty(n, dtype=np.int32)
        force = np.empty((n, 6), dtype=np.float64)
        self._ck(self.L.fy_coupling_proc(self.h, pdata.ctypes.data, n, found.ctypes.data, force.ctypes.data))
        return found, force

    def coupling_end(self):
        self._ck(self.L.fy_coupling_end(self.h))

    def coupling_proc_device(self, d_pdata, n, d_found, d_force):
        self._ck(self.L.fy_coupling_proc_device(self.h, d_pdata, n, d_found, d_force))

    def coupling_pass_device(self, p, d_pdata, n, d_found, d_force):
        self._ck(self.L.fy_coupling_pass_device(self.h, int(p), C.c_void_p(d_pdata), int(n), C.c_void_p(d_found),
                                                 C.c_void_p(d_force)))

    def device_accumulators(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.L.fy_device_accumulators(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def stream(self):
        s = C.c_void_p()
        self._ck(self.L.fy_stream(self.h, C.byref(s)))
        return s.value or 0

    def set_source_zero(self):
        self._ck(self.L.fy_set_source_zero(self.h))

    def last_lists(self, n):
        cnt = np.empty(n, dtype=np.int32)
        ids = np.empty((n, 12), dtype=np.int32)
        w = np.empty((n, 12), dtype=np.float64)
        self._ck(self.L.fy_get_last_lists(self.h, n, _i(cnt), _i(ids), _d(w)))
        return cnt, ids, w

    def last_counts(self, n):
        """cells per particle of the last Gaussian pass (the only list data the full-support mode keeps)"""
        cnt = np.empty(n, dtype=np.int32)
        self._ck(self.L.fy_get_last_lists(self.h, n, _i(cnt), None, None))
        return cnt

    # ---- multi-GPU: Py x Pz decomposition of the pressure solve (fycuda.h, fy_dist_*)
    @staticmethod
    def dist_unique_id():
        buf = C.create_string_buffer(128)
        rc = lib().fy_dist_unique_id(buf)
        if rc != 0:
            raise FyError("fy_dist_unique_id failed (%d): is libnccl.so.2 on the loader path?" % rc)
        return buf.raw

    def dist_init(self, rank, nranks, uid, py=0):
        """py = 0: the library picks the Py x Pz grid of ranks (y first); py = 1: z slabs only."""
        self._ck(self.L.fy_dist_init_grid(self.h, int(rank), int(nranks), int(py), C.c_char_p(bytes(uid))))

    def dist_info(self):
        out = (C.c_longlong * 6)()
        self._ck(self.L.fy_dist_info(self.h, out))
        g = (C.c_longlong * 10)()
        self._ck(self.L.fy_dist_grid(self.h, g))
        return dict(rank=out[0], nranks=out[1], kLo=out[2], kHi=out[3], collectives=out[4], halo_bytes=out[5],
                    Py=g[0], Pz=g[1], ry=g[2], rz=g[3], jLo=g[4], jHi=g[5], peer=bool(g[8]))

    def synchronize(self):
        self._ck(self.L.fy_synchronize(self.h))

    def set_profiling(self, on):
        self._ck(self.L.fy_set_profiling(self.h, int(on)))

    def phase_ms(self):
        out = np.zeros(8)
        self._ck(self.L.fy_get_phase_ms(self.h, _d(out)))
        return out

    def timer_start(self):
        self._ck(self.L.fy_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self.L.fy_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def set_gaussian_options(self, support_full=False, added_mass=False, torque=False):
        """full-support Gaussian cell sets (range-based search) and the reference's dormant forces (fycuda.h)"""
        self._ck(self.L.fy_set_gaussian_options(self.h, 1 if support_full else 0, int(added_mass), int(torque)))

    def launch_count(self):
        return int(self.L.fy_launch_count(self.h))

    # ---- the fluid half (icoFoamYade time step) -------------------------------------------------
    def fv_supported(self):
        return bool(self.L.fy_fv_supported(self.h))

    def set_piso_controls(self, **kw):
        c = PisoControls()
        self.L.fy_piso_default_controls(C.byref(c))
        for k, v in kw.items():
            if k == "preconditioner" and isinstance(v, str):
                v = PRECOND[v]
            if k == "nu":
                self._ck(self.L.fy_set_viscosity(self.h, float(v)))
                continue
            if not hasattr(c, k):
                raise KeyError(k)
            setattr(c, k, v)
        self._ck(self.L.fy_set_piso_controls(self.h, C.byref(c)))

    def set_pimple_controls(self, nOuterCorrectors=1, relaxU=0.0, relaxUFinal=0.0, relaxP=0.0, relaxPFinal=0.0):
        """PIMPLE nOuterCorrectors (pimpleFoamYade.C:91) and relaxationFactors (UcEqn.H:13, pEqn.H:41); <= 0: no entry"""
        self._ck(self.L.fy_set_pimple_controls(self.h, int(nOuterCorrectors), float(relaxU), float(relaxUFinal), float(relaxP),
                                               float(relaxPFinal)))

    def create_phi(self):
        self._ck(self.L.fy_create_phi(self.h))

    def ico_pre(self, dt):
        self._ck(self.L.fy_ico_pre(self.h, dt))

    def pimple_pre(self, dt):
        """pimpleFoamYade.C:71-76 on the device fields: CourantNo, ddtU, gradP, divT, vGrad"""
        self._ck(self.L.fy_pimple_pre(self.h, dt))

    def pimple_solve(self, dt, g=(0.0, 0.0, 0.0)):
        """pimpleFoamYade.C:82-104 (UcEqn.H, pEqn.H, continuityErrs.H) on the device fields"""
        gv = _c64(g)
        self._ck(self.L.fy_pimple_solve(self.h, dt, _d(gv)))

    def ico_solve(self, dt):
        self._ck(self.L.fy_ico_solve(self.h, dt))

    def fluid_step(self, dt):
        """the fluid part of one icoFoamYade time step after setParticleAction (icoFoamYade.C:79-140)"""
        self._ck(self.L.fy_ico_solve(self.h, dt))

    def ico_stats(self):
        st = _IcoStats()
        self._ck(self.L.fy_get_ico_stats(self.h, C.byref(st)))
        perf = lambda q: dict(initial=q.initialResidual, final=q.finalResidual, iters=q.nIterations)  # noqa: E731
        return dict(CoNum=st.CoNum, meanCoNum=st.meanCoNum, U=[perf(st.U[j]) for j in range(3)],
                    p=[perf(st.p[k]) for k in range(min(st.nPSolves, 8))], nPSolves=st.nPSolves,
                    sumLocalContErr=st.sumLocalContErr, globalContErr=st.globalContErr,
                    cumulativeContErr=st.cumulativeContErr, corrSumLocal=list(st.corrSumLocal),
                    corrGlobal=list(st.corrGlobal))

    def grad_vector(self, U):
        out = np.empty((self.N, 9))
        self._ck(self.L.fy_fvc_grad_vector(self.h, _d(_c64(U)), _d(out)))
        return out

    def grad_scalar(self, p):
        out = np.empty((self.N, 3))
        self._ck(self.L.fy_fvc_grad_scalar(self.h, _d(_c64(p)), _d(out)))
        return out

    def div_flux(self, phi):
        out = np.empty(self.N)
        self._ck(self.L.fy_fvc_div_flux(self.h, _d(_c64(phi)), _d(out)))
        return out

    def div_phi_vector(self, phi, U):
        out = np.empty((self.N, 3))
        self._ck(self.L.fy_fvc_div_phi_vector(self.h, _d(_c64(phi)), _d(_c64(U)), _d(out)))
        return out

    def laplacian_gamma_vector(self, gamma, U, gammaB=1.0):
        out = np.empty((self.N, 3))
        self._ck(self.L.fy_fvc_laplacian_gamma_vector(self.h, _d(_c64(gamma)), gammaB, _d(_c64(U)), _d(out)))
        return out

    def pcg(self, diag, upper, source, psi0, tol=1e-6, relTol=0.0, maxIter=1000, preconditioner="DIC"):
        psi = _c64(psi0).copy()
        out = np.zeros(3)
        self._ck(self.L.fy_pcg_solve(self.h, _d(_c64(diag)), _d(_c64(upper)), _d(_c64(source)), _d(psi), tol, relTol,
                                     maxIter, PRECOND[preconditioner], _d(out)))
        return psi, dict(initial=out[0], final=out[1], iters=int(out[2]))

    def smooth(self, diag, lower, upper, source, psi0, tol=1e-5, relTol=0.0, maxIter=1000):
        psi = _c64(psi0).copy()
        out = np.zeros(3)
        self._ck(self.L.fy_smooth_solve(self.h, _d(_c64(diag)), _d(_c64(lower)), _d(_c64(upper)), _d(_c64(source)),
                                        _d(psi), tol, relTol, maxIter, _d(out)))
        return psi, dict(initial=out[0], final=out[1], iters=int(out[2]))

    def dic(self, diag, upper, rA):
        out = np.empty(self.N)
        self._ck(self.L.fy_dic_precondition(self.h, _d(_c64(diag)), _d(_c64(upper)), _d(_c64(rA)), _d(out)))
        return out

    def fv_get(self, name):
        nF = int(self.mesh["nInternalFaces"])
        nB = sum(int(p["faceCells"].shape[0]) for p in self.mesh["patches"])
        shape = dict(rAU=(self.N,), HbyA=(self.N, 3), gradP=(self.N, 3), diagU=(self.N,), sourceU=(self.N, 3),
                     phiHbyA=(nF + nB,), phi=(nF + nB,), upperP=(nF,), upperU=(nF,), lowerU=(nF,),
                     phicForces=(nF + nB,), divDev=(self.N, 3), bGradP=(nF + nB,))[name]
        out = np.empty(shape)
        self._ck(self.L.fy_fv_get(self.h, name.encode(), _d(out)))
        return out

    def fluid_ms(self):
        out = np.zeros(4)
        self._ck(self.L.fy_get_fluid_ms(self.h, _d(out)))
        return out

    def kernel_ms(self, reset=True):
        out = np.zeros(8)
        self._ck(self.L.fy_get_kernel_ms(self.h, _d(out), int(reset)))
        return dict(precond_fwd=out[0], precond_bwd=out[1], direction=out[2], amul=out[3], update=out[4],
                    samples=int(out[5]), pcg_iterations=int(out[6]), fused_tail=bool(out[7]))
