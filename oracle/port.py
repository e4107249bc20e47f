"""oracle/port.py -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper of oracle/_build/liboracle.so (oracle/fv_oracle.cc): this repo's CPU restatement of the
OpenFOAM-6 finite-volume operators and linear solvers behind icoFoamYade.C:65-149 and pimpleFoamYade/.
Build with `make -C oracle oracle` (done by __graft_entry__.build(); the built .so travels to the GPU box).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

PRECOND = dict(DIC=0, diagonal=1, none=2)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.fvo_create.restype = C.c_void_p
        L.fvo_create.argtypes = [C.c_int, _dp, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_int, _ip, _ip, _dp, _dp, _dp,
                                 _ip, _dp, _ip, _dp]
        L.fvo_destroy.argtypes = [C.c_void_p]
        L.fvo_set_controls.argtypes = [C.c_void_p, _ip, _dp]
        L.fvo_set_pimple_controls.argtypes = [C.c_void_p, C.c_int, _dp]
        L.fvo_field.restype = _dp
        L.fvo_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
        L.fvo_create_phi.argtypes = [C.c_void_p]
        L.fvo_ico_pre.argtypes = [C.c_void_p, C.c_double]
        L.fvo_ico_solve.restype = C.c_int
        L.fvo_ico_solve.argtypes = [C.c_void_p, C.c_double]
        L.fvo_get_stats.argtypes = [C.c_void_p, _dp]
        L.fvo_get_times.argtypes = [C.c_void_p, _dp]
        L.fvo_grad_vector.argtypes = [C.c_void_p, _dp, _dp]
        L.fvo_grad_scalar.argtypes = [C.c_void_p, _dp, _dp]
        L.fvo_div_flux.argtypes = [C.c_void_p, _dp, _dp]
        L.fvo_div_phi_vector.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.fvo_laplacian_gamma_vector.argtypes = [C.c_void_p, _dp, C.c_double, _dp, _dp]
        L.fvo_pimple_solve.restype = C.c_int
        L.fvo_pimple_solve.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp, _dp]
        L.fvo_pimple_field.restype = _dp
        L.fvo_pimple_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
        L.fvo_reconstruct.argtypes = [C.c_void_p, _dp, _dp]
        L.fvo_div_dev.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double, _dp, _dp]
        L.fvo_pimple_pre.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp, _dp, _dp]
        L.fvo_pcg.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp]
        L.fvo_smooth.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int, _dp]
        L.fvo_dic.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.fvo_amul.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
        L.fvo_set_partition.argtypes = [C.c_void_p, _ip]
        L.fvo_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


DEFAULT_CTL = dict(nCorrectors=2, nNonOrthogonalCorrectors=0, momentumPredictor=1, pRefCell=0, pRefValue=0.0,
                   pTol=1e-6, pRelTol=0.05, pFinalTol=1e-6, pFinalRelTol=0.0, UTol=1e-5, URelTol=0.0, maxIter=1000,
                   preconditioner="DIC", nu=0.01)


def parse_stats(out):
    st = dict(CoNum=out[0], meanCoNum=out[1], sumLocalContErr=out[2], globalContErr=out[3], cumulativeContErr=out[4],
              nPSolves=int(out[5]))
    st["U"] = [dict(initial=out[6 + 3 * j], final=out[7 + 3 * j], iters=int(out[8 + 3 * j])) for j in range(3)]
    st["p"] = [dict(initial=out[15 + 3 * k], final=out[16 + 3 * k], iters=int(out[17 + 3 * k]))
               for k in range(min(st["nPSolves"], 8))]
    st["corrSumLocal"] = [float(v) for v in out[39:47]]
    st["corrGlobal"] = [float(v) for v in out[47:55]]
    return st


class IcoOracle:
    """icoFoamYade's fluid step on a mesh dict from oracle.meshgen.hex_box_ldu()."""

    def __init__(self, mesh, **ctl):
        self.L = lib()
        self.mesh = mesh
        self.N = int(mesh["nCells"])
        self.Fi = int(mesh["nInternalFaces"])
        pl = mesh["patches"]
        self.nB = int(sum(p["faceCells"].size for p in pl))
        sizes = _c([p["faceCells"].size for p in pl], np.int32)
        bCell = _c(np.concatenate([p["faceCells"] for p in pl]), np.int32)
        bSf = _c(np.concatenate([p["Sf"] for p in pl]))
        bMag = _c(np.concatenate([p["magSf"] for p in pl]))
        bDc = _c(np.concatenate([p["deltaCoeffs"] for p in pl]))
        bcU = _c([p["bcU"] for p in pl], np.int32)
        bcP = _c([p["bcP"] for p in pl], np.int32)
        vU = _c([p["valueU"] for p in pl]).reshape(-1)
        vP = _c([p["valueP"] for p in pl])
        self.h = self.L.fvo_create(self.N, _d(_c(mesh["V"])), self.Fi, _i(_c(mesh["owner"], np.int32)),
                                   _i(_c(mesh["neighbour"], np.int32)), _d(_c(mesh["Sf"])), _d(_c(mesh["magSf"])),
                                   _d(_c(mesh["weights"])), _d(_c(mesh["deltaCoeffs"])), len(pl), _i(sizes), _i(bCell),
                                   _d(bSf), _d(bMag), _d(bDc), _i(bcU), _d(vU), _i(bcP), _d(vP))
        self.ctl = dict(DEFAULT_CTL)
        self.set_controls(**ctl)

    def close(self):
        if self.h:
            self.L.fvo_destroy(self.h)
            self.h = None

    def set_slabs(self, nslabs):
        """Decomposed-run semantics (decomposePar `simple` along z): nslabs z slabs with the planes
        [r nz / nslabs, (r+1) nz / nslabs); the DIC preconditioner of the pressure solve then factorises each slab's
        own matrix, as OpenFOAM does per processor.  nslabs <= 1 restores the single domain."""
        if nslabs <= 1:
            self.L.fvo_set_partition(self.h, None)
            return
        nx, ny, nz = (int(v) for v in self.mesh["boxN"])
        k = np.arange(self.N) // (nx * ny)
        bounds = [(r * nz) // nslabs for r in range(nslabs + 1)]
        proc = np.searchsorted(np.asarray(bounds[1:]), k, side="right").astype(np.int32)
        self.L.fvo_set_partition(self.h, _i(_c(proc, np.int32)))

    def set_grid(self, py, pz):
        """Decomposed-run semantics on a Py x Pz grid of processors (decomposePar `simple`, n (1 Py Pz), with the y cuts
        on multiples of 32 rows as csrc/fv_dist.cu places them): processor rz*Py + ry owns the 32-row blocks
        [ry nJB / Py, (ry+1) nJB / Py) of the planes [rz nz / Pz, (rz+1) nz / Pz)."""
        if py * pz <= 1:
            self.L.fvo_set_partition(self.h, None)
            return
        nx, ny, nz = (int(v) for v in self.mesh["boxN"])
        c = np.arange(self.N)
        k, jb = c // (nx * ny), ((c // nx) % ny) // 32
        njb = (ny + 31) // 32
        kb = np.asarray([((r + 1) * nz) // pz for r in range(pz)])
        jbb = np.asarray([((r + 1) * njb) // py for r in range(py)])
        proc = (np.searchsorted(kb, k, side="right") * py + np.searchsorted(jbb, jb, side="right")).astype(np.int32)
        self.L.fvo_set_partition(self.h, _i(_c(proc, np.int32)))

    def set_threads(self, n):
        """host threads of the decomposed PCG's TIMING path (bench.py's all-core CPU baseline); needs set_slabs(>1).
        Process-wide; 1 restores the sequential checker."""
        self.L.fvo_set_threads(int(n))

    def set_controls(self, **kw):
        self.ctl.update(kw)
        c = self.ctl
        ic = _c([c["nCorrectors"], c["nNonOrthogonalCorrectors"], c["momentumPredictor"], c["pRefCell"], c["maxIter"],
                 PRECOND[c["preconditioner"]]], np.int32)
        dc = _c([c["pRefValue"], c["pTol"], c["pRelTol"], c["pFinalTol"], c["pFinalRelTol"], c["UTol"], c["URelTol"],
                 c["nu"]])
        self.L.fvo_set_controls(self.h, _i(ic), _d(dc))

    def set_pimple_controls(self, nOuterCorrectors=1, relaxU=0.0, relaxUFinal=0.0, relaxP=0.0, relaxPFinal=0.0):
        """PIMPLE nOuterCorrectors (pimpleFoamYade.C:91 `while (pimple.loop())`) and fvSolution relaxationFactors
        (UcEqn.relax() pim/UcEqn.H:13, p.relax() pim/pEqn.H:41); a factor <= 0 means `no entry` (no-op)."""
        self.L.fvo_set_pimple_controls(self.h, int(nOuterCorrectors), _d(_c([relaxU, relaxUFinal, relaxP, relaxPFinal])))

    def field(self, name):
        """numpy VIEW of a state / intermediate field (shape by name)."""
        n = C.c_long()
        p = self.L.fvo_field(self.h, name.encode(), C.byref(n))
        if not p or n.value == 0:
            raise KeyError(name)
        a = np.ctypeslib.as_array(p, shape=(n.value,))
        if name in ("U", "uSource", "HbyA", "gradP", "sourceU", "icU", "bcU"):
            return a.reshape(-1, 3)
        if name == "vGrad":
            return a.reshape(-1, 9)
        return a

    def create_phi(self):
        self.L.fvo_create_phi(self.h)

    def pre(self, dt):
        self.L.fvo_ico_pre(self.h, dt)

    def solve(self, dt):
        rc = self.L.fvo_ico_solve(self.h, dt)
        if rc != 0:
            raise RuntimeError("adjustPhi: continuity error cannot be removed by adjusting the outflow")

    def stats(self):
        out = np.zeros(64)
        self.L.fvo_get_stats(self.h, _d(out))
        return parse_stats(out)

    def times(self):
        out = np.zeros(3)
        self.L.fvo_get_times(self.h, _d(out))
        return dict(momentum=out[0], pressure=out[1], other=out[2])

    # stand-alone operators
    def grad_vector(self, U):
        out = np.empty((self.N, 9))
        self.L.fvo_grad_vector(self.h, _d(_c(U)), _d(out))
        return out

    def grad_scalar(self, p):
        out = np.empty((self.N, 3))
        self.L.fvo_grad_scalar(self.h, _d(_c(p)), _d(out))
        return out

    def div_flux(self, phi):
        out = np.empty(self.N)
        self.L.fvo_div_flux(self.h, _d(_c(phi)), _d(out))
        return out

    def div_phi_vector(self, phi, U):
        out = np.empty((self.N, 3))
        self.L.fvo_div_phi_vector(self.h, _d(_c(phi)), _d(_c(U)), _d(out))
        return out

    def laplacian_gamma_vector(self, gamma, U, gammaB=1.0):
        out = np.empty((self.N, 3))
        self.L.fvo_laplacian_gamma_vector(self.h, _d(_c(gamma)), gammaB, _d(_c(U)), _d(out))
        return out

    def reconstruct(self, ssf):
        out = np.empty((self.N, 3))
        self.L.fvo_reconstruct(self.h, _d(_c(ssf)), _d(out))
        return out

    def div_dev(self, alpha, U, nu, alphaB=1.0):
        """fvc::div((alpha*nuEff)*dev2(T(fvc::grad(U)))) -- the explicit part of the laminar divDevRhoReff"""
        out = np.empty((self.N, 3))
        self.L.fvo_div_dev(self.h, _d(_c(alpha)), alphaB, nu, _d(_c(U)), _d(out))
        return out

    def pimple_solve(self, dt, alpha, uSourceDrag, g=(0.0, 0.0, 0.0), alpha0=None):
        """UcEqn.H + pEqn.H + continuityErrs.H (pimpleFoamYade.C:82-104); reads the state's uSource"""
        a = _c(alpha).reshape(-1)
        a0 = a if alpha0 is None else _c(alpha0).reshape(-1)
        rc = self.L.fvo_pimple_solve(self.h, dt, _d(a), _d(a0), _d(_c(uSourceDrag).reshape(-1)), _d(_c(g)))
        if rc == -2:
            raise RuntimeError("p.relax(): previous-iteration field not stored (nOuterCorrectors 1 with a p relaxation factor < 1)")
        if rc != 0:
            raise RuntimeError("adjustPhi: continuity error cannot be removed by adjusting the outflow")

    def pimple_field(self, name):
        n = C.c_long()
        p = self.L.fvo_pimple_field(self.h, name.encode(), C.byref(n))
        if not p or n.value == 0:
            raise KeyError(name)
        a = np.ctypeslib.as_array(p, shape=(n.value,))
        return a.reshape(-1, 3) if name in ("recon", "divDev") else a

    def pimple_pre(self, dt, alpha):
        """pimpleFoamYade.C:73-76 on the state's U, p, phi: returns ddtU, gradP, divT, vGrad."""
        ddtU, gradP, divT = np.empty((self.N, 3)), np.empty((self.N, 3)), np.empty((self.N, 3))
        vGrad = np.empty((self.N, 9))
        self.L.fvo_pimple_pre(self.h, dt, _d(_c(alpha)), _d(ddtU), _d(gradP), _d(divT), _d(vGrad))
        return ddtU, gradP, divT, vGrad

    def pcg(self, diag, upper, source, psi0, tol=1e-6, relTol=0.0, maxIter=1000, preconditioner="DIC"):
        psi = _c(psi0).copy()
        out = np.zeros(3)
        self.L.fvo_pcg(self.h, _d(_c(diag)), _d(_c(upper)), _d(_c(source)), _d(psi), tol, relTol, maxIter,
                       PRECOND[preconditioner], _d(out))
        return psi, dict(initial=out[0], final=out[1], iters=int(out[2]))

    def smooth(self, diag, lower, upper, source, psi0, tol=1e-5, relTol=0.0, maxIter=1000):
        psi = _c(psi0).copy()
        out = np.zeros(3)
        self.L.fvo_smooth(self.h, _d(_c(diag)), _d(_c(lower)), _d(_c(upper)), _d(_c(source)), _d(psi), tol, relTol,
                          maxIter, _d(out))
        return psi, dict(initial=out[0], final=out[1], iters=int(out[2]))

    def dic(self, diag, upper, rA):
        out = np.empty(self.N)
        self.L.fvo_dic(self.h, _d(_c(diag)), _d(_c(upper)), _d(_c(rA)), _d(out))
        return out

    def amul(self, diag, lower, upper, psi):
        out = np.empty(self.N)
        self.L.fvo_amul(self.h, _d(_c(diag)), _d(_c(lower)), _d(_c(upper)), _d(_c(psi)), _d(out))
        return out
