#!/usr/bin/env python
"""bench.py -- coupled timesteps/s of the FoamYade hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--coupling gaussian|point]
  python bench.py --impl reference ...      # the reference's own CPU path (oracle/_ref + oracle port)

One "step" = one icoFoamYade time step (icoFoamYade.C:65-149) on one batch of synthetic particles:
  CourantNo, vGrad = grad(U)  ->  setParticleAction(dt)  ->  UEqn + momentum predictor + PISO correctors
  (pressure PCG)  ->  setSourceZero()
`value` times it with the particle records already resident in HBM (fy_coupling_proc_device); `e2e` times the
same step through the reference-facing call fy_set_particle_action with pinned HOST wire buffers
(80 B/particle in, 52 B/particle out inside the timed region).
Timing: CUDA events on the engine's own stream, max over ranks; L2 is flushed between timed steps.
N > 1 (default --partition domain): ONE domain and ONE particle set on N GPUs (strong scaling): the pressure solves are
z-slab decomposed (NCCL halo exchange + all-reduces, csrc/fv_dist.cu), the particles migrate every step to the rank that
owns their slab (all-to-all) and the per-cell coupling sums are all-reduced; see DESIGN.md "multi-GPU".
--partition replicas keeps round 1's N independent replicas (weak scaling, no collective on the data path).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "coupled timesteps/sec @1M particles/128^3 cells; achieved HBM GB/s vs peak"
WORKLOADS = {
    # name: (nx, ny, nz, particles, seed, flow, dt, nu, description)
    "C1": (32, 32, 32, 1000, 42, "cavity", 5e-3, 0.01, "icoFoamYade lid-driven cavity 32^3 cells, 1k particles"),
    "C2": (128, 128, 128, 1000000, 7, "channel", 5e-3, 1e-6, "icoFoamYade channel 128^3 cells, 1M particles, fp64, 1xB200"),
    "C2s": (64, 64, 64, 125000, 7, "channel", 1e-2, 1e-6, "icoFoamYade channel 64^3 cells, 125k particles (reduced C2)"),
    "C3": (256, 256, 256, 10000000, 1001, "closedbox", 2.5e-3, 1e-6, "pimpleFoamYade 4-way coupling, closed box under gravity, 256^3 cells, 10M particles, void fraction + UcEqn/pEqn, fp64, 1xB200 (use --solver pimple)"),
    "C3s": (128, 128, 128, 1000000, 1001, "closedbox", 5e-3, 1e-6, "pimpleFoamYade closed box under gravity, 128^3 cells, 1M particles (reduced C3; use --solver pimple)"),
    "C3c": (256, 256, 256, 10000000, 1001, "channel", 2.5e-3, 1e-6, "pimpleFoamYade channel 256^3 cells, 10M particles (round 1's C3 flow; use --solver pimple)"),
    "C5": (256, 256, 256, 100000000, 1003, "closedbox", 2.5e-3, 1e-6, "pimpleFoamYade settling suspension, closed box under gravity, 256^3 cells, 100M particles (use --solver pimple)"),
    "C2p": (128, 128, 128, 10000000, 7, "channel", 5e-3, 1e-6, "icoFoamYade channel 128^3 cells, 10M particles (particle-bound, for --partition particles)"),
    "C4": (512, 256, 256, 50000000, 1002, "channel", 1.25e-3, 1e-6, "icoFoamYade box 512x256x256 cells, 50M particles, domain-decomposed 8xB200 (BASELINE configs[3])"),
    "C4s": (256, 128, 128, 6250000, 1002, "channel", 2.5e-3, 1e-6, "icoFoamYade box 256x128x128 cells, 6.25M particles (C4 at 1/8 size)"),
}
UIN = 0.3
GRAVITY = (0.0, 0.0, -9.81)          # closed-box workloads (pimpleFoamYade's g, pimpleFoamYade/createFields.H readGravitationalAcceleration)


def gravity_of(wl):
    return GRAVITY if WORKLOADS[wl][5] == "closedbox" else (0.0, 0.0, 0.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(kernel_label):
    """(DRAM bytes per launch (read + write), source file) of the kernel class from the NEWEST committed ncu --set full
    capture that knows the kernel (profiles/r*_ncu_traffic.json, written by tools/ncu_summary.py --traffic from a
    capture of this command); (None, None) when no capture names it -- a stale figure is not reported."""
    import glob
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")), reverse=True):
        try:
            ks = json.load(open(p))["kernels"]
        except Exception:
            continue
        for name, v in ks.items():
            if kernel_label.startswith(name):
                return (v["dram_read_MB"] + v["dram_write_MB"]) * 1e6, os.path.relpath(p, ROOT)
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def solve_grid(ny, world, py=0):
    """(Py, Pz) of the decomposed pressure solve -- yade-openfoam-coupling_b200/domain.py solve_grid, restated here so that
    the reference arm (no package import needed) prints the same config."""
    njb = (int(ny) + 31) // 32
    if py == 0:
        py = max(c for c in range(1, world + 1) if world % c == 0 and c <= njb)
    return py, world // py


def partition_text(partition, world, ny=0, py=0):
    if world == 1:
        return "single domain"
    gy, gz = solve_grid(ny, world, py) if partition == "domain" else (1, world)
    return {"domain": "one domain on %d GPUs: pressure solve decomposed on a %d (y) x %d (z) grid of ranks (NCCL halo exchange of the "
                      "PCG search direction + 3 all-reduces per iteration, rank-local DIC), particles migrate to the owner z slab's "
                      "rank every step (all-to-all), per-cell coupling sums all-reduced, FV assembly replicated" % (world, gy, gz),
            "particles": "one domain, particle buffer sharded over the GPUs, NCCL all-reduce of the cell sums, fluid solve replicated",
            "replicas": "one domain replica per GPU"}[partition]


def config_of(args, world):
    """The `config` of the JSON line -- the SAME dict in the engine arm and in the reference arm (the reference arm runs
    the CPU path "on the engine arm's config")."""
    nx, ny, nz, P, seed, flow, dt, nu, desc = WORKLOADS[args.workload]
    N = nx * ny * nz
    pimple = args.solver == "pimple"
    replicas = args.partition == "replicas" and world > 1
    return {"workload": "%s: %s" % (args.workload, desc), "cells": N, "internal_faces": 3 * N - nx * ny - ny * nz - nx * nz,
            "particles_total": P * world if replicas else P,
            "coupling": args.coupling + (" (full support: every cell inside the search bound)" if getattr(args, "support", "trail") == "full" else ""),
            "fluid_solve": not args.coupling_only, "flow": flow, "gravity": list(gravity_of(args.workload)),
            "dt": dt, "nu": nu,
            "solver": ("pimpleFoamYade (UcEqn.H/pEqn.H, nOuterCorrectors 1, laminar)" if pimple else "icoFoamYade"),
            "fvSolution": "PISO nCorrectors 2; p PCG/DIC 1e-06 relTol 0.05 (pFinal 0); U smoothSolver symGaussSeidel 1e-05",
            "l2": "flushed between timed steps (256 MiB write)",
            "partition": partition_text(args.partition, world, ny, args.py)}


def flow_case(wl, pkg=None):
    """(oracle mesh, product mesh, U0, p0) of the workload's flow.  With a package (the engine arm) only the product's
    mesh is built -- nothing under oracle/ is imported there; without one (the CPU arm) only the oracle's."""
    from tests import cases_fv
    nx, ny, nz, P, seed, flow, dt, nu, _ = WORKLOADS[wl]
    if flow == "cavity":
        mo, mp = cases_fv.cavity3d(pkg, (nx, ny, nz), (1.0, 1.0, 1.0), oracle=pkg is None)
        U0 = np.zeros((nx * ny * nz, 3))
    elif flow == "closedbox":
        mo, mp = cases_fv.closed_box(pkg, (nx, ny, nz), (1.0, 1.0, 1.0), oracle=pkg is None)
        U0 = np.zeros((nx * ny * nz, 3))
    else:
        mo, mp = cases_fv.channel(pkg, (nx, ny, nz), (1.0, 1.0, 1.0), UIN, oracle=pkg is None)
        U0 = np.tile(np.array([UIN, 0.0, 0.0]), (nx * ny * nz, 1))
    return mo, mp, U0, np.zeros(nx * ny * nz)


def run_engine(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from tests import cases

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = g.load_package()
    wl = args.workload
    nx, ny, nz, P, seed, flow, dt, nu, desc = WORKLOADS[wl]
    gaussian = args.coupling == "gaussian"
    _, mp, U0, p0 = flow_case(wl, pkg)
    N = mp["nCells"]
    domain = args.partition == "domain" and world > 1
    if domain and (args.solver == "pimple" or args.coupling_only):
        raise SystemExit("bench.py: --partition domain decomposes icoFoamYade's pressure solve (use --partition particles / replicas)")
    sharded = (args.partition == "particles" or domain) and world > 1
    if sharded:
        # ONE domain, ONE particle buffer split over the ranks (strong scaling of the coupling half; the fluid
        # solve is replicated on every rank from identical inputs)
        lo, hi = pkg.sharded.shard_range(P, rank, world)
        pd = cases.particles(P, seed, radius=0.1 / nx, moving=True)[lo:hi].copy()
        P_total, P = P, hi - lo
    else:
        pd = cases.particles(P, pkg.replicas.particle_seed(seed, rank), radius=0.1 / nx, moving=True)
        P_total = P
    E = pkg.Engine(mp, device=local)
    E.set_properties(cases.RHOP, cases.RHOF, nu, gaussian)
    full_support = gaussian and args.support == "full"
    if full_support:
        E.set_gaussian_options(support_full=True)
    fluid = not args.coupling_only
    if fluid and not E.fv_supported():
        raise SystemExit("bench.py: " + E.L.fy_last_error(E.h).decode())
    E.set_piso_controls(nu=nu)
    E.upload("U", U0)
    E.upload("p", p0)
    E.create_phi()
    L = E.L
    dinfo = pkg.domain.init_domain(E, dist, "cuda", args.py) if domain else None

    # device-resident wire buffers (value) and pinned host wire buffers (e2e)
    d_pd = torch.from_numpy(pd).cuda()
    d_found = torch.empty(P, dtype=torch.int32, device="cuda")
    d_force = torch.empty(P, 6, dtype=torch.float64, device="cuda")
    h_pd = torch.from_numpy(pd).pin_memory()
    h_found = torch.empty(P, dtype=torch.int32).pin_memory()
    h_force = torch.empty(P, 6, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    pimple = args.solver == "pimple"
    if pimple and not gaussian:
        raise SystemExit("bench.py: pimpleFoamYade constructs the operator with gaussianInterp = true (pimpleFoamYade.C:53)")
    fluid_pre = E.pimple_pre if pimple else E.ico_pre
    if flow == "closedbox" and not pimple:
        raise SystemExit("bench.py: the closed-box workloads have fixedFluxPressure walls (pimpleFoamYade/pEqn.H:21): use --solver pimple")
    grav = gravity_of(wl)
    fluid_solve = (lambda dt_: E.pimple_solve(dt_, grav)) if pimple else E.ico_solve
    S = None
    if sharded:
        S = pkg.sharded.ShardedCoupling(E, dist, pkg.sharded.device_views(E), gaussian, pkg.sharded.external_stream_ctx(E))

    hz, bounds = 1.0 / nz, None
    if domain:
        bounds = torch.tensor([((r + 1) * nz) // world for r in range(world)], dtype=torch.int64, device="cuda")
    moved = {"records": 0}

    def coupling_domain(src):
        """particle migration to the owner slab's rank, coupling there, results back to the rank Yade handed them to"""
        k = torch.clamp(torch.floor(src[:, 2] / hz).to(torch.int64), 0, nz - 1)
        owner = torch.searchsorted(bounds, k, right=True)
        mine, route = pkg.domain.migrate(dist, src, owner, "cuda")
        n = mine.shape[0]
        fo = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
        Fo = torch.empty(max(n, 1), 6, dtype=torch.float64, device="cuda")
        S.step(dt, mine.data_ptr(), n, fo.data_ptr(), Fo.data_ptr())
        E.synchronize()
        d_force.copy_(pkg.domain.migrate_back(dist, Fo[:n], route, "cuda"))
        d_found.copy_(pkg.domain.migrate_back(dist, fo[:n].reshape(-1, 1), route, "cuda")[:, 0])
        moved["records"] = int((owner != rank).sum())

    def step_device():
        if fluid:
            fluid_pre(dt)
        if domain:
            coupling_domain(d_pd)
        elif S is not None:
            S.step(dt, d_pd.data_ptr(), P, d_found.data_ptr(), d_force.data_ptr())
        else:
            E.coupling_begin(dt)
            E.coupling_proc_device(d_pd.data_ptr(), P, d_found.data_ptr(), d_force.data_ptr())
        if fluid:
            fluid_solve(dt)
        E.set_source_zero()

    def step_e2e():
        if fluid:
            fluid_pre(dt)
        if domain:
            d_pd.copy_(h_pd, non_blocking=True)
            coupling_domain(d_pd)
            h_found.copy_(d_found, non_blocking=True)
            h_force.copy_(d_force, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        elif S is not None:
            with S.stream_ctx():
                d_pd.copy_(h_pd, non_blocking=True)
            S.step(dt, d_pd.data_ptr(), P, d_found.data_ptr(), d_force.data_ptr())
            with S.stream_ctx():
                h_found.copy_(d_found, non_blocking=True)
                h_force.copy_(d_force, non_blocking=True)
        else:
            L.fy_set_particle_action(E.h, dt, ctypes.c_void_p(h_pd.data_ptr()), P, ctypes.c_void_p(h_found.data_ptr()),
                                     ctypes.c_void_p(h_force.data_ptr()))
        if fluid:
            fluid_solve(dt)
        E.set_source_zero()
        E.synchronize()

    def step_e2e_overlapped():
        # the same host buffers, host<->device copies inside the timed region, on the engine's second stream: the
        # records come up under fy_ico_pre, the forces go down under fy_ico_solve (fycuda.h, "overlapped wire transfers")
        E._ck(L.fy_particles_upload_async(E.h, ctypes.c_void_p(h_pd.data_ptr()), P))
        if fluid:
            fluid_pre(dt)
        E.coupling_begin(dt)
        E._ck(L.fy_coupling_proc_staged(E.h, ctypes.c_void_p(h_found.data_ptr()), ctypes.c_void_p(h_force.data_ptr())))
        if fluid:
            fluid_solve(dt)
        E.set_source_zero()
        E._ck(L.fy_results_wait(E.h))
        E.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms = {}

    def timed(fn, steps):
        tot, its = 0.0, []
        per = step_ms.setdefault(fn.__name__, [])
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            E.timer_start()
            fn()
            per.append(E.timer_stop())
            tot += per[-1]
            if fluid:
                st = E.ico_stats()
                its.append(sum(q["iters"] for q in st["p"]))
        return tot, its

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    E.synchronize()
    barrier()
    l0 = E.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # both timed series integrate the SAME time steps: the flow state after the warm-up is restored in between
    state0 = dict(U=E.download("U"), p=E.download("p"), phi=E.download("phi")) if fluid else None
    ms, p_iters = timed(step_device, args.steps)
    barrier()
    launches = E.launch_count() - l0
    # e2e: same steps through the host-buffer call.  That path has one-time costs of its own (the library's staging
    # buffers are allocated on its first call, the pinned host buffers are touched by DMA for the first time): it gets
    # its own untimed warm-up before the state is restored and the series is timed.
    overlap = S is None and fluid and not args.e2e_blocking
    e2e_fn = step_e2e_overlapped if overlap else step_e2e
    for _ in range(warm):
        e2e_fn()
    E.synchronize()
    if fluid:
        E.upload("U", state0["U"]); E.upload("p", state0["p"]); E.upload("phi", state0["phi"])
        E.synchronize()
    ms_e2e, p_iters_e2e = timed(e2e_fn, args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel device times: separate profiling pass (extra events), not part of the headline
    E.set_profiling(True)
    E.kernel_ms(reset=True) if fluid else None
    phase = np.zeros(8)
    fl_ms = np.zeros(4)
    nprof = min(args.steps, 3)
    for _ in range(nprof):
        flush.fill_(1)
        torch.cuda.synchronize()
        if fluid:
            fluid_pre(dt)
        E.coupling_begin(dt)
        L.fy_coupling_proc(E.h, ctypes.c_void_p(h_pd.data_ptr()), P, ctypes.c_void_p(h_found.data_ptr()),
                           ctypes.c_void_p(h_force.data_ptr()))
        phase += E.phase_ms()
        if fluid:
            fluid_solve(dt)
            fl_ms += E.fluid_ms()
        E.set_source_zero()
    phase /= nprof
    fl_ms /= nprof
    kms = E.kernel_ms(reset=True) if fluid else None
    E.set_profiling(False)

    ms, ms_e2e = pkg.replicas.slowest_rank_ms([ms, ms_e2e], dist if world > 1 else None, "cuda")
    if rank == 0:
        peak, which = peaks()
        Fi = 3 * N - nx * ny - ny * nz - nx * nz
        # kernel classes with their algorithmic bytes per launch (DESIGN.md "kernels and rooflines")
        cand = {}
        if full_support:
            # (compulsory bytes as for the trail kernels; the ~370-cell gathers and REDs per particle are L2 work on top)
            cand["k_range_accumulate"] = (phase[1], 84.0 * P + 32.0 * N)
            cand["k_range_force"] = (phase[3], 48.0 * P + 136.0 * N)
        elif gaussian:
            cand["k_locate_gauss"] = (phase[1], 84.0 * P + 32.0 * N)
            cand["k_force_gauss"] = (phase[3], 48.0 * P + 136.0 * N)
        else:
            cand["k_point_force"] = (phase[3], 132.0 * P + 128.0 * N)
        if fluid and kms and kms["samples"] > 0:
            it_step = kms["pcg_iterations"] / float(nprof)
            # pencil-layout kernels: algorithmic bytes = 8 B x (streams read + written) per cell (kernels)
            Nl = N / float(world) if domain else N          # cells of this rank's part of the decomposed solve
            cand["k_pen2<Op2DicFwd> (DIC forward sweep)"] = (kms["precond_fwd"], 48.0 * Nl, it_step)       # {rD, rD low[3]} rA -> y
            cand["k_pen2<Op2DicBwd> (DIC backward sweep + wA.rA)"] = (kms["precond_bwd"], 56.0 * Nl, it_step)  # y {rD up[3]} rA -> z, re-arm y
            if kms["fused_tail"]:
                # z p -> p, re-arm z | dg up[3] p -> w | p w x r -> x r, in one cooperative launch
                cand["k_pen_tail<8> (direction + Amul + update)"] = (kms["direction"], 128.0 * Nl, it_step)
            else:
                cand["k_pen_amul_rows<8>"] = (kms["amul"], 48.0 * Nl, it_step)                             # dg up[3] p -> w (symmetric: lower = the neighbours' upper)
                cand["k_pen_update"] = (kms["update"], 48.0 * Nl, it_step)                                 # p w x r -> x r
                cand["k_pen_dir"] = (kms["direction"], 32.0 * Nl, it_step)                                 # z p -> p, re-arm z
        # dominant = largest share of the step
        def share(v):
            return v[0] * (v[2] if len(v) > 2 else 1.0)
        kname = max(cand, key=lambda k: share(cand[k]))
        kms_dom, kbytes = cand[kname][0], cand[kname][1]
        traffic, traffic_src = ncu_traffic(kname) if wl == "C2" else (None, None)
        ach = kbytes / (kms_dom * 1e-3) / 1e9 if kms_dom > 0 else 0.0
        line = {
            "metric": METRIC,
            "value": pkg.replicas.job_throughput(1 if sharded else world, args.steps, ms), "unit": "coupled timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_of(args, world),
            "e2e": {"value": pkg.replicas.job_throughput(1 if sharded else world, args.steps, ms_e2e), "unit": "coupled timesteps/s",
                    "h2d_bytes_per_step": 80 * P, "d2h_bytes_per_step": 52 * P,
                    "call": ("fy_particles_upload_async + fy_coupling_proc_staged + fy_results_wait (copies on the engine's second "
                             "stream, overlapped with fy_ico_pre / fy_ico_solve)" if overlap else
                             ("sharded / domain path: copies on the engine's stream" if S is not None else "fy_set_particle_action (blocking copies)"))},
            "gpu_launches": int(launches),
            "ms_steps": {k: [round(x, 3) for x in v] for k, v in step_ms.items()},
            "pcg_iterations_per_step": (float(np.mean(p_iters)) if p_iters else None),
            "pcg_iterations_per_step_e2e": (float(np.mean(p_iters_e2e)) if p_iters_e2e else None),
            "phase_ms": {"h2d": phase[0], "locate+weights+accumulate": phase[1], "void_fraction": phase[2],
                         "forces": phase[3], "d2h": phase[4], "UEqn+momentum_predictor": fl_ms[0],
                         "pressure_solves": fl_ms[1], "corrector_rest": fl_ms[2]},
            "kernel_ms": {k: {"ms": v[0], "alg_GBps": (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None),
                              "launches_per_step": (v[2] if len(v) > 2 else 1)} for k, v in cand.items()},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic, "peak_source": which, "kernel_ms": kms_dom,
                         "algorithmic_bytes": kbytes,
                         "traffic_source": ("%s (ncu --set full, dram read+write per launch)" % traffic_src) if traffic else None},
            "clocks": clocks,
            "particles_per_gpu": P,
        }
        if domain:
            di = E.dist_info()
            line["domain"] = {"grid_y_z": [dinfo["Py"], dinfo["Pz"]],
                              "iteration_collectives": ("inside the iteration's kernels over NVLink peer memory (CUDA IPC), CUDA-graph replayed"
                                                        if dinfo["peer"] else "NCCL calls between the kernels"), "rows_rank0": [dinfo["jLo"], dinfo["jHi"]],
                              "planes_rank0": [dinfo["kLo"], dinfo["kHi"]], "collectives_issued_rank0": di["collectives"],
                              "halo_bytes_sent_rank0": di["halo_bytes"], "particles_migrated_last_step_rank0": moved["records"],
                              "limiting_collective": "the three 1-double all-reduces + one halo exchange of every PCG iteration "
                                                     "(latency-bound: a few hundred KB per exchange at most)"}
        if not args.no_cpu_baseline and world == 1:
            state = dict(U=E.download("U"), p=E.download("p"), phi=E.download("phi")) if fluid else None
            if fluid and flow == "closedbox":
                # the fixedFluxPressure patches carry state from one time step to the next (the gradient constrainPressure
                # set, which fvc::grad(p) of the next pre-coupling block reads): it is part of the flow state
                state["bGradP"] = E.fv_get("bGradP")[int(mp["nInternalFaces"]):]
            ref_out = {}
            line["cpu_baseline"] = cpu_baseline(args, state=state, out=ref_out)
            if fluid and ref_out:
                # the same step (same start fields, same particle sample) on the engine: parity at the full mesh size
                Ps = ref_out["pd"].shape[0]
                E.upload("U", state["U"]); E.upload("p", state["p"]); E.upload("phi", state["phi"])
                fluid_pre(dt)
                fe, Fe = E.set_particle_action(dt, ref_out["pd"])
                fluid_solve(dt)
                E.set_source_zero()
                st = E.ico_stats()
                line["parity_full_size"] = {
                    "cells": N, "particles": int(Ps),
                    "U_rel_l2": cases.rel_l2(E.download("U"), ref_out["U"]), "p_rel_l2": cases.rel_l2(E.download("p"), ref_out["p"]),
                    "force_rel_l2": cases.rel_l2(Fe, ref_out["force"]), "found_equal": bool(np.array_equal(fe, ref_out["found"])),
                    "p_iters_engine": [q["iters"] for q in st["p"]], "p_iters_cpu": ref_out["p_iters"],
                    # (scale of the comparison: in a closed box under gravity the fluid is almost at rest -- U is the small
                    # difference HbyA - rAU grad p of terms of size |g| dt, and its relative error is amplified accordingly)
                    "U_rms": float(np.sqrt(np.mean(ref_out["U"] ** 2))), "g_dt": float(np.linalg.norm(grav) * dt),
                    "against": "oracle/_ref (unmodified FoamYade.C, canonical <=12 lists) + oracle/fv_oracle.cc, one step from the engine's state"}
        if (world == 1 and wl == "C2" and gaussian and fluid and not pimple and not full_support and not args.no_extra):
            line["extra_keys"] = extra_lines(args)
        print(json.dumps(line))
    torch.cuda.synchronize()
    S = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()         # before the engine (and the stream NCCL was ordered on) goes away
    E.close()


def extra_lines(args):
    """The other single-GPU configurations SURVEY.md 8(d) asks to see next to the headline, each measured by this same
    script in a process of its own (same timing rules; CPU arm and full-size parity skipped) and condensed: C2 with the
    point-force branch icoFoamYade hard-codes (icoFoamYade.C:53), and pimpleFoamYade on the closed box under gravity at
    C2's size (C3s; BASELINE configs[2] itself, 256^3 / 10 M, takes minutes to set up: profiles/).  A failure of an extra
    run is reported in its entry and never touches the headline line."""
    out = {}
    runs = {"C2_point_force": ["--coupling", "point"],
            "C3s_pimpleFoamYade_closed_box": ["--workload", "C3s", "--solver", "pimple"]}
    for name, extra in runs.items():
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--no-cpu-baseline", "--no-extra"] + extra
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
            ln = [x for x in r.stdout.splitlines() if x.startswith("{")]
            if r.returncode != 0 or not ln:
                out[name] = {"error": (r.stderr or r.stdout)[-300:]}
                continue
            d = json.loads(ln[-1])
            out[name] = {"value": d["value"], "unit": d["unit"], "e2e": d["e2e"]["value"], "ms_per_step": d["ms_per_step"],
                         "pcg_iterations_per_step": d["pcg_iterations_per_step"], "workload": d["config"]["workload"],
                         "coupling": d["config"]["coupling"], "solver": d["config"]["solver"],
                         "roofline_kernel": d["roofline"]["kernel"], "roofline_frac": d["roofline"]["frac"],
                         "command": "python bench.py " + " ".join(cmd[2:])}
        except Exception as e:                                   # noqa: BLE001
            out[name] = {"error": repr(e)[-300:]}
    return out


def cpu_baseline(args, state=None, steps=1, warmup=0, out=None):
    """The reference CPU path of one coupled step on the box's host cores (1 core: FoamYade.C and an OpenFOAM
    rank are single-threaded by construction): the reference's own coupling code (oracle/_ref, unmodified
    FoamYade.C) on a bounded particle sample + the oracle's restatement of the OpenFOAM fluid step on the
    full mesh.  `state` = fields to start from (the engine's current U, p, phi), else the workload's start."""
    from oracle import port, ref
    from tests import cases
    wl = args.workload
    nx, ny, nz, P, seed, flow, dt, nu, desc = WORKLOADS[wl]
    gaussian = args.coupling == "gaussian"
    Ps = min(P, args.cpu_particles)
    mo, _, U0, p0 = flow_case(wl, None)
    pd = cases.particles(P, seed, radius=0.1 / nx, moving=True)[:Ps]
    t0 = time.time()
    R = ref.RefFoamYade(mo, gaussian)
    t_tree = time.time() - t0
    R.set_properties(cases.RHOP, cases.RHOF, nu)
    if gaussian and getattr(args, "support", "trail") == "full":
        R.set_gaussian_options(True, False, False)
    fluid = not args.coupling_only
    O = None
    if fluid:
        O = port.IcoOracle(mo, nu=nu)
        O.field("U")[:] = state["U"] if state else U0
        O.field("p")[:] = state["p"] if state else p0
        if state:
            O.field("phi")[:] = state["phi"]
            if "bGradP" in state:
                O.field("bGradP")[:] = state["bGradP"]
        else:
            O.create_phi()
    t_cpl = t_fl = 0.0
    iters = []
    pimple = getattr(args, "solver", "ico") == "pimple"
    for it in range(warmup + steps):
        t0 = time.time()
        if fluid and pimple:
            ddtU, gradP, divT, vGrad = O.pimple_pre(dt, R.field("alpha").reshape(-1))
            for k, v in (("U", O.field("U")), ("ddtU", ddtU), ("gradP", gradP), ("divT", divT), ("vGrad", vGrad)):
                R.field(k)[:] = v.reshape(R.field(k).shape)
        elif fluid:
            O.pre(dt)
            R.field("U")[:] = O.field("U")
            R.field("vGrad")[:] = O.field("vGrad")
        t1 = time.time()
        found_cpu, force_cpu = R.step(dt, pd, pieces=True, truncate12=True, dense=True)
        if fluid:
            O.field("uSource")[:] = R.field("uSource")
        if pimple:
            alpha_c, drag_c = R.field("alpha").reshape(-1).copy(), R.field("uSourceDrag").reshape(-1).copy()
        R.set_source_zero()
        t2 = time.time()
        if fluid and pimple:
            O.pimple_solve(dt, alpha_c, drag_c, gravity_of(wl))
        elif fluid:
            O.solve(dt)
        t3 = time.time()
        if it >= warmup:
            t_cpl += t2 - t1
            t_fl += (t1 - t0) + (t3 - t2)
            if fluid:
                iters.append(sum(q["iters"] for q in O.stats()["p"]))
    if out is not None and fluid and steps == 1 and warmup == 0:
        out.update(pd=pd, U=O.field("U").copy(), p=O.field("p").copy(), found=found_cpu.copy(), force=force_cpu.copy(),
                   p_iters=[q["iters"] for q in O.stats()["p"]])
    R.close()
    if O:
        O.close()
    t_full = (t_cpl * (P / float(Ps)) + t_fl) / steps
    one = {"value": 1.0 / t_full, "coupling_seconds_scaled": t_cpl * (P / float(Ps)) / steps, "fluid_seconds": t_fl / steps}
    cores = host_cores()
    allc = cpu_all_cores(args, cores, state) if (cores > 1 and not args.cpu_one_core) else None
    fluid_txt = ("oracle port of the OpenFOAM-6 fluid step on the full %dx%dx%d mesh" % (nx, ny, nz)) if fluid else "no fluid solve"
    base = {"unit": "coupled timesteps/s", "kind": "port", "pcg_iterations_per_step": (float(np.mean(iters)) if iters else None),
            "single_core": {"value": one["value"], "coupling_seconds_scaled": one["coupling_seconds_scaled"], "fluid_seconds": one["fluid_seconds"],
                            "sample": "%d step(s): reference coupling operator (unmodified FoamYade.C, quadratic buildCellPartList replaced by its "
                                      "order-preserving dense accumulate) on %d of %d particles scaled linearly to P, %s; one-time k-d build "
                                      "%.1f s excluded" % (steps, Ps, P, fluid_txt, t_tree)}}
    if allc is None:
        base.update(value=one["value"], cores=1, sample=base["single_core"]["sample"],
                    coupling_seconds_scaled=one["coupling_seconds_scaled"], fluid_seconds=one["fluid_seconds"])
    else:
        base.update(allc)
    return base


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_all_cores(args, cores, state=None):
    """The same step with the host's cores used the way the reference uses them (`mpiexec -n <cores> icoFoamYade -parallel`,
    README.md:29): the particles are split over `cores` reference FoamYade objects (one process each, as the Foam ranks of a
    decomposed run each take the particles in their sub-domain), and the pressure solve runs decomposed into `cores` z slabs
    with one host thread per slab (oracle pcgSolvePar: slab-local DIC, OpenFOAM's decomposed preconditioner).  The FV assembly
    and the momentum predictor of the oracle stay on one core (they are 10-15 % of its step) -- stated in `sample`."""
    from oracle import port
    wl = args.workload
    nx, ny, nz, P, seed, flow, dt, nu, desc = WORKLOADS[wl]
    fluid = not args.coupling_only
    pimple = getattr(args, "solver", "ico") == "pimple"
    Ps = min(P, max(args.cpu_particles, 20000 * cores))
    # coupling: one process per core, each with its share of the particle sample
    procs = []
    for i in range(cores):
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", wl, "--coupling", args.coupling,
               "--support", getattr(args, "support", "trail"), "--cpl-worker", "%d,%d,%d" % (i, cores, Ps)]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=ROOT))
    t_fl, slabs = 0.0, min(cores, nz)
    if fluid and not pimple:
        mo, _, U0, p0 = flow_case(wl, None)
        O = port.IcoOracle(mo, nu=nu)
        O.field("U")[:] = state["U"] if state else U0
        O.field("p")[:] = state["p"] if state else p0
        if state:
            O.field("phi")[:] = state["phi"]
        else:
            O.create_phi()
        O.set_slabs(slabs)
        O.set_threads(slabs)
    t_cpl = 0.0
    for pr in procs:                                      # (the workers have the cores to themselves while they run)
        out, _ = pr.communicate(timeout=1800)
        for ln in out.splitlines():
            if ln.startswith("CPLWORKER "):
                t_cpl = max(t_cpl, float(ln.split()[1]))
    if fluid and not pimple:
        t0 = time.time()
        O.pre(dt)
        O.solve(dt)
        t_fl = time.time() - t0
        O.set_threads(1)
        O.close()
    elif fluid:
        return None                                       # (pimpleFoamYade: single-core baseline only)
    t_cpl_scaled = t_cpl * (P / float(Ps))
    return {"value": 1.0 / (t_cpl_scaled + t_fl), "cores": cores, "coupling_seconds_scaled": t_cpl_scaled, "fluid_seconds": t_fl,
            "sample": "1 step on %d host cores: %d reference FoamYade objects (unmodified FoamYade.C, dense accumulate), one process per "
                      "core, %d of %d particles split between them and scaled linearly to P (slowest process counts); oracle fluid step "
                      "with the pressure solve decomposed into %d z slabs, one thread per slab (slab-local DIC as in OpenFOAM -parallel); "
                      "the oracle's FV assembly and momentum predictor run on one core" % (cores, cores, Ps, P, slabs)}


def cpl_worker(args):
    """one of the all-core baseline's coupling processes: `i,n,Ps` = worker i of n on its share of the first Ps particles"""
    from oracle import ref
    from tests import cases
    i, n, Ps = (int(x) for x in args.cpl_worker.split(","))
    nx, ny, nz, P, seed, flow, dt, nu, desc = WORKLOADS[args.workload]
    gaussian = args.coupling == "gaussian"
    mo, _, U0, p0 = flow_case(args.workload, None)
    pd = cases.particles(P, seed, radius=0.1 / nx, moving=True)[:Ps]
    lo, hi = (i * Ps) // n, ((i + 1) * Ps) // n
    R = ref.RefFoamYade(mo, gaussian)
    R.set_properties(cases.RHOP, cases.RHOF, nu)
    if gaussian and getattr(args, "support", "trail") == "full":
        R.set_gaussian_options(True, False, False)
    R.field("U")[:] = U0
    t0 = time.time()
    R.step(dt, pd[lo:hi], pieces=True, truncate12=True, dense=True)
    R.set_source_zero()
    print("CPLWORKER %.6f" % (time.time() - t0), flush=True)
    R.close()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.cpl_worker:
        return cpl_worker(args)
    # bounded: at 128^3 one CPU step is ~10-20 s
    steps = max(1, min(args.steps, 2))
    warm = max(0, min(args.warmup, 1))
    cb = cpu_baseline(args, state=None, steps=steps, warmup=warm)
    v = cb["value"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nx, ny, nz, P, seed, flow, dt, nu, desc = WORKLOADS[args.workload]
    print(json.dumps({
        "impl": "reference", "metric": METRIC,
        "value": v, "unit": "coupled timesteps/s", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 / v, "higher_is_better": True,
        "scaling": "strong" if (world > 1 and args.partition != "replicas") else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args, world),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "coupled timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--coupling", default="gaussian", choices=["gaussian", "point"])
    ap.add_argument("--solver", default="ico", choices=["ico", "pimple"],
                    help="fluid step: icoFoamYade (default, the headline) or pimpleFoamYade (UcEqn.H/pEqn.H; Gaussian coupling)")
    ap.add_argument("--coupling-only", action="store_true")
    ap.add_argument("--support", default="trail", choices=["trail", "full"],
                    help="Gaussian cell sets: the reference's <= 12-cell k-d trail (default) or every cell inside the search bound "
                         "(fy_set_gaussian_options FY_SUPPORT_FULL, ~370 cells per particle)")
    ap.add_argument("--partition", default="domain", choices=["domain", "particles", "replicas"],
                    help="N > 1: one domain, pressure solve z-slab decomposed + particle migration (strong scaling, default); "
                         "one domain with only the particle buffer sharded; or independent domain replicas (weak scaling)")
    ap.add_argument("--py", type=int, default=0,
                    help="--partition domain: y slabs of the rank grid (0: the library's choice, y first; 1: z slabs only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra single-GPU configurations reported under extra_keys")
    ap.add_argument("--e2e-blocking", action="store_true", help="e2e through the blocking fy_set_particle_action instead of the overlapped calls")
    ap.add_argument("--cpu-particles", type=int, default=200000)
    ap.add_argument("--cpu-one-core", action="store_true", help="CPU baseline on one core only (skip the all-core arm)")
    ap.add_argument("--cpl-worker", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
