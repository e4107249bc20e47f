#!/usr/bin/env python
"""bench.py -- coupled timesteps/s of the FoamYade hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--coupling gaussian|point]
  python bench.py --impl reference ...      # the reference's own CPU path (oracle/_ref + oracle port)

One "step" = one pass of the hot path on one batch of synthetic particles:
  vGrad = grad(U)  ->  setParticleAction(dt)  ->  UEqn + PISO correctors (PCG)  ->  setSourceZero()
(icoFoamYade.C:65-149).  `value` times it with the particle records already resident in HBM
(fy_coupling_proc_device); `e2e` times the same step through the reference-facing call
fy_set_particle_action with pinned HOST wire buffers (80 B/particle in, 52 B/particle out).
Timing: CUDA events on the engine's own stream, max over ranks; L2 is flushed between timed steps.
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

WORKLOADS = {
    # name: (nx, ny, nz, particles, seed, description)
    "C1": (32, 32, 32, 1000, 42, "icoFoamYade lid-driven cavity 32^3 cells, 1k particles"),
    "C2": (128, 128, 128, 1000000, 7, "icoFoamYade channel 128^3 cells, 1M particles, fp64, 1xB200"),
    "C3": (256, 256, 256, 10000000, 1001, "256^3 cells, 10M particles"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


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


def build_case(pkg, wl, coupling, rank=0):
    from tests import cases
    nx, ny, nz, P, seed, _ = WORKLOADS[wl]
    mesh = pkg.box_mesh(nx, ny, nz, faces=True)
    flds = cases.fields_for(mesh["C"])
    pd = cases.particles(P, seed + 1000 * rank, radius=0.1 / nx, moving=True)
    return mesh, flds, pd


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
    gaussian = args.coupling == "gaussian"
    mesh, flds, pd = build_case(pkg, wl, args.coupling, rank)
    N, P = mesh["nCells"], pd.shape[0]
    E = pkg.Engine(mesh, device=local)
    E.set_properties(cases.RHOP, cases.RHOF, cases.NU, gaussian)
    for k in ("U", "gradP", "divT", "vGrad"):
        E.upload(k, flds[k])
    L = E.L

    # device-resident wire buffers (value) and pinned host wire buffers (e2e)
    d_pd = torch.from_numpy(pd).cuda()
    d_found = torch.empty(P, dtype=torch.int32, device="cuda")
    d_force = torch.empty(P, 6, dtype=torch.float64, device="cuda")
    h_pd = torch.from_numpy(pd).pin_memory()
    h_found = torch.empty(P, dtype=torch.int32).pin_memory()
    h_force = torch.empty(P, 6, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    fluid = hasattr(E, "fluid_step") and not args.coupling_only
    dt = 1e-3

    def step_device():
        E.coupling_begin(dt)
        E.coupling_proc_device(d_pd.data_ptr(), P, d_found.data_ptr(), d_force.data_ptr())
        if fluid:
            E.fluid_step(dt)
        E.set_source_zero()

    def step_e2e():
        L.fy_set_particle_action(E.h, dt, ctypes.c_void_p(h_pd.data_ptr()), P, ctypes.c_void_p(h_found.data_ptr()),
                                 ctypes.c_void_p(h_force.data_ptr()))
        if fluid:
            E.fluid_step(dt)
        E.set_source_zero()
        E.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # engine-stream event timing through the ABI
    def timed(fn, steps):
        tot = 0.0
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            E.timer_start()
            fn()
            tot += E.timer_stop()
        return tot

    for _ in range(max(args.warmup, 3)):
        step_device()
    E.synchronize()
    barrier()
    l0 = E.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, args.steps)
    barrier()
    launches = E.launch_count() - l0
    # per-kernel times (events around each phase, separate pass so that the headline has no extra events)
    E.set_profiling(True)
    phase = np.zeros(8)
    for _ in range(min(args.steps, 5)):
        flush.fill_(1)
        torch.cuda.synchronize()
        E.coupling_begin(dt)
        L.fy_coupling_proc(E.h, ctypes.c_void_p(h_pd.data_ptr()), P, ctypes.c_void_p(h_found.data_ptr()),
                           ctypes.c_void_p(h_force.data_ptr()))
        phase += E.phase_ms()
        E.set_source_zero()
    phase /= min(args.steps, 5)
    E.set_profiling(False)
    for _ in range(2):
        step_e2e()
    barrier()
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        peak, which = peaks()
        # dominant kernel of the coupling operator and its algorithmic bytes (DESIGN.md "roofline")
        if gaussian:
            kname, kms = "k_locate_gauss", phase[1]
            kbytes = 80.0 * P + 4 * P + 32.0 * N            # particle record in, found out, tree nodes once
        else:
            kname, kms = "k_point_force", phase[3]
            kbytes = 132.0 * P + (24 + 72 + 8 + 24) * float(N)
        ach = kbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
        line = {
            "metric": "coupled timesteps/sec @1M particles/128^3 cells; achieved HBM GB/s vs peak",
            "value": world * args.steps / (ms * 1e-3), "unit": "coupled timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (wl, WORKLOADS[wl][5]), "cells": N, "particles_per_gpu": P,
                       "coupling": args.coupling, "fluid_solve": bool(fluid), "l2": "flushed between timed steps (256 MiB write)",
                       "partition": "replicas" if world > 1 else "single domain"},
            "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "coupled timesteps/s",
                    "h2d_bytes_per_step": 80 * P, "d2h_bytes_per_step": 52 * P},
            "gpu_launches": int(launches),
            "phase_ms": {"h2d": phase[0], "locate+weights+accumulate": phase[1], "void_fraction": phase[2],
                         "forces": phase[3], "d2h": phase[4]},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": None, "peak_source": which, "kernel_ms": kms},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, sample_only=True)
        print(json.dumps(line))
    E.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, sample_only):
    """The reference's own coupling code (oracle/_ref, unmodified FoamYade.C; 1 core: it is single-threaded by
    construction) on a bounded sample of the workload's particles, full mesh."""
    from oracle import meshgen, ref
    from tests import cases
    nx, ny, nz, P, seed, desc = WORKLOADS[args.workload]
    gaussian = args.coupling == "gaussian"
    Ps = min(P, args.cpu_particles)
    mo = meshgen.hex_box(nx, ny, nz)
    flds = cases.fields_for(mo["C"])
    pd = cases.particles(P, seed, radius=0.1 / nx, moving=True)[:Ps]
    t0 = time.time()
    R = ref.RefFoamYade(mo, gaussian)
    t_tree = time.time() - t0
    R.set_properties(cases.RHOP, cases.RHOF, cases.NU)
    for k in ("U", "gradP", "divT", "vGrad"):
        R.field(k)[:] = flds[k]
    t0 = time.time()
    R.step(1e-3, pd, pieces=True, truncate12=True, dense=True)
    R.set_source_zero()
    t_step = time.time() - t0
    R.close()
    t_full = t_step * (P / float(Ps))
    return {"value": 1.0 / t_full, "unit": "coupled timesteps/s", "cores": 1, "kind": "reference",
            "sample": "%d of %d particles on the full %dx%dx%d mesh, coupling operator only, quadratic "
                      "buildCellPartList replaced by its order-preserving dense accumulate; scaled linearly to P; "
                      "one-time k-d build %.1f s excluded" % (Ps, P, nx, ny, nz, t_tree),
            "sample_seconds": t_step}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    vals = []
    for _ in range(steps):
        vals.append(cpu_baseline(args, sample_only=True))
    v = float(np.median([b["value"] for b in vals]))
    cb = dict(vals[0])
    cb["value"] = v
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nx, ny, nz, P, seed, desc = WORKLOADS[args.workload]
    print(json.dumps({
        "impl": "reference", "metric": "coupled timesteps/sec @1M particles/128^3 cells; achieved HBM GB/s vs peak",
        "value": v, "unit": "coupled timesteps/s", "n_gpus": world, "steps": steps, "warmup": 0,
        "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, desc), "cells": nx * ny * nz, "particles_per_gpu": P,
                   "coupling": args.coupling},
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
    ap.add_argument("--coupling-only", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-particles", type=int, default=200000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
