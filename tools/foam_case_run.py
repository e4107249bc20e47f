#!/usr/bin/env python
"""Run an OpenFOAM case directory on the engine without OpenFOAM.

  python tools/foam_case_run.py <case> [--solver icoFoamYade|pimpleFoamYade] [--steps N] [--particles file.npy]
                                       [--gaussian] [--rhoP 2500 --rhoF 1000] [--no-write]

Reads constant/polyMesh, 0/U, 0/p, transportProperties, controlDict and fvSolution (yade-openfoam-coupling_b200/foamcase.py),
runs the solver's time loop (icoFoamYade.C:65-149 / pimpleFoamYade.C:65-110) with every field resident on the GPU, prints
the solver log in OpenFOAM's own format (so that it diffs against a log.icoFoam) and writes the time directories the
controlDict asks for.  --particles: a [P][10] wire-record array (x y z vx vy vz wx wy wz radius) held fixed, standing in
for the Yade side; without it the coupling call runs with zero particles (the reference solver without a Yade peer
would not start at all)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def fmt(x):
    return "%.6g" % x


def log_step(t, st, valid=(True, True, True), uname="U", n_non_orth=0, precond="DIC", out=None):
    """the lines icoFoam / pimpleFoam print for one time step (components in an empty direction are not solved)"""
    w = (out or sys.stdout).write
    w("Time = %s\n\n" % fmt(t))
    w("Courant Number mean: %s max: %s\n" % (fmt(st["meanCoNum"]), fmt(st["CoNum"])))
    for j, ok in enumerate(valid):
        if ok:
            u = st["U"][j]
            w("smoothSolver:  Solving for %s%s, Initial residual = %s, Final residual = %s, No Iterations %d\n"
              % (uname, "xyz"[j], fmt(u["initial"]), fmt(u["final"]), u["iters"]))
    pname = {"DIC": "DICPCG", "diagonal": "diagonalPCG", "none": "PCG"}[precond]
    for q, p in enumerate(st["p"]):
        w("%s:  Solving for p, Initial residual = %s, Final residual = %s, No Iterations %d\n"
          % (pname, fmt(p["initial"]), fmt(p["final"]), p["iters"]))
        if (q + 1) % (n_non_orth + 1) == 0:                  # continuityErrs.H follows the non-orthogonal corrector loop
            c = (q + 1) // (n_non_orth + 1) - 1
            w("time step continuity errors : sum local = %s, global = %s\n" % (fmt(st["corrSumLocal"][c]), fmt(st["corrGlobal"][c])))
    w("\n")


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--solver", default="icoFoamYade", choices=["icoFoamYade", "pimpleFoamYade"])
    ap.add_argument("--steps", type=int, default=0, help="time steps to run (0: until controlDict's endTime)")
    ap.add_argument("--particles", default="")
    ap.add_argument("--gaussian", action="store_true", help="Gaussian coupling (pimpleFoamYade always; icoFoamYade hard-codes point force)")
    ap.add_argument("--rhoP", type=float, default=2500.0)
    ap.add_argument("--rhoF", type=float, default=1000.0)
    ap.add_argument("--no-write", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    import __graft_entry__ as g
    pkg = g.load_package()
    fc = pkg.foamcase
    case = fc.load_case(args.case, solver=args.solver)
    mesh = fc.build_mesh(case, pkg.box_mesh, pkg.set_bc, pkg)
    pimple = args.solver == "pimpleFoamYade"
    gaussian = pimple or args.gaussian
    E = pkg.Engine(mesh, device=args.device)
    if not E.fv_supported():
        raise SystemExit(E.L.fy_last_error(E.h).decode())
    rhoP = case["props"].get("partDensity", args.rhoP)
    rhoF = case["props"].get("rhocValue", case["props"].get("fluidDensity", args.rhoF))
    E.set_properties(rhoP, rhoF, case["nu"], gaussian)
    E.set_piso_controls(nu=case["nu"], **case["piso"])
    if pimple:
        E.set_pimple_controls(**case["pimple"])
    E.upload("U", case["U"])
    E.upload("p", case["p"])
    E.create_phi()
    pd = np.load(args.particles) if args.particles else np.zeros((0, 10))
    ctl = case["control"]
    dt, t = ctl["deltaT"], ctl["startTime"]
    nsteps = args.steps or int(round((ctl["endTime"] - t) / dt))
    owners = None if args.no_write else fc.boundary_owner_cells(case)
    empty = set(sd for pt in case["patches"] if pt["bcU"] == "empty" for sd in pt["sides"])
    valid = tuple(not ({a + "min", a + "max"} <= empty) for a in "xyz")
    for it in range(1, nsteps + 1):
        t += dt
        if pimple:
            E.pimple_pre(dt)
        else:
            E.ico_pre(dt)
        E.set_particle_action(dt, pd)
        if pimple:
            E.pimple_solve(dt, case["g"])
        else:
            E.ico_solve(dt)
        E.set_source_zero()
        log_step(t, E.ico_stats(), valid, case["Uname"], case["piso"]["nNonOrthogonalCorrectors"], case["piso"]["preconditioner"])
        if not args.no_write and ctl["writeControl"] == "timeStep" and it % max(1, int(round(ctl["writeInterval"]))) == 0:
            fc.write_time(case, t, E.download("U"), E.download("p"), owners=owners)
    if not args.no_write and nsteps % max(1, int(round(ctl["writeInterval"]))) != 0 and args.steps:
        fc.write_time(case, t, E.download("U"), E.download("p"), owners=owners)      # a shortened run still leaves its last state
    print("End\n")
    E.close()


if __name__ == "__main__":
    main()
