"""GPU x2 (NCCL): the decomposed (z slabs, and y slabs of 32-row blocks) pressure solve and the domain-decomposed coupled icoFoamYade step
(csrc/fv_dist.cu, domain.py) against the oracle run with the SAME partition (OpenFOAM's decomposed semantics: the DIC
preconditioner factorises each processor's own matrix) -- identical iteration counts, fields within 1e-10 -- and against
the single-domain run at solver tolerance.  Needs two devices; the round-end single-GPU run skips it
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_domain.py -m gpu`)."""
import os
import socket
import sys

import numpy as np
import pytest

from tests import cases, cases_fv

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOXES = {"z": (20, 14, 12), "y": (24, 70, 12)}      # ny = 70: three 32-row blocks, split 1 + 2 over two ranks
P = 6000


def _matrix(mo):
    rng = np.random.default_rng(4)
    N, Fi = mo["nCells"], mo["nInternalFaces"]
    upper = rng.uniform(0.5, 1.5, Fi)
    diag = np.zeros(N)
    np.subtract.at(diag, mo["owner"], upper)
    np.subtract.at(diag, mo["neighbour"], upper)
    diag -= rng.uniform(0.001, 0.01, N)
    return diag, upper, rng.standard_normal(N)


def _worker(rank, world, port, q, NBOX, peer):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    torch.cuda.set_device(rank)
    if not peer:
        os.environ["FY_DIST_PEER"] = "0"              # the iteration's collectives as NCCL calls (the fallback path)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    _, mp_ = cases_fv.channel(pkg, NBOX, oracle=False)
    E = pkg.Engine(mp_, device=rank)
    nu = 1e-3
    E.set_properties(cases.RHOP, cases.RHOF, nu, True)
    info = pkg.domain.init_domain(E, dist, "cuda")
    out = dict(info=info)
    # (1) the linear solver alone
    diag, upper, b = _matrix(mp_)
    x, perf = E.pcg(diag, upper, b, np.zeros(mp_["nCells"]), tol=1e-10, relTol=0.0, preconditioner="DIC")
    out["pcg"] = (x, perf)
    x2, perf2 = E.pcg(diag, upper, b, np.zeros(mp_["nCells"]), tol=1e-8, relTol=0.0, preconditioner="diagonal")
    out["pcg_diag"] = (x2, perf2)
    # (2) coupled steps: particles migrate to the rank owning their slab, per-cell sums are all-reduced, the pressure
    # solves run decomposed
    U0, p0 = cases_fv.channel_init(mp_["C"])
    E.set_piso_controls(nu=nu)
    E.upload("U", U0)
    E.upload("p", p0)
    E.create_phi()
    sh = pkg.sharded
    pd_all = cases.particles(P, 31, radius=0.1 / NBOX[0], moving=True)
    pd_all[:, 0] *= 2.0                                   # the channel is 2 x 1 x 1
    lo, hi = sh.shard_range(P, rank, world)               # what "Yade" hands this rank
    d_in = torch.from_numpy(pd_all[lo:hi].copy()).cuda()
    S = sh.ShardedCoupling(E, dist, sh.device_views(E), True, sh.external_stream_ctx(E))
    steps = []
    for it in range(3):
        dt = 2e-3
        E.ico_pre(dt)
        owner = torch.from_numpy(pkg.domain.owner_slab(d_in[:, 2].cpu().numpy(), 0.0, 1.0 / NBOX[2], NBOX[2], world).astype(np.int64)).cuda()
        mine, route = pkg.domain.migrate(dist, d_in, owner, "cuda")
        n = mine.shape[0]
        d_found = torch.zeros(max(n, 1), dtype=torch.int32, device="cuda")
        d_force = torch.zeros(max(n, 1), 6, dtype=torch.float64, device="cuda")
        S.step(dt, mine.data_ptr(), n, d_found.data_ptr(), d_force.data_ptr())
        E.synchronize()
        force = pkg.domain.migrate_back(dist, d_force[:n], route, "cuda")
        found = pkg.domain.migrate_back(dist, d_found[:n].reshape(-1, 1), route, "cuda")[:, 0]
        E.ico_solve(dt)
        E.set_source_zero()
        steps.append(dict(U=E.download("U"), p=E.download("p"), force=force.cpu().numpy(), found=found.cpu().numpy(),
                          iters=[qq["iters"] for qq in E.ico_stats()["p"]], uiters=[qq["iters"] for qq in E.ico_stats()["U"]],
                          owned=n))
    out["steps"] = steps
    out["info_end"] = E.dist_info()
    dist.barrier()
    dist.destroy_process_group()
    E.close()
    q.put((rank, out))


@pytest.mark.parametrize("cut,peer", [("z", True), ("y", True), ("y", False)])
def test_two_gpu_domain_decomposed_solve_and_step(pkg, cut, peer):
    import torch
    import torch.multiprocessing as mp
    from oracle import port, ref
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    NBOX = BOXES[cut]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port_no, q, NBOX, peer)) for r in range(2)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=600) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    if cut == "z":
        assert [(out[r]["info"]["Py"], out[r]["info"]["kLo"], out[r]["info"]["kHi"]) for r in range(2)] == [(1, 0, 6), (1, 6, 12)]
    else:
        assert [(out[r]["info"]["Py"], out[r]["info"]["jLo"], out[r]["info"]["jHi"]) for r in range(2)] == [(2, 0, 32), (2, 32, 70)]
        assert all(out[r]["info"]["kLo"] == 0 and out[r]["info"]["kHi"] == NBOX[2] for r in range(2))
    grid = (1, 2) if cut == "z" else (2, 1)
    # the collectives of the PCG iteration ran inside its kernels over peer memory (or, asked so, as NCCL calls)
    assert all(out[r]["info"]["peer"] == peer for r in range(2)), out[0]["info"]
    assert out[0]["info_end"]["collectives"] > 0 and (out[0]["info_end"]["halo_bytes"] > 0) == (not peer)   # (NCCL calls only)

    mo, mp_ = cases_fv.channel(pkg, NBOX)
    nu = 1e-3
    # (1) PCG: decomposed == the oracle with the same partition; both ranks hold the same solution
    diag, upper, b = _matrix(mo)
    O = port.IcoOracle(mo, nu=nu)
    O.set_grid(*grid)
    xo, po = O.pcg(diag, upper, b, np.zeros(mo["nCells"]), tol=1e-10, relTol=0.0, preconditioner="DIC")
    O.set_grid(1, 1)
    x1, p1 = O.pcg(diag, upper, b, np.zeros(mo["nCells"]), tol=1e-10, relTol=0.0, preconditioner="DIC")
    assert po["iters"] != p1["iters"] or not np.array_equal(xo, x1)      # the partition really changes the preconditioner
    for r in range(2):
        xe, pe = out[r]["pcg"]
        assert pe["iters"] == po["iters"]
        assert cases.rel_l2(xe, xo) <= cases.TOL
        assert cases.rel_l2(xe, x1) <= 1e-7                              # same converged solution as the single domain
    assert np.array_equal(out[0]["pcg"][0], out[1]["pcg"][0])
    xd, pd_ = O.pcg(diag, upper, b, np.zeros(mo["nCells"]), tol=1e-8, relTol=0.0, preconditioner="diagonal")
    for r in range(2):
        assert out[r]["pcg_diag"][1]["iters"] == pd_["iters"] and cases.rel_l2(out[r]["pcg_diag"][0], xd) <= cases.TOL

    # (2) coupled steps against (unmodified reference coupling + oracle fluid step with the same partition)
    O.set_grid(*grid)
    U0, p0 = cases_fv.channel_init(mo["C"])
    O.field("U")[:] = U0
    O.field("p")[:] = p0
    O.create_phi()
    R = ref.RefFoamYade(mo, True)
    R.set_properties(cases.RHOP, cases.RHOF, nu)
    pd_all = cases.particles(P, 31, radius=0.1 / NBOX[0], moving=True)
    pd_all[:, 0] *= 2.0
    sh = pkg.sharded
    for it in range(3):
        dt = 2e-3
        O.pre(dt)
        R.field("U")[:] = O.field("U")
        R.field("vGrad")[:] = O.field("vGrad")
        fo, Fo = R.step(dt, pd_all, pieces=True)
        O.field("uSource")[:] = R.field("uSource")
        O.solve(dt)
        R.set_source_zero()
        so = O.stats()
        found = np.concatenate([out[r]["steps"][it]["found"] for r in range(2)])
        force = np.concatenate([out[r]["steps"][it]["force"] for r in range(2)])
        assert np.array_equal(found, fo)
        assert cases.rel_l2(force, Fo) <= cases.TOL
        assert sum(out[r]["steps"][it]["owned"] for r in range(2)) == P
        for r in range(2):
            st = out[r]["steps"][it]
            assert st["iters"] == [qq["iters"] for qq in so["p"]], (it, r, st["iters"], so["p"])
            # (the momentum components are solved one per rank and travel with their solver statistics)
            assert st["uiters"] == [qq["iters"] for qq in so["U"]], (it, r, st["uiters"], so["U"])
            assert cases.rel_l2(st["U"], O.field("U")) <= cases.TOL
            assert cases.rel_l2(st["p"], O.field("p")) <= cases.TOL
        assert np.array_equal(out[0]["steps"][it]["U"], out[1]["steps"][it]["U"])
    R.close()
    O.close()
