"""CPU, world_size 2, gloo: the host-side plumbing of the domain-decomposed multi-GPU step
(yade-openfoam-coupling_b200/domain.py) -- slab arithmetic, the NCCL-id hand-over, particle migration by owner
slab and the way back.  The decomposed solve itself needs GPUs (tests/test_gpu_domain.py)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_ranges_partition_the_planes(pkg):
    d = pkg.domain
    for nz in (1, 5, 12, 128, 131):
        for world in (1, 2, 3, 8):
            if world > nz:
                continue
            r = [d.slab_range(nz, q, world) for q in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nz
            assert all(r[q][1] == r[q + 1][0] for q in range(world - 1))
            assert all(hi > lo for lo, hi in r)
            z = (np.arange(nz) + 0.5) * 0.25 + 1.0
            own = d.owner_slab(z, 1.0, 0.25, nz, world)
            for q, (lo, hi) in enumerate(r):
                assert np.all(own[lo:hi] == q)
    # particles outside the box go to the nearest slab (they are reported as not found there)
    assert list(d.owner_slab([-3.0, 99.0], 0.0, 0.1, 10, 2)) == [0, 1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = pkg.domain
    uid = bytes(range(128)) if rank == 0 else b"\0" * 128
    got = d.broadcast_id(dist, uid)
    rng = np.random.default_rng(100 + rank)
    n = 50 + 13 * rank
    pd = torch.from_numpy(rng.uniform(0, 1, (n, 10)))
    pd[:, 9] = torch.arange(n) + 1000 * rank               # a tag to follow every record
    owner = torch.from_numpy(d.owner_slab(pd[:, 2].numpy(), 0.0, 0.1, 10, world).astype(np.int64))
    mine, route = d.migrate(dist, pd, owner, "cpu")
    lo, hi = d.slab_range(10, rank, world)
    ok_owner = bool(np.all((np.floor(mine[:, 2].numpy() / 0.1) >= lo) & (np.floor(mine[:, 2].numpy() / 0.1) < hi)))
    back = d.migrate_back(dist, mine[:, 9:10] * 2.0, route, "cpu")        # a per-particle result: 2 x tag
    ok_back = bool(torch.equal(back[:, 0], pd[:, 9] * 2.0))
    tot = torch.tensor([float(mine.shape[0])])
    dist.all_reduce(tot)
    dist.destroy_process_group()
    q.put((rank, got == bytes(range(128)), ok_owner, ok_back, int(tot.item())))


def test_id_handover_and_particle_migration_gloo(pkg):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, id_ok, ok_owner, ok_back, tot in out:
        assert id_ok and ok_owner and ok_back
        assert tot == 50 + 63                              # no record lost or duplicated


def test_solve_grid_prefers_y_cuts(pkg):
    d = pkg.domain
    assert d.solve_grid(128, 1) == (1, 1)
    assert d.solve_grid(128, 2) == (2, 1)
    assert d.solve_grid(128, 4) == (4, 1)
    assert d.solve_grid(128, 8) == (4, 2)             # four 32-row blocks: the rest of the ranks cut z
    assert d.solve_grid(14, 2) == (1, 2)              # one block: z slabs
    assert d.solve_grid(70, 4) == (2, 2)              # three blocks: the largest divisor of 4 that fits
    assert d.solve_grid(128, 8, py=1) == (1, 8)


def test_oracle_grid_partition_semantics(pkg):
    """The oracle's decomposed DIC (the checker of tests/test_gpu_domain.py): a 1 x Pz grid IS the z-slab partition; a y
    cut changes the preconditioner (other iteration count) but not the converged solution."""
    from oracle import port
    from tests import cases_fv
    mo, _ = cases_fv.channel(pkg, (10, 70, 6))
    rng = np.random.default_rng(4)
    N, Fi = mo["nCells"], mo["nInternalFaces"]
    upper = rng.uniform(0.5, 1.5, Fi)
    diag = np.zeros(N)
    np.subtract.at(diag, mo["owner"], upper)
    np.subtract.at(diag, mo["neighbour"], upper)
    diag -= rng.uniform(0.001, 0.01, N)
    b = rng.standard_normal(N)
    O = port.IcoOracle(mo, nu=1e-3)
    run = lambda: O.pcg(diag, upper, b, np.zeros(N), tol=1e-10, relTol=0.0, preconditioner="DIC")
    x1, p1 = run()
    O.set_slabs(2)
    xs, ps = run()
    O.set_grid(1, 2)
    xg, pg = run()
    assert np.array_equal(xs, xg) and ps["iters"] == pg["iters"]
    O.set_grid(2, 1)
    xy, py_ = run()
    O.set_grid(3, 2)
    xyz, pyz = run()
    O.set_grid(1, 1)
    x0, p0 = run()
    assert np.array_equal(x0, x1) and p0["iters"] == p1["iters"]
    assert py_["iters"] >= p1["iters"] and pyz["iters"] >= py_["iters"]
    assert not np.array_equal(xy, x1)
    for x in (xs, xy, xyz):
        assert np.linalg.norm(x - x1) <= 1e-7 * np.linalg.norm(x1)
    O.close()
