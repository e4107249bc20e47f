"""CPU, world_size 2 over gloo: the particle-sharded coupling plumbing (yade-openfoam-coupling_b200/sharded.py) around
a numpy stand-in for the three device passes: two ranks with half the buffer each + the all-reduces between the
passes give what one rank computes from the whole buffer."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CELLS, P = 64, 400


class FakeEngine:
    """Same call surface as Engine for the passes; arithmetic: every particle adds to 3 pseudo-random cells."""

    def __init__(self, views):
        self.v = views
        self.serial = 0

    def coupling_begin(self, dt):
        self.dt = dt

    def coupling_pass_device(self, p, pd, n, found, force):
        v = self.v
        if p == 0:
            self.serial += 1
            for q in range(n):
                for s in range(3):
                    c = int(pd[q, 0] * 7919 + s * 31) % N_CELLS
                    v["pvol"][c] += pd[q, 1]
                    v["upAcc"][c] += pd[q, 1] * pd[q, 2:5]
                    v["stamp"][c] = self.serial
                found[q] = 1
        elif p == 1:
            m = v["stamp"] == self.serial
            self.alpha = torch.ones(N_CELLS, dtype=torch.float64)
            self.alpha[m] = 1.0 - v["pvol"][m]
            v["pvol"][m] = 0
            v["upAcc"][m] = 0
        else:
            for q in range(n):
                c = int(pd[q, 0] * 7919) % N_CELLS
                f = self.alpha[c] * pd[q, 2:5]
                force[q, :3] = f
                v["uSource"][c] -= f
                v["uSourceDrag"][c] -= self.alpha[c]


def _views():
    z = lambda *s: torch.zeros(*s, dtype=torch.float64)
    return dict(pvol=z(N_CELLS), upAcc=z(N_CELLS, 3), stamp=torch.zeros(N_CELLS, dtype=torch.int32), uSource=z(N_CELLS, 3),
                uSourceDrag=z(N_CELLS))


def _particles():
    g = torch.Generator().manual_seed(3)
    pd = torch.rand(P, 5, generator=g, dtype=torch.float64)
    pd[:, 0] = torch.arange(P, dtype=torch.float64)
    pd[:, 1] *= 1e-3
    return pd


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    sh = pkg.sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pd = _particles()
    lo, hi = sh.shard_range(P, rank, world)
    v = _views()
    S = sh.ShardedCoupling(FakeEngine(v), dist, v, gaussian=True)
    found = torch.zeros(hi - lo, dtype=torch.int32)
    force = torch.zeros(hi - lo, 6, dtype=torch.float64)
    for _ in range(2):                                   # two steps: the accumulators must come back clean
        v["uSource"].zero_(); v["uSourceDrag"].zero_()
        S.step(1e-3, pd[lo:hi], hi - lo, found, force)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, lo, hi, force.numpy(), v["uSource"].numpy(), v["uSourceDrag"].numpy(), S.E.alpha.numpy()))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_cover_the_buffer():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    sh = g.load_package().sharded
    for n, w in ((10, 3), (7, 8), (1000000, 8), (0, 2)):
        r = [sh.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_shards_equal_one_domain_over_gloo():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    sh = g.load_package().sharded
    # single rank, whole buffer
    pd = _particles()
    v = _views()
    S = sh.ShardedCoupling(FakeEngine(v), None, v, gaussian=True)
    found = torch.zeros(P, dtype=torch.int32)
    force1 = torch.zeros(P, 6, dtype=torch.float64)
    for _ in range(2):
        v["uSource"].zero_(); v["uSourceDrag"].zero_()
        S.step(1e-3, pd, P, found, force1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted((q.get(timeout=180) for _ in ps), key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    force2 = np.concatenate([o[3] for o in out])
    np.testing.assert_allclose(force2, force1.numpy(), rtol=1e-12, atol=0)
    for o in out:                                        # every rank ends with the reduced fields
        np.testing.assert_allclose(o[4], v["uSource"].numpy(), rtol=1e-12, atol=1e-18)
        np.testing.assert_allclose(o[5], v["uSourceDrag"].numpy(), rtol=1e-12, atol=1e-18)
        np.testing.assert_allclose(o[6], S.E.alpha.numpy(), rtol=1e-12, atol=0)
