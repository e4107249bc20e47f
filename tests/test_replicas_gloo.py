"""CPU, world_size 2 over gloo: the multi-process plumbing bench.py uses for --gpus N (replicas):
distinct particle batches per rank, MAX-over-ranks timing, whole-job throughput."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    from tests import cases
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    pkg = g.load_package()
    rep = pkg.replicas
    assert rep.world() == (rank, world, rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seed = rep.particle_seed(7, rank)
    pd = cases.particles(64, seed, radius=0.01)
    ms = [10.0 + 5.0 * rank, 20.0 - 3.0 * rank]          # rank 1 is slower on the first, faster on the second
    red = rep.slowest_rank_ms(ms, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, seed, float(pd[:, :3].sum()), red, rep.job_throughput(world, 10, red[0])))


def test_two_replicas_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, c0, red0, v0), (r1, s1, c1, red1, v1) = out
    assert (s0, s1) == (7, 1007) and c0 != c1             # distinct particle batches
    assert red0 == red1 == [15.0, 20.0]                   # MAX over ranks, element-wise
    assert v0 == v1 == 2 * 10 / 15e-3                     # whole-job steps/s from the slowest rank
