"""GPU x2 (NCCL): particle-sharded coupling == the single-GPU result.  Needs two devices; the
round-end single-GPU run skips it (run it with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_BOX, P = 24, 20000


def _case(pkg):
    mp_ = pkg.box_mesh(N_BOX, N_BOX, N_BOX)
    pd = cases.particles(P, 21, radius=0.1 / N_BOX, moving=True)
    pd[:11, :3] += 3.0                                    # some particles outside the mesh
    f = cases.fields_for(mp_["C"])
    return mp_, pd, f


def _setup(pkg, mp_, f, gaussian, device):
    E = pkg.Engine(mp_, device=device)
    E.set_properties(cases.RHOP, cases.RHOF, cases.NU, gaussian)
    for k in ("U", "gradP", "divT", "vGrad"):
        E.upload(k, f[k])
    return E


def _worker(rank, world, port, gaussian, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    torch.cuda.set_device(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    mp_, pd, f = _case(pkg)
    E = _setup(pkg, mp_, f, gaussian, rank)
    sh = pkg.sharded
    lo, hi = sh.shard_range(P, rank, world)
    d_pd = torch.from_numpy(pd[lo:hi].copy()).cuda()
    d_found = torch.zeros(hi - lo, dtype=torch.int32, device="cuda")
    d_force = torch.zeros(hi - lo, 6, dtype=torch.float64, device="cuda")
    S = sh.ShardedCoupling(E, dist, sh.device_views(E), gaussian, sh.external_stream_ctx(E))
    outs = []
    for it in range(2):
        S.step(1e-3, d_pd.data_ptr(), hi - lo, d_found.data_ptr(), d_force.data_ptr())
        E.synchronize()
        outs.append(dict(found=d_found.cpu().numpy(), force=d_force.cpu().numpy(), uSource=E.download("uSource"),
                         uSourceDrag=E.download("uSourceDrag"), alpha=E.download("alpha"), uParticle=E.download("uParticle")))
        E.set_source_zero()
    dist.barrier()
    dist.destroy_process_group()
    E.close()
    q.put((rank, outs))


@pytest.mark.parametrize("gaussian", [True, False])
def test_two_gpu_shards_equal_single_gpu(pkg, gaussian):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp_, pd, f = _case(pkg)
    E = _setup(pkg, mp_, f, gaussian, 0)
    ref = []
    for it in range(2):
        fo, Fo = E.set_particle_action(1e-3, pd)
        ref.append(dict(found=fo.copy(), force=Fo.copy(), uSource=E.download("uSource"), uSourceDrag=E.download("uSourceDrag"),
                        alpha=E.download("alpha"), uParticle=E.download("uParticle")))
        E.set_source_zero()
    E.close()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, gaussian, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for it in range(2):
        found = np.concatenate([out[r][it]["found"] for r in range(2)])
        force = np.concatenate([out[r][it]["force"] for r in range(2)])
        assert np.array_equal(found, ref[it]["found"])
        assert cases.rel_l2(force, ref[it]["force"]) <= cases.TOL
        for r in range(2):                                # every rank holds the reduced fields
            for k in ("uSource", "uSourceDrag", "alpha", "uParticle"):
                assert cases.rel_l2(out[r][it][k], ref[it][k]) <= cases.TOL, (r, k)
