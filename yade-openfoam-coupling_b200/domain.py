"""Domain-decomposed multi-GPU step: one process per GPU, torch.distributed as plumbing.

What runs where in one coupled icoFoamYade time step on N GPUs:

    pressure solves (PCG, 90 % of the step)   z-slab decomposed inside libfycuda.so (csrc/fv_dist.cu): NCCL halo exchange
                                              of the search direction before every Amul, three 1-double all-reduces per
                                              iteration, slab-local DIC (OpenFOAM's decomposed preconditioner), gather of
                                              the solution planes at the end of a solve
    coupling (locate, weights, forces)        particles partitioned over the ranks -- by owner slab of the particle's
                                              cell ("migration": `owner_slab`, `migrate`) or by index -- per-cell sums
                                              all-reduced over NVLink (sharded.ShardedCoupling)
    FV assembly around the solves             replicated: every rank holds the whole box (the global k-d tree has to be
                                              replicated anyway for the cell lists to equal the single-domain ones,
                                              SURVEY.md H9), 8 % of the step

This module is the host-side plumbing: it hands the NCCL id from rank 0 to the others and creates the partition."""
import numpy as np


def slab_range(nz, rank, world):
    """k-planes [lo, hi) of rank's z slab -- the same arithmetic as csrc/fv_dist.cu (fvSlabRange)."""
    return (rank * int(nz)) // int(world), ((rank + 1) * int(nz)) // int(world)


def owner_slab(z, z0, hz, nz, world):
    """rank owning the k-plane of the cell containing height z (array): particle migration target."""
    k = np.clip(np.floor((np.asarray(z) - z0) / hz).astype(np.int64), 0, nz - 1)
    bounds = np.array([((r + 1) * nz) // world for r in range(world)])
    return np.searchsorted(bounds, k, side="right").astype(np.int32)


def broadcast_id(dist, uid, device=None):
    """rank 0's NCCL id (bytes) on every rank, through the job's torch.distributed group (any backend)."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(uid), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def init_domain(engine, dist, device=None, py=0):
    """Decomposes the engine's pressure solve over the ranks of `dist` (initialised torch.distributed) as a Py x Pz grid
    (py = 0: the library's choice, y first; py = 1: z slabs).  The particles' owners stay the z slabs of slab_range."""
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = engine.dist_unique_id() if rank == 0 else b"\0" * 128
    uid = broadcast_id(dist, uid, device)
    engine.dist_init(rank, world, uid, py)
    return engine.dist_info()


def solve_grid(ny, world, py=0):
    """(Py, Pz) the library picks for `world` ranks on a box with ny rows -- the arithmetic of csrc/fv_dist.cu."""
    njb = (int(ny) + 31) // 32
    if py == 0:
        py = max(c for c in range(1, world + 1) if world % c == 0 and c <= njb)
    return py, world // py


def _all_to_all(dist, recv, send, rc, sc):
    """rows of `send` split by sc -> rows of `recv` split by rc.  NCCL: one all_to_all_single; gloo (the CPU tests) has
    no all-to-all, so the same exchange is a batch of point-to-point operations."""
    import torch
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc)
        return
    me, ops = dist.get_rank(), []
    so = np.concatenate([[0], np.cumsum(sc)]).astype(int)
    ro = np.concatenate([[0], np.cumsum(rc)]).astype(int)
    recv[ro[me]:ro[me + 1]] = send[so[me]:so[me + 1]]
    for r in range(dist.get_world_size()):
        if r == me:
            continue
        if sc[r] > 0:
            ops.append(dist.P2POp(dist.isend, send[so[r]:so[r + 1]].contiguous(), r))
        if rc[r] > 0:
            ops.append(dist.P2POp(dist.irecv, recv[ro[r]:ro[r + 1]], r))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def migrate(dist, pdata, owner, device):
    """Particle migration: every rank sends each of its records to the rank that owns it (all-to-all of 80-byte
    records after a count exchange).  pdata [n][10] float64 tensor on `device`, owner [n] int tensor.  Returns the
    records this rank owns now and, for the way back, (send order, send counts, receive counts)."""
    import torch
    world = dist.get_world_size()
    order = torch.argsort(owner, stable=True)
    send = pdata[order].contiguous()
    scount = torch.bincount(owner, minlength=world).to(torch.int64)
    counts = [torch.empty_like(scount) for _ in range(world)]
    dist.all_gather(counts, scount)                       # counts[r][q] = records rank r sends to rank q
    me = dist.get_rank()
    sc, rc = [int(x) for x in scount.tolist()], [int(counts[r][me]) for r in range(world)]
    recv = torch.empty((sum(rc), pdata.shape[1]), dtype=pdata.dtype, device=device)
    _all_to_all(dist, recv, send, rc, sc)
    if recv.is_cuda:
        torch.cuda.current_stream().synchronize()         # the engine reads the records on its own stream
    return recv, (order, sc, rc)


def migrate_back(dist, values, route, device):
    """The inverse route for per-particle results (forces [m][6], found [m]): back to the rank the record came from,
    in that rank's original order."""
    import torch
    order, sc, rc = route
    out = torch.empty((sum(sc),) + tuple(values.shape[1:]), dtype=values.dtype, device=device)
    _all_to_all(dist, out, values.contiguous(), sc, rc)
    res = torch.empty_like(out)
    res[order] = out
    if res.is_cuda:
        torch.cuda.current_stream().synchronize()
    return res
