"""Particle-sharded multi-GPU coupling.

ONE Yade particle buffer is split into contiguous slices, one per rank (one process per GPU).  Every rank holds the
whole mesh, the k-d tree and the fluid fields (replicated read-only data) and runs the passes of the
coupling operator on its slice; the per-cell partial sums are added up over NVLink with NCCL all-reduces between
the passes, which is the one real exchange step of this half of the path:

    pass 0  locate + weights + accumulate   ->  all-reduce SUM  pvol [N], upAcc [N][3];  MAX stamp [N]
    pass 1  void fraction (identical on every rank)
    pass 2  forces of the slice             ->  all-reduce SUM  uSource [N][3], uSourceDrag [N]

The result is the single-domain, single-buffer result (the reference's serial-Yade semantics) up to the order of
the floating-point additions.  The collectives are issued on the engine's own CUDA stream, so they are ordered
with its kernels without host synchronisation.  torch.distributed is plumbing only; on CPU (gloo) the same class
drives a stand-in engine in the tests."""


def shard_range(n, rank, world):
    """[lo, hi) of rank's contiguous slice of n particles (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevArray:
    """A device buffer owned by the engine, exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2,
                                             strides=None)


def device_views(engine):
    """torch views of the buffers the passes reduce: dict(pvol, upAcc, stamp, uSource, uSourceDrag)."""
    import torch
    N = engine.N
    pvol, up, stamp = engine.device_accumulators()
    mk = lambda ptr, shape, ts: torch.as_tensor(_DevArray(ptr, shape, ts), device="cuda")
    return dict(pvol=mk(pvol, (N,), "<f8"), upAcc=mk(up, (N, 3), "<f8"), stamp=mk(stamp, (N,), "<i4"),
                uSource=mk(engine.device_field("uSource"), (N, 3), "<f8"),
                uSourceDrag=mk(engine.device_field("uSourceDrag"), (N,), "<f8"))


class ShardedCoupling:
    """setParticleAction over `dist` (torch.distributed, initialised) with this rank's slice of the buffer.

    engine: an object with coupling_begin(dt), coupling_pass_device(pass, d_pdata, n, d_found, d_force), gaussian
    (bool) -- yade_openfoam_coupling_b200.Engine on GPUs.  views: the tensors to reduce (device_views(engine)).
    stream: a torch stream context factory (torch.cuda.stream(ExternalStream(engine.stream()))) or None."""

    def __init__(self, engine, dist, views, gaussian, stream_ctx=None):
        self.E, self.dist, self.v, self.gaussian = engine, dist, views, bool(gaussian)
        self.stream_ctx = stream_ctx

    def _reduce(self, names_ops):
        d = self.dist
        if d is None or not d.is_initialized() or d.get_world_size() == 1:
            return
        if self.stream_ctx is not None:
            with self.stream_ctx():
                for name, op in names_ops:
                    d.all_reduce(self.v[name], op=op)
        else:
            for name, op in names_ops:
                d.all_reduce(self.v[name], op=op)

    def step(self, dt, d_pdata, n, d_found, d_force):
        """d_pdata / d_found / d_force: this rank's slice (device pointers on GPUs)."""
        d = self.dist
        SUM = d.ReduceOp.SUM if d is not None else None
        MAX = d.ReduceOp.MAX if d is not None else None
        self.E.coupling_begin(dt)
        self.E.coupling_pass_device(0, d_pdata, n, d_found, d_force)
        if self.gaussian:
            self._reduce([("pvol", SUM), ("upAcc", SUM), ("stamp", MAX)])
        self.E.coupling_pass_device(1, d_pdata, n, d_found, d_force)
        self.E.coupling_pass_device(2, d_pdata, n, d_found, d_force)
        self._reduce([("uSource", SUM)] + ([("uSourceDrag", SUM)] if self.gaussian else []))


def external_stream_ctx(engine):
    """Context factory that makes the engine's CUDA stream torch's current stream (orders NCCL with the kernels)."""
    import torch
    ext = torch.cuda.ExternalStream(engine.stream())
    return lambda: torch.cuda.stream(ext)
