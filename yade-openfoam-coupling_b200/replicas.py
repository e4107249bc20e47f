"""Multi-GPU plumbing of this round: N independent domain replicas, one process per GPU.

No collective sits on the data path: every rank owns a whole domain and its own particle batch.  torch.distributed
(NCCL on GPUs, gloo in the CPU tests) only (a) hands every rank a distinct particle seed and (b) reduces the timed
region to the slowest rank, from which the whole-job throughput is derived."""
import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def particle_seed(base_seed, rank):
    """Every replica draws its own particle batch: seeds are base + 1000 * rank (distinct for < 1000 ranks)."""
    return int(base_seed) + 1000 * int(rank)


def slowest_rank_ms(values_ms, dist=None, device=None):
    """Element-wise MAX over ranks of a list of per-rank times (ms).  `dist` = torch.distributed (initialised)
    or None for a single process."""
    import torch
    t = torch.tensor(list(values_ms), dtype=torch.float64, device=device if device is not None else "cpu")
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def job_throughput(world_size, steps, ms):
    """Whole-job coupled timesteps/s: every replica advanced `steps` steps in `ms` (the slowest rank's time)."""
    return world_size * steps / (ms * 1e-3)
