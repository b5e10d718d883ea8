"""Multi-GPU plumbing of the path: channels shard, one sum-reduce of the stereo master.

Chains share no state (the reference already runs one goroutine per channel,
controller/controller.go:3339-3341, 2682-2705), so each rank owns a contiguous range
of channels with its chains, spatializer slice and (on the owner rank) the tuner.
The only exchange step is the spatializer master (spatializer/spatializer.go:140-335):
every rank mixes its own channels into a partial [2][n] and the partials are summed
onto rank 0 - torch.distributed.reduce (NCCL over NVLink on the GPU box, gloo in the
CPU tests).
"""
import torch.distributed as dist


def shard_channels(total, world, rank):
    """Contiguous channel range [lo, hi) of `rank`; the first total % world ranks get one more."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of_channel(channel, total, world):
    for r in range(world):
        lo, hi = shard_channels(total, world, r)
        if lo <= channel < hi:
            return r
    raise ValueError(channel)


def reduce_stereo_master(master, dst=0):
    """Sum the per-rank partial masters ([2][n] float64 tensor, in place) onto rank `dst`."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(master, dst=dst, op=dist.ReduceOp.SUM)
    return master
