"""Batch partitioning for one-process-per-GPU runs: the batch (tiles, m_blocks, or the columns of N
for fsspmdm) is the only shard axis, ranks work on contiguous ranges and exchange nothing on the data path.
Only torch.distributed plumbing lives here; the kernels are reached through the C ABI by the caller."""


def shard_range(total, world, rank, granule=1):
    """[begin, end) of `rank`: contiguous, sizes differ by at most one granule, multiples of `granule`
    (fsspmdm needs N-slices that are multiples of the vector length 64/typesize; BCSC groups of m_blocks)."""
    if world <= 0 or not (0 <= rank < world) or total < 0 or granule <= 0:
        raise ValueError("bad shard request")
    units, rem = divmod(total, granule)
    base, extra = divmod(units, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    b, e = begin * granule, end * granule
    if rank == world - 1:
        e += rem                      # a ragged tail stays with the last rank
    return b, e


def weak_batch(per_gpu, world):
    """weak scaling: every rank processes `per_gpu` units; the job total is per_gpu * world"""
    return per_gpu, per_gpu * world


def aggregate(dist, local_units, local_ms, device=None):
    """whole-job figures: units summed over ranks, time = MAX over ranks (the bench contract)"""
    import torch
    t = torch.tensor([float(local_units), 0.0], dtype=torch.float64, device=device)
    m = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(t[0].item()), float(m[0].item())
