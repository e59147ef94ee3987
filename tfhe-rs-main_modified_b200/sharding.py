"""Batch sharding across GPUs: independent polynomials, contiguous slices, no collective on the
data path.  The split rule is the reference CUDA backend's
(backends/tfhe-cuda-backend/cuda/src/utils/helper_multi_gpu.cu:57-88): batch // G each, the
remainder goes one-by-one to the first GPUs."""


def shard_range(batch, world_size, rank):
    """[begin, end) of the polynomials rank `rank` of `world_size` owns."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(batch, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(batch, world_size):
    return [shard_range(batch, world_size, r)[1] - shard_range(batch, world_size, r)[0] for r in range(world_size)]


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised).
    Used for device-timed intervals: a multi-GPU number is the slowest rank's."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
