"""Batch sharding for N GPUs.  Lanes are independent, so the only multi-GPU strategy is a contiguous
split with no collective.  The same split is used by run_sharded() in
csrc/engine.cu (mldsa_*_batch_multi) and by bench.py's one-process-per-GPU mode."""


def shard_range(n, rank, world):
    """Lanes [lo, hi) of rank `rank` out of `world`: contiguous, disjoint, covering, balanced to +-1."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return n * rank // world, n * (rank + 1) // world


def all_shards(n, world):
    return [shard_range(n, r, world) for r in range(world)]


def max_over_ranks(value, dist=None):
    """Timing rule for multi-GPU numbers: the slowest rank defines the step."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
