"""Batch partitioning across the GPUs of one box: independent trajectories, contiguous blocks,
SDF grid replicated, no collective on the data path — only a host-side gather / the timing reduction.

Trajectory b (global index) belongs to problem (b // starts_per_problem) % num_problems; rank r of `world`
owns the contiguous block [lo, hi) and calls the evaluator with traj_offset = lo, so per-trajectory results
are bit-identical for every world size."""


def shard_range(total, world, rank):
    """Contiguous block of `total` items for `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def problem_of(global_traj, starts_per_problem, num_problems):
    return (global_traj // starts_per_problem) % num_problems


def max_over_ranks(value, dist=None):
    """Timing reduction: MAX over ranks (identity when torch.distributed is not initialised)."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return value
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
