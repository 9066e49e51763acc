"""Batch mode over several GPUs: a formula does not shard, so a batch of independent CNF
instances is partitioned over the ranks - one process, one sigma context and one stream per GPU, no
collective on the data path.  torch.distributed is used only to agree on the timing (barrier + max over
ranks) and to add up the work done."""
from __future__ import annotations


def rank_instances(n_instances, rank, world):
    """Indices of the instances rank `rank` owns: round-robin, so sizes sorted descending spread evenly."""
    return list(range(rank, n_instances, world))


def rank_seed(base_seed, rank):
    """Weak-scaling bench: every rank generates its own instance of the same shape."""
    return base_seed + rank


def max_over_ranks(x, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def throughput(units_all_ranks, ms_max_over_ranks):
    """Whole-job rate: the units ALL ranks processed divided by the slowest rank's time."""
    return units_all_ranks / (ms_max_over_ranks / 1e3)
