"""Multi-GPU plumbing for the hot path: independent replicas, one process per GPU, no data-path collective.

TTA adapts every target volume independently (dg_tta/tta/tta.py:157-182: fresh deepcopy of the network and a fresh
optimiser per sample and ensemble member) and MIND / GIN / the sampler are per-sample functions, so the path shards
by volume.  torch.distributed is only used to agree on the work split and to combine timings
(max over ranks) — on NCCL for GPUs, on gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_items(n_items, rank, world_size):
    """Round-robin assignment of volume (or ensemble-member) indices to ranks — the same split a per-GPU launcher
    of `run_tta` would apply to `tta_data_filepaths` (restartable thanks to tta.py:166-173's skip-if-exists)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_items, world_size))


def max_over_ranks(value, device="cpu"):
    """Slowest rank decides (timing rule: max over ranks).  Identity when not running under torch.distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def aggregate_throughput(units_this_rank, seconds_this_rank, device="cpu"):
    """Whole-job throughput = units processed by all ranks / time of the slowest rank."""
    total = sum_over_ranks(units_this_rank, device)
    slowest = max_over_ranks(seconds_this_rank, device)
    return total / slowest if slowest > 0 else float("inf")
