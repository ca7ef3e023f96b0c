"""Host-side logic for running the encoder on several GPUs, one process per GPU.

BZip2 streams (archive entries) and the chunks inside them are independent units:
ranks never exchange data on the encode path.  What the ranks do share is bookkeeping: which
entries each rank takes, and the slowest rank's time (the job's time).  These helpers work with any
torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_entries(sizes, world_size, rank):
    """Deterministic longest-processing-time-first assignment of entries (by size) to ranks.
    Returns the sorted list of entry indices this rank encodes.  Every rank computes the same
    partition from the same `sizes`, so no communication is needed."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    owner = [0] * len(sizes)
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += int(sizes[i])
    return sorted(i for i in range(len(sizes)) if owner[i] == rank)


def stream_seed(base_seed, rank):
    """Seed of the synthetic stream a rank encodes in the weak-scaling bench."""
    return int(base_seed) + int(rank)


def max_over_ranks(value, device="cpu"):
    """The job's time is the slowest rank's time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_mbps(bytes_this_rank, seconds_this_rank, device="cpu"):
    """Whole-job throughput: bytes of all ranks / time of the slowest rank."""
    total = sum_over_ranks(bytes_this_rank, device)
    t = max_over_ranks(seconds_this_rank, device)
    return total / 1e6 / t if t > 0 else 0.0
