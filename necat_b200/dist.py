"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed.

The path shards without a data-path collective except for one exchange step: the 28-byte candidate
records found by the rank that mapped a query shard must reach every rank that corrects one of the two
reads involved -- what `cat` + oc2pcan do through the file system in the reference
(partition_candidates/pcan.c:58-99).  Records are tiny (3.4 MB for E. coli), so an all-gather is used and
each rank then keeps the templates it owns.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous query-read range [lo, hi) of a rank for the mapping half."""
    return (n * rank) // world, (n * (rank + 1)) // world


def owns_template(ids, rank, world):
    """Template ownership for the correction half: id mod world (matches nb_consensus_job_sharded)."""
    return (np.asarray(ids) % world) == rank


def all_gather_records(packed, device, world):
    """All-gather (n_i, 7) uint32 record arrays of unequal length; returns their concatenation in rank order.
    Works with NCCL (device = cuda) and gloo (device = cpu)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return packed
    mine = torch.from_numpy(np.ascontiguousarray(packed, np.uint32).reshape(-1).view(np.int32).copy()).to(device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([mine.numel()], dtype=torch.int64, device=device))
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    pad = torch.zeros(width, dtype=torch.int32, device=device)
    pad[:mine.numel()] = mine
    bufs = [torch.empty(width, dtype=torch.int32, device=device) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = [b[:c].cpu().numpy().view(np.uint32).reshape(-1, 7) for b, c in zip(bufs, counts)]
    return np.concatenate(parts) if parts else np.zeros((0, 7), np.uint32)
