"""Multi-GPU plumbing: sample-index sharding and the single verification-histogram reduction.

The hot path shards by global sample index with no inter-GPU traffic: rank g owns
[g*ceil(n/G), min(n, (g+1)*ceil(n/G))) and the Philox stream is keyed by the *global* index, so the
union over ranks is identical for any GPU count.  The only collective is one all-reduce (NCCL on
GPUs, gloo in the CPU tests) of the uint64 verification histograms."""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int):
    """-> (first_sample, count) of `rank`; contiguous, ceil-sized shards, empty tail shards allowed."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-int(n_total) // world)
    first = min(int(n_total), rank * per)
    return first, max(0, min(int(n_total), first + per) - first)


def allreduce_histograms(*hists):
    """Sum the given int64 histogram tensors over all ranks with ONE collective call (they are packed
    into one flat buffer first).  No-op when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return hists
    flat = torch.cat([h.reshape(-1) for h in hists])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    o = 0
    for h in hists:
        h.copy_(flat[o:o + h.numel()].reshape(h.shape))
        o += h.numel()
    return hists
