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


def sample_tracks_multi(model, n: int, T: int, seed: int = 0, first_sample: int = 0, opts=None, n_devices: int = 0,
                        want_hist: bool = True):
    """One process, all GPUs: emb_sample_tracks_multi (include/emb200.h).  Host-memory outputs, one TrackResult per device
    shard, plus the globally reduced verification histograms (the single NCCL collective happens inside the library).
    -> (list of TrackResult in shard order, hist_initial (n_initial, 64) uint64 or None, hist_transition (n_dyn, 64) or None)"""
    import ctypes as C

    import numpy as np

    from . import _lib as L
    from .model import TrackResult, _ptr
    lib = L.lib()
    D = n_devices if n_devices > 0 else lib.emb_device_count()
    if D <= 0:
        raise L.EmbError(L.EMB_E_CUDA, "no CUDA device available (libemb200 has no CPU path)")
    o = opts if opts is not None else model._opts()
    o.mem = L.EMB_MEM_HOST
    outs = (L.TrackOut * D)()
    res = []
    ni = model.n_initial
    for d in range(D):
        f, c = C.c_int64(), C.c_int64()
        lib.emb_shard_range(n, d, D, C.byref(f), C.byref(c))
        cnt = int(c.value)
        nb, nv = int(lib.emb_tracks_bins_len(model._h, cnt, T)), int(lib.emb_tracks_values_len(model._h, cnt, T))
        r = TrackResult(n=cnt, T=T, dyn_vars=[int(v) for v in model.temporal_map[:, 0]], tv_vars=list(model.timevarying_vars),
                        bins_tiled=np.zeros(max(nb, 1), dtype=np.int8), values_tiled=np.zeros(max(nv, 1), dtype=np.float32),
                        init_bins=np.zeros((ni, cnt), dtype=np.int8), init_values=np.zeros((ni, cnt), dtype=np.float64),
                        attempts=np.zeros(cnt, dtype=np.uint16))
        res.append(r)
        outs[d] = L.TrackOut(_ptr(r.bins_tiled), _ptr(r.values_tiled), _ptr(r.init_bins), _ptr(r.init_values), _ptr(r.attempts),
                             None, None)
    hi = np.zeros((ni, 64), dtype=np.uint64) if want_hist else None
    ht = np.zeros((model.n_dyn, 64), dtype=np.uint64) if want_hist else None
    rng = L.Rng(int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_sample))
    rc = lib.emb_sample_tracks_multi(model._h, C.byref(rng), n, T, C.byref(o), D, outs, _ptr(hi), _ptr(ht))
    if rc:
        raise L.EmbError(rc, (lib.emb_multi_last_error() or b"").decode() or "emb_sample_tracks_multi failed")
    return res, hi, ht
