"""N>1 host logic on CPU: world_size-2 gloo ranks shard by global sample index (no data-path
collective) and reduce verification histograms with ONE all-reduce.  The per-rank samples come from
the TEST-ONLY host emulation of the device code (no GPU here); on the B200 the same helper runs over
NCCL inside bench.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from em_model_manned_bayes_b200.shard import allreduce_histograms, shard_range


def test_shard_range_covers_everything():
    for n, w in [(10, 1), (10, 2), (10, 3), (10_000_000, 8), (3, 8), (0, 4)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert sum(c for _, c in spans) == n
        pos = 0
        for first, cnt in spans:
            assert first == pos or cnt == 0
            pos += cnt
    assert shard_range(10_000_000, 3, 8) == (3_750_000, 1_250_000)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, path, tmp):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import EmuModel
    from oracle.em_read import em_read
    p = em_read(path)
    n_total, T, seed = 37, 40, 3
    first, cnt = shard_range(n_total, rank, world)
    em = EmuModel(path)
    r = em.sample_tracks(p.n_initial, 3, 4, cnt, T, seed, first, EmuModel.opts(p.n_initial), hist=True)
    hi = torch.from_numpy(r["hist_initial"].astype(np.int64))
    ht = torch.from_numpy(r["hist_transition"].astype(np.int64))
    allreduce_histograms(hi, ht)
    np.savez(os.path.join(tmp, "rank%d.npz" % rank), bins=r["bins"], values=r["values"], hi=hi.numpy(), ht=ht.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_histogram_reduce(model_paths, tmp_path):
    path = model_paths["uncor_1200code_v2p1"]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, path, str(tmp_path)), nprocs=2, join=True)
    from helpers import EmuModel
    from oracle.em_read import em_read
    p = em_read(path)
    whole = EmuModel(path).sample_tracks(p.n_initial, 3, 4, 37, 40, 3, 0, EmuModel.opts(p.n_initial), hist=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(2)]
    assert np.array_equal(np.concatenate([q["bins"] for q in parts]), whole["bins"])
    assert np.array_equal(np.concatenate([q["values"] for q in parts]), whole["values"])
    for q in parts:   # every rank holds the global histogram after the single all-reduce
        assert np.array_equal(q["hi"], whole["hist_initial"].astype(np.int64))
        assert np.array_equal(q["ht"], whole["hist_transition"].astype(np.int64))


def _terminal_worker(rank, world, port, traj_dir, geo_path, tmp):
    here = os.path.dirname(os.path.abspath(__file__))
    for q in (os.path.dirname(here), here):
        if q not in sys.path:
            sys.path.insert(0, q)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import test_terminal_traj as TT
    from em_model_manned_bayes_b200.synthetic import write_terminal_model_set
    paths = write_terminal_model_set(traj_dir)
    geo = np.load(geo_path)
    first, cnt = shard_range(geo.shape[1], rank, world)
    rc, traj, ln = TT.emu_propagate(paths, np.ascontiguousarray(geo[:, first:first + cnt]), 77, first, 60)
    assert rc == 0
    states = torch.tensor([int(ln.astype(np.int64).sum())])
    dist.all_reduce(states, op=dist.ReduceOp.SUM)           # what bench.py reduces for configs[4]: total trajectory states
    np.savez(os.path.join(tmp, "term%d.npz" % rank), traj=traj, len=ln, states=states.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_terminal_chains_are_shard_invariant(golden, tmp_path):
    """configs[4] shards by encounter index like everything else: rank r propagates encounters [first, first+cnt) with
    first_sample = first, and the union equals the single-rank result bit for bit (the chains are keyed by the global
    encounter index, emb_terminal.cuh)."""
    import test_terminal_traj as TT
    from em_model_manned_bayes_b200.synthetic import write_terminal_model_set
    traj_dir = str(tmp_path / "traj")
    paths = write_terminal_model_set(traj_dir)
    geo = TT.geo_from_golden(golden, 13)
    geo_path = str(tmp_path / "geo.npy")
    np.save(geo_path, geo)
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_terminal_worker, args=(2, port, traj_dir, geo_path, str(tmp_path)), nprocs=2, join=True)
    rc, traj, ln = TT.emu_propagate(paths, geo, 77, 0, 60)
    assert rc == 0
    parts = [np.load(os.path.join(str(tmp_path), "term%d.npz" % r)) for r in range(2)]
    got = np.concatenate([q["traj"] for q in parts], axis=3)
    assert np.array_equal(np.concatenate([q["len"] for q in parts], axis=1), ln)
    assert np.array_equal(np.isnan(got), np.isnan(traj)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(traj))
    assert all(int(q["states"][0]) == int(ln.astype(np.int64).sum()) for q in parts)
