"""Scratch timing helper for gpurun sessions (not the bench contract; see bench.py)."""
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200.model import EncounterModel, UncorEncounterModel  # noqa: E402
from em_model_manned_bayes_b200.model_archive import materialize  # noqa: E402

paths = materialize(tempfile.mkdtemp(prefix="emb_models_"))
dev = "cuda:0"


def time_tracks(name, n, T, reps=3):
    m = UncorEncounterModel(paths[name])
    res = m.sample_compact(n, T, seed=1, device=dev, want_init=False)
    torch.cuda.synchronize()
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.sample_compact(n, T, seed=2 + r, device=dev, out=res)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # checksum of the last pass (same seed for every library variant): variants must agree bit for bit
    cs = (int(res.bins_tiled.view(torch.int32).sum(dtype=torch.int64)), int(res.values_tiled.view(torch.int32).sum(dtype=torch.int64)))
    print("%s n=%d T=%d: %.3f ms  %.3e track-timesteps/s  checksum %x %x" % (name, n, T, best, n * T / best * 1e3, cs[0] & (2**64 - 1), cs[1] & (2**64 - 1)), flush=True)


def time_events(name, n, T, reps=3):
    m = UncorEncounterModel(paths[name])
    r = m.sample_events_uncor(n, T, seed=1, device=dev, want_init=False)
    cap = int(r.total * 1.05)
    torch.cuda.synchronize()
    best = 1e9
    for k in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = m.sample_events_uncor(n, T, seed=2 + k, device=dev, want_init=False, capacity=cap)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s events n=%d T=%d: %.3f ms  %.3e track-timesteps/s  (%.1f rows/track, %.1f MB)" %
          (name, n, T, best, n * T / best * 1e3, r.total / n, r.total * 8 / 1e6), flush=True)


def time_initial(name, n, reps=3, want_values=True, fp32=False):
    m = EncounterModel(paths[name])
    m.sample_initial(n, seed=1, device=dev, want_attempts=False, want_values=want_values, values_fp32=fp32)
    torch.cuda.synchronize()
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.sample_initial(n, seed=2 + r, device=dev, want_attempts=False, want_values=want_values, values_fp32=fp32)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s initial n=%d values=%s%s: %.3f ms  %.3e samples/s" % (name, n, want_values, " fp32" if fp32 else "", best, n / best * 1e3), flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if len(sys.argv) > 1 and sys.argv[1] == "tracks":
        time_tracks("uncor_allcode_fwsingle_v1", 1250000, 600, reps=5)
        sys.exit(0)
    time_tracks("uncor_1200code_v2p1", 1 << 20, 300)
    time_tracks("uncor_allcode_fwsingle_v1", 1 << 20, 600)
    time_tracks("glider_v1", 1 << 20, 300)
    time_events("uncor_allcode_fwsingle_v1", 1 << 20, 600)
    time_initial("glider_v1", 1 << 24)
    time_initial("glider_v1", 1 << 24, fp32=True)
    time_initial("glider_v1", 1 << 24, want_values=False)
    time_initial("uncor_allcode_fwsingle_v1", 1 << 24, want_values=False)
    time_initial("terminal_v3_radar_encounter_model", 1 << 22)
