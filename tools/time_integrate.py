"""Times sample2track's Euler loop on the GPU (emb_tracks_integrate) after emb_sample_tracks: python tools/time_integrate.py [n] [T]"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200.model import UncorEncounterModel  # noqa: E402
from em_model_manned_bayes_b200.model_archive import materialize  # noqa: E402
from em_model_manned_bayes_b200.sample2track import integrate_tracks  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 600
paths = materialize(tempfile.mkdtemp(prefix="emb_models_"))
m = UncorEncounterModel(paths["uncor_allcode_fwsingle_v1"])
res = m.sample_compact(n, T, seed=1, device="cuda:0")
xyz, good = integrate_tracks(m, res, device="cuda:0")
torch.cuda.synchronize()
best = 1e9
for r in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    integrate_tracks(m, res, device="cuda:0")
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("emb_tracks_integrate n=%d T=%d: %.3f ms  %.3e track-timesteps/s (reads 12 B + writes 12 B per unit: %.0f GB/s), good %.3f"
      % (n, T, best, n * T / best * 1e3, 24.0 * n * T / best / 1e6, float(good.float().mean())))
