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

from em_model_manned_bayes_b200.sample2track import sample_tracks_xyz  # noqa: E402


def timed(fn):
    fn()
    torch.cuda.synchronize()
    b = 1e9
    for r in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        b = min(b, e0.elapsed_time(e1))
    return b


del xyz, good
t2 = timed(lambda: integrate_tracks(m, m.sample_compact(n, T, seed=5, device="cuda:0", out=res, want_init=True), device="cuda:0"))
print("two passes (emb_sample_tracks + emb_tracks_integrate, incl. output allocation): %.3f ms" % t2)
del res
torch.cuda.empty_cache()
tf = timed(lambda: sample_tracks_xyz(m, n, T, seed=5, sample_opts=m.uncor_opts(), device="cuda:0"))
print("fused (emb_sample_tracks_xyz, xyz + is_good only, incl. output allocation): %.3f ms  %.3e track-timesteps/s" % (tf, n * T / tf * 1e3))
tg = timed(lambda: sample_tracks_xyz(m, n, T, seed=5, sample_opts=m.uncor_opts(), device="cuda:0", want_xyz=False))
print("fused, is_good only: %.3f ms" % tg)
