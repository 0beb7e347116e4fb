"""Scratch timing of the slow-branch shapes (glider_v1, cor_v1) -- see tools/quick_time.py."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200.model import EncounterModel, UncorEncounterModel
from em_model_manned_bayes_b200.model_archive import materialize
paths = materialize(tempfile.mkdtemp(prefix="emb_models_"), names=["glider_v1", "cor_v1"])
def run(m, n, T, fn):
    res = fn(n, T, 1, None)
    torch.cuda.synchronize()
    best = 1e9
    for r in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(n, T, 2 + r, res); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
g = UncorEncounterModel(paths["glider_v1"])
ms = run(g, 1 << 20, 300, lambda n, T, s, out: g.sample_compact(n, T, seed=s, device="cuda:0", want_init=False, out=out))
print("glider_v1 n=%d T=300: %.3f ms %.3e/s" % (1 << 20, ms, (1 << 20) * 300 / ms * 1e3))
c = EncounterModel(paths["cor_v1"])
ms = run(c, 10_000_000, 60, lambda n, T, s, out: c.sample_tracks(n, T, seed=s, device="cuda:0", out=out))
print("cor_v1 n=1e7 T=60: %.3f ms %.3e/s" % (ms, 1e7 * 60 / ms * 1e3))
