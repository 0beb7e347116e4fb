"""Scratch: one launch each of the non-headline kernels, for ncu (slow-branch tracks on cor_v1 and glider_v1, initial network)."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200.model import EncounterModel, UncorEncounterModel
from em_model_manned_bayes_b200.model_archive import materialize
paths = materialize(tempfile.mkdtemp(prefix="emb_models_"))
dev = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1


def run(label, fn, units):
    fn(1)
    torch.cuda.synchronize()
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(2 + r); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-40s %.3f ms  %.3e units/s" % (label, best, units / best * 1e3), flush=True)


cm = EncounterModel(paths["cor_v1"])
r = cm.sample_tracks(1 << 20, 60, seed=1, device=dev)
run("cor_v1 tracks 1M x 60", lambda k: cm.sample_tracks(1 << 20, 60, seed=k, device=dev, out=r), (1 << 20) * 60)
gm = UncorEncounterModel(paths["glider_v1"])
r2 = gm.sample_compact(1 << 20, 300, seed=1, device=dev, want_init=False)
run("glider_v1 tracks 1M x 300", lambda k: gm.sample_compact(1 << 20, 300, seed=k, device=dev, out=r2), (1 << 20) * 300)
g = EncounterModel(paths["glider_v1"])
run("glider_v1 initial 16M bins", lambda k: g.sample_initial(1 << 24, seed=k, device=dev, want_values=False, want_attempts=False), 1 << 24)
run("glider_v1 initial 16M bins+values", lambda k: g.sample_initial(1 << 24, seed=k, device=dev, want_values=True, want_attempts=False), 1 << 24)
