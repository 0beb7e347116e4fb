"""Scratch A/B timing of the initial-network kernels on one GPU (smem-staged vs gathered vs generic)."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200 import _lib as L
from em_model_manned_bayes_b200.model import EncounterModel
from em_model_manned_bayes_b200.model_archive import materialize
paths = materialize(tempfile.mkdtemp(prefix="emb_models_"))
lib = L.lib()
def t(name, n, vals, generic=False, reps=5):
    m = EncounterModel(paths[name])
    lib.emb_debug_force_generic(int(generic))
    m.sample_initial(n, seed=1, device="cuda:0", want_attempts=False, want_values=vals)
    torch.cuda.synchronize()
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.sample_initial(n, seed=2 + r, device="cuda:0", want_attempts=False, want_values=vals); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    lib.emb_debug_force_generic(0)
    print("%-36s n=%9d values=%d generic=%d  %.3f ms  %.3e samples/s" % (name, n, vals, generic, best, n / best * 1e3), flush=True)
for name in ("glider_v1", "uncor_allcode_fwsingle_v1", "terminal_v3_radar_encounter_model"):
    for vals in (False, True):
        t(name, 1 << 24, vals)
t("glider_v1", 100_000_000, False)
t("glider_v1", 100_000_000, True)
t("glider_v1", 1 << 24, True, generic=True)
