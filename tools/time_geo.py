"""Scratch: terminal encounter-geometry sampling (CorTerminalModel.sample: 15 variables, r <= 36, box rejection)."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200.model import CorTerminalModel
from em_model_manned_bayes_b200.model_archive import materialize
p = materialize(tempfile.mkdtemp(prefix="emb_models_"), names=["terminal_v3_radar_encounter_model"])["terminal_v3_radar_encounter_model"]
m = CorTerminalModel(p)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
m.sample_raw(n, seed=1, device="cuda:0")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for k in range(3):
    e0.record(); m.sample_raw(n, seed=2 + k, device="cuda:0"); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("terminal geometry n=%d: %.3f ms %.3e encounters/s" % (n, best, n / best * 1e3))
