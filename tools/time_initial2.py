"""Scratch: glider_v1 initial network (BASELINE configs[1]) timing at 100M samples: bins only, fp32 values, fp64 values."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200.model import EncounterModel
from em_model_manned_bayes_b200.model_archive import materialize
paths = materialize(tempfile.mkdtemp(prefix="emb_models_"), names=["glider_v1"])
g = EncounterModel(paths["glider_v1"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
for kw in (dict(want_values=False), dict(want_values=True, values_fp32=True), dict(want_values=True)):
    buf = g.sample_initial(n, seed=1, device="cuda:0", want_attempts=False, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(3):
        g.sample_initial(n, seed=2 + k, device="cuda:0", want_attempts=False, out=buf, enqueue_only=True, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("glider_v1 initial n=%d %s: %.3f ms %.3e samples/s" % (n, kw, ms, n / ms * 1e3), flush=True)
    del buf
