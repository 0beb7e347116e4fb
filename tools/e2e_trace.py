"""Scratch: phase times of the e2e (events, host buffers) call.  EMB200_TRACE=1 python tools/e2e_trace.py"""
import ctypes as C, os, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200 import _lib as L
from em_model_manned_bayes_b200.model import UncorEncounterModel
from em_model_manned_bayes_b200.model_archive import materialize
p = materialize(tempfile.mkdtemp(prefix="emb_models_"), names=["uncor_allcode_fwsingle_v1"])["uncor_allcode_fwsingle_v1"]
m = UncorEncounterModel(p)
lib = L.lib()
n, T = 1250000, 600
probe = m.sample_events(n, T, seed=6, opts=m.uncor_opts(), device="cuda:0", want_init=False)
cap = int(probe.total * 1.03)
del probe
torch.cuda.empty_cache()
t0 = time.perf_counter()
h_ev = torch.empty(cap, dtype=torch.int64).pin_memory()
h_off = torch.empty(n + 1, dtype=torch.int64).pin_memory()
h_iv = torch.empty((m.n_initial, n), dtype=torch.float64).pin_memory()
print("pinned alloc %.1f ms" % ((time.perf_counter() - t0) * 1e3))
o = m.uncor_opts()
o.mem, o.device = L.EMB_MEM_HOST, 0
tot = C.c_int64(0)
init_only = L.TrackOut(None, None, None, h_iv.data_ptr(), None, None, None)
for k in range(4):
    rng = L.Rng(3000 + k, 0)
    t0 = time.perf_counter()
    L.check(lib.emb_sample_track_events(m._h, C.byref(rng), n, T, C.byref(o), cap, h_ev.data_ptr(), h_off.data_ptr(),
                                        C.byref(init_only), C.byref(tot)))
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms  %.3e track-timesteps/s  rows %d" % (k, dt * 1e3, n * T / dt, tot.value), flush=True)
# plain pinned D2H bandwidth for reference
d = torch.empty(1 << 28, dtype=torch.int64, device="cuda:0")
h = torch.empty(1 << 28, dtype=torch.int64).pin_memory()
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("pinned D2H 2 GiB: %.1f GB/s" % (2.147 / dt))
os.system("nvidia-smi topo -m 2>/dev/null | head -5; nproc; numactl -H 2>/dev/null | head -4")
