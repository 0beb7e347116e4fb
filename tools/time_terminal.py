"""Times the terminal path on one GPU: geometry sampling (CorTerminalModel.sample) -> trajectory chains
(createEncounter) on the synthetic trajectory DBNs.  Usage: python tools/time_terminal.py [n] [tmax] [reps]"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from em_model_manned_bayes_b200.model import CorTerminalModel  # noqa: E402
from em_model_manned_bayes_b200.model_archive import materialize  # noqa: E402
from em_model_manned_bayes_b200.synthetic import write_terminal_model_set  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    tmax = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    d = tempfile.mkdtemp(prefix="emb_term_")
    geo_path = materialize(d, names=["terminal_v3_radar_encounter_model"])["terminal_v3_radar_encounter_model"]
    write_terminal_model_set(os.path.join(d, "traj"))
    m = CorTerminalModel(geo_path, parameters_directory=os.path.join(d, "traj"))
    vals, _, _ = m.sample_raw(n, seed=1, device="cuda:0")
    geo = vals.T.contiguous()
    res = m.create_encounters(geo, tmax, seed=2, device="cuda:0")
    torch.cuda.synchronize()
    steps = int(res.len.to(torch.int64).sum().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for r in range(reps):
        e0.record()
        m.create_encounters(geo, tmax, seed=3 + r, device="cuda:0", out=res)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    steps = int(res.len.to(torch.int64).sum().item())
    out_bytes = res.traj.numel() * 4 + res.len.numel() * 2
    print("terminal chains: n=%d tmax=%d  %.3f ms  %.3e trajectory states/s  (%.1f states/encounter, %.1f GB/s written)" %
          (n, tmax, best, steps / (best * 1e-3), steps / n, out_bytes / (best * 1e-3) / 1e9))
    e0.record()
    m.sample_raw(n, seed=9, device="cuda:0")
    e1.record()
    torch.cuda.synchronize()
    print("terminal geometry: n=%d %.3f ms %.3e encounters/s" % (n, e0.elapsed_time(e1), n / (e0.elapsed_time(e1) * 1e-3)))


if __name__ == "__main__":
    main()
