"""Fixtures for matlab/verify_parity.m -- the one-command check that closes "parity unpinned" for anyone with MATLAB.

For each case the Python ORACLE (oracle/, the restatement the CUDA path is tested against) is run with a recording
uniform provider; the uniforms are written in the order in which the *reference* consumes `rand`, together
with the outputs the oracle produced from them.  verify_parity.m replays the tape into the UNMODIFIED reference through
matlab/inject/rand.m and compares the reference's outputs with these files.  A match pins the oracle (and with it the CUDA
sampler, which is bit-identical to the oracle) to the real reference.

    python tests/golden/make_matlab_parity.py        (needs /root/reference/model; writes tests/golden/matlab_parity/)

Files per case <c>:  <c>_tape.txt (one uniform per line, %.17g), <c>_inits.txt (n x n_initial), <c>_events.txt (rows
[track dt var value], tracks 1-based), <c>_meta.txt (n T).  Cases: `uncor_fast` (uncor_1200code_v2p1, fast branch of
dbn_sample.m), `glider_slow` (glider_v1, slow branch), `terminal_geo` (terminal_v3_radar_encounter_model, sample.m:29-77
with the GENERIC speed limits), each under the keyed Philox stream, and `uncor_mt` = `mdl.sample(n, T, 'seed', 1)` under
MATLAB's own rng(1,'twister') emulated by MT19937 genrand_res53 (no tape: pins oracle/uniforms.py: MTStream).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.drivers import terminal_sample, uncor_sample  # noqa: E402
from oracle.em_read import em_read  # noqa: E402
from oracle.uniforms import KeyedPhilox, MTStream  # noqa: E402

OUT = os.path.join(HERE, "matlab_parity")
REF = "/root/reference/model"
CASES = {
    "uncor_fast": ("uncor_1200code_v2p1.txt", 6, 80, 1),
    "glider_slow": ("glider_v1.txt", 6, 80, 2),
    "terminal_geo": ("correlated_terminal/terminalradar/terminal_v3_radar_encounter_model.txt", 40, 0, 3),
    "uncor_mt": ("uncor_1200code_v2p1.txt", 4, 60, 1),
}


def build(case, model_dir=REF):
    """-> dict(tape, inits, events, meta) as numpy arrays (what the files hold)."""
    fname, n, T, seed = CASES[case]
    p = em_read(os.path.join(model_dir, fname))
    if case == "terminal_geo":
        U = KeyedPhilox(seed, record=True)
        inits, bins, att = terminal_sample(p, n, U)
        events = np.zeros((0, 4))
    else:
        U = MTStream(seed, record=True) if case == "uncor_mt" else KeyedPhilox(seed, record=True)
        out = uncor_sample(p, n, T, U)
        inits = np.stack([s.initial for s in out])
        events = np.concatenate([np.column_stack([np.full(s.events.shape[0], k + 1.0), s.events]) for k, s in enumerate(out)])
    tape = np.array([u for _, u in U.tape], dtype=np.float64)
    return dict(tape=tape, inits=inits, events=events, meta=np.array([[n, T, seed]], dtype=np.float64))


def main():
    os.makedirs(OUT, exist_ok=True)
    for case in CASES:
        d = build(case)
        for k, a in d.items():
            np.savetxt(os.path.join(OUT, "%s_%s.txt" % (case, k)), np.atleast_2d(a) if k != "tape" else a, fmt="%.17g")
        print(case, "tape", d["tape"].size, "events", d["events"].shape[0])


if __name__ == "__main__":
    main()
