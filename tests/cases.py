"""Golden cases (tests/golden/vectors.npz, produced by the oracle): how to re-run each through the
product API.  Shared by the CPU emulation suite and the GPU suite so both exercise the same inputs."""
import numpy as np

LAYERS = np.array([[500, 1200], [1200, 3000], [3000, 5000], [5000, 12500]], dtype=np.float64)

# name -> dict(model, kind, n, T, seed, first, and options)
TRACK_CASES = {
    "uncor_v2p1_n24_T300_seed1": dict(model="uncor_1200code_v2p1", uncor=True, n=24, T=300, seed=1),
    "uncor_fwsingle_n8_T600_seed2_first": dict(model="uncor_allcode_fwsingle_v1", uncor=True, n=8, T=600, seed=2, first=12345678901),
    "uncor_v1p2_n12_T100_seed3_q500": dict(model="uncor_1200only_fwse_v1p2", uncor=True, n=12, T=100, seed=3, q500=True),
    "uncor_v2p1_layers_n12_T40_seed4": dict(model="uncor_1200code_v2p1", uncor=True, n=12, T=40, seed=4, q500=True, layers=LAYERS,
                                            overwrite=(1, 2, 3)),
    "uncor_v2p1_start142_n12_T50_seed5": dict(model="uncor_1200code_v2p1", uncor=True, n=12, T=50, seed=5,
                                              start=[1, 4, 2, None, None, None, None]),
    "glider_n16_T120_seed6": dict(model="glider_v1", uncor=True, n=16, T=120, seed=6),
    "paramotor_n16_T75_seed7": dict(model="paramotor_v1", uncor=True, n=16, T=75, seed=7),
    "littoral_uncor_n8_T33_seed8": dict(model="littoral_uncor_v1", uncor=True, n=8, T=33, seed=8),
    "cor_v1_n12_T60_seed9": dict(model="cor_v1", uncor=False, n=12, T=60, seed=9),
    "balloon_n16_T50_seed10": dict(model="balloon_v1", uncor=False, n=16, T=50, seed=10),
    "glider_dbe_n8_T40_seed11": dict(model="glider_v1", uncor=True, n=8, T=40, seed=11, prior="dbe"),
}

INITIAL_CASES = {
    "glider_initial_n256_seed12_first77": dict(model="glider_v1", n=256, seed=12, first=77),
}

TERMINAL_CASES = {
    "terminal_geo_n64_seed13": dict(model="terminal_v3_radar_encounter_model", n=64, seed=13, start=None),
    "terminal_geo_start213_n32_seed14": dict(model="terminal_v3_radar_encounter_model", n=32, seed=14,
                                             start=[2, 1, 3] + [None] * 12),
}

# GENERIC aircraft speed limits (@CorTerminalModel/getDynamicLimits.m:16-17)
GENERIC_VEL = (50.0, 506.0)


def label_index(labels, name):
    return labels.index(name) + 1 if name in labels else 0


def check_tracks(got, want, rtol=1e-6):
    """got: dict with bins/values/init_bins/init_values/attempts as numpy; want: golden case dict."""
    assert np.array_equal(np.asarray(got["init_bins"]), want["init_bins"]), "initial bins differ"
    assert np.array_equal(np.asarray(got["attempts"]).astype(np.int64), want["attempts"].astype(np.int64)), "attempts differ"
    assert np.array_equal(np.asarray(got["bins"]), want["bins"]), "track bins differ (must be bit-exact)"
    # initial values are fp64 on both sides and must be identical; dense values are fp32 (tolerance 1e-6 rel)
    assert np.array_equal(np.asarray(got["init_values"]), want["init_values"]), "initial continuous values differ"
    gv, wv = np.asarray(got["values"], dtype=np.float64), want["values"]
    assert gv.shape == wv.shape
    assert np.all(np.abs(gv - wv) <= rtol * np.abs(wv)), "continuous values differ by more than 1e-6 relative"


def check_events(ev, off, want, rtol=1e-6):
    """ev: structured emb_event rows, off: int64 [n+1]; want: golden case dict (events k x 3, event_bins, event_offsets).
    dt, var and bin must be identical row by row; values within rtol (fp32 on the device)."""
    ev, off = np.asarray(ev), np.asarray(off)
    assert np.array_equal(off, want["event_offsets"]), "rows per track differ"
    w = want["events"]
    assert np.array_equal(ev["dt"].astype(np.float64), w[:, 0]), "event dt differ"
    assert np.array_equal(ev["var"].astype(np.float64), w[:, 1]), "event variables differ"
    assert np.array_equal(ev["bin"].astype(np.float64), want["event_bins"]), "event bins differ"
    gv = ev["value"].astype(np.float64)
    assert np.all(np.abs(gv - w[:, 2]) <= rtol * np.abs(w[:, 2])), "event values differ by more than 1e-6 relative"
