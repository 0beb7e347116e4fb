"""Generates the committed fixtures (run in the build container, where /root/reference exists):

  tests/golden/models.npz   -- the count tables etc. of the models the GPU tests / bench need, read
                               from /root/reference/model/**.txt with the ORACLE reader and stored in
                               a compact binary form.  /root/reference does not exist on the GPU box;
                               there the tests materialise these back into model/*.txt files with
                               em_model_manned_bayes_b200.em_write and load them through the product
                               reader.
  tests/golden/vectors.npz  -- golden outputs of the Python ORACLE (keyed-Philox provider) for small
                               seeded cases of every sampled path.

    python tests/golden/make_fixtures.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.drivers import dbn_tracks, initial_sample, terminal_sample, uncor_sample  # noqa: E402
from oracle.em_read import em_read  # noqa: E402
from oracle.uniforms import KeyedPhilox  # noqa: E402
from helpers import oracle_dense  # noqa: E402

REF = "/root/reference/model"
MODELS = {
    "balloon_v1": "balloon_v1.txt",
    "glider_v1": "glider_v1.txt",
    "paramotor_v1": "paramotor_v1.txt",
    "littoral_uncor_v1": "littoral_uncor_v1.txt",
    "cor_v1": "cor_v1.txt",
    "uncor_1200code_v2p1": "uncor_1200code_v2p1.txt",
    "uncor_allcode_fwsingle_v1": "uncor_allcode_fwsingle_v1.txt",
    "uncor_1200only_fwse_v1p2": "uncor_1200only_fwse_v1p2.txt",
    "terminal_v3_radar_encounter_model": "correlated_terminal/terminalradar/terminal_v3_radar_encounter_model.txt",
    # one model per remaining compiled kernel shape (GPU tests compare them with the C oracle; no golden vectors)
    "fai1_v1": "fai1_v1.txt",                      # slow branch, order dh', dv', dpsi'
    "uncor_1200code_v1": "uncor_1200code_v1.txt",  # slow branch, 6 initial variables
    "blimp_v1": "blimp_v1.txt",                    # fast branch, 3 gated variables
    "dueregard_v1": "dueregard_v1.txt",            # bins (5,9,7)
    "haa_v1": "haa_v1.txt",                        # bins (7,7,5), 9 initial variables, 7 gated
    "littoral_cor_v1": "littoral_cor_v1.txt",      # 4 dynamic variables, fast branch
    "weatherballoon_v1": "weatherballoon_v1.txt",  # 1 dynamic variable
}


def pack_model(p):
    d = {}
    d["labels_initial"] = np.array("\n".join(p.labels_initial))
    d["G_initial"] = p.G_initial.astype(np.uint8)
    d["r_initial"] = p.r_initial.astype(np.int32)
    flat = np.concatenate([x.ravel(order="F") for x in p.N_initial])
    assert np.all(flat == np.round(flat)) and flat.max() < 2 ** 32
    d["N_initial"] = flat.astype(np.uint32)
    if p.n_transition:
        d["labels_transition"] = np.array("\n".join(p.labels_transition))
        d["G_transition"] = p.G_transition.astype(np.uint8)
        d["r_transition"] = p.r_transition.astype(np.int32)
        flat = np.concatenate([x.ravel(order="F") for x in p.N_transition if x is not None])
        assert np.all(flat == np.round(flat)) and flat.max() < 2 ** 32
        d["N_transition"] = flat.astype(np.uint32)
    d["boundaries"] = np.concatenate([b for b in p.boundaries] + [np.zeros(0)])
    d["boundaries_len"] = np.array([len(b) for b in p.boundaries], dtype=np.int32)
    d["resample_rates"] = np.asarray(p.resample_rates, dtype=np.float64)
    return d


def tracks_case(p, out):
    bins, vals, dyn, tv = oracle_dense(p, out)
    # sparse event lists (out_events) of all tracks, concatenated: [dt, var, value] + the bin each value was drawn in
    ev = np.concatenate([np.asarray(s.events, dtype=np.float64).reshape(-1, 3) for s in out])
    evb = np.concatenate([np.asarray(s.event_bins, dtype=np.float64).ravel() for s in out])
    evo = np.concatenate([[0], np.cumsum([len(s.events) for s in out])]).astype(np.int64)
    return dict(bins=bins, values=vals, dyn=np.array(dyn), tv=np.array(tv), events=ev, event_bins=evb, event_offsets=evo,
                init_bins=np.stack([s.initial_bins for s in out]).astype(np.int8),
                init_values=np.stack([s.initial for s in out]),
                attempts=np.array([s.attempts for s in out], dtype=np.int32))


def main():
    models = {}
    parms = {}
    for name, rel in MODELS.items():
        p = em_read(os.path.join(REF, rel))
        parms[name] = p
        for k, v in pack_model(p).items():
            models[name + "/" + k] = v
    np.savez_compressed(os.path.join(HERE, "models.npz"), **models)
    if "--models-only" in sys.argv:
        print("models.npz", os.path.getsize(os.path.join(HERE, "models.npz")))
        return

    vec = {}

    def put(case, d):
        for k, v in d.items():
            vec[case + "/" + k] = np.asarray(v)

    # config 1 shape (small): UncorEncounterModel.sample on uncor_1200code_v2p1, seed 1
    p = parms["uncor_1200code_v2p1"]
    put("uncor_v2p1_n24_T300_seed1", tracks_case(p, uncor_sample(p, 24, 300, KeyedPhilox(1))))
    # config 3 shape: uncor_allcode_fwsingle_v1, T=600, 64-bit global sample index
    p = parms["uncor_allcode_fwsingle_v1"]
    put("uncor_fwsingle_n8_T600_seed2_first", tracks_case(p, uncor_sample(p, 8, 600, KeyedPhilox(2), first_sample=12345678901)))
    # default model of UncorEncounterModel (5 gated variables), quantize500
    p = parms["uncor_1200only_fwse_v1p2"]
    put("uncor_v1p2_n12_T100_seed3_q500", tracks_case(p, uncor_sample(p, 12, 100, KeyedPhilox(3), isQuantize500=True)))
    # layers (requires L as bin index): overwrite boundaries 1..3
    p2 = em_read(os.path.join(REF, MODELS["uncor_1200code_v2p1"]), isOverwriteZeroBoundaries=True, idxZeroBoundaries=(1, 2, 3))
    layers = np.array([[500, 1200], [1200, 3000], [3000, 5000], [5000, 12500]], dtype=np.float64)
    put("uncor_v2p1_layers_n12_T40_seed4", tracks_case(p2, uncor_sample(p2, 12, 40, KeyedPhilox(4), layers=layers, isQuantize500=True)))
    # start presets (RUN_uncor.m:20-50 uses start = {1,4,2,...})
    p = parms["uncor_1200code_v2p1"]
    st = [1, 4, 2, None, None, None, None]
    put("uncor_v2p1_start142_n12_T50_seed5", tracks_case(p, uncor_sample(p, 12, 50, KeyedPhilox(5), start=st)))
    # slow branch (dynamic -> dynamic edges): glider, paramotor (non-identity order_initial), littoral
    p = parms["glider_v1"]
    put("glider_n16_T120_seed6", tracks_case(p, uncor_sample(p, 16, 120, KeyedPhilox(6))))
    p = parms["paramotor_v1"]
    put("paramotor_n16_T75_seed7", tracks_case(p, uncor_sample(p, 16, 75, KeyedPhilox(7))))
    p = parms["littoral_uncor_v1"]
    put("littoral_uncor_n8_T33_seed8", tracks_case(p, uncor_sample(p, 8, 33, KeyedPhilox(8))))
    # config 4 stand-in: correlated model, 4 dynamic variables, slow branch, T = 60 (em_sample.m:26)
    p = parms["cor_v1"]
    put("cor_v1_n12_T60_seed9", tracks_case(p, dbn_tracks(p, 12, 60, KeyedPhilox(9))))
    # one dynamic variable, fast branch
    p = parms["balloon_v1"]
    put("balloon_n16_T50_seed10", tracks_case(p, dbn_tracks(p, 16, 50, KeyedPhilox(10))))
    # priors: dbe (fractional weights) on glider tracks
    p = parms["glider_v1"]
    put("glider_dbe_n8_T40_seed11", tracks_case(p, uncor_sample(p, 8, 40, KeyedPhilox(11), prior="dbe")))
    # config 2: initial network only
    S, V = initial_sample(parms["glider_v1"], 256, KeyedPhilox(12), first_sample=77)
    put("glider_initial_n256_seed12_first77", dict(bins=S.astype(np.int8), values=V))
    # config 5 geometry: terminal encounter model with GENERIC speed limits and a start preset
    p = parms["terminal_v3_radar_encounter_model"]
    I, B, A = terminal_sample(p, 64, KeyedPhilox(13))
    put("terminal_geo_n64_seed13", dict(values=I, bins=B.astype(np.int8), attempts=A.astype(np.int32)))
    st = [2, 1, 3] + [None] * 12
    I, B, A = terminal_sample(p, 32, KeyedPhilox(14), start=st)
    put("terminal_geo_start213_n32_seed14", dict(values=I, bins=B.astype(np.int8), attempts=A.astype(np.int32)))
    # config 5 trajectories: createEncounter.m chains on the synthetic trajectory DBNs
    # (em_model_manned_bayes_b200/synthetic.py -- the reference's 20 trajectory files are not in the checkout),
    # geometry = the golden terminal-geometry samples above
    import tempfile
    from em_model_manned_bayes_b200.synthetic import write_terminal_model_set
    from test_terminal_traj import geo_from_golden, oracle_propagate
    tp = write_terminal_model_set(tempfile.mkdtemp(prefix="traj_models_"))
    gold = {"terminal_geo_n64_seed13": {"values": vec["terminal_geo_n64_seed13/values"]}}
    geo = geo_from_golden(gold, 16)
    tr, ln = oracle_propagate(tp, geo, 31, 9000000000, 120)
    put("terminal_traj_n16_T120_seed31", dict(geo=geo, traj=tr, len=ln, first=np.int64(9000000000)))
    geo = geo_from_golden(gold, 12)
    geo[5] = np.clip(geo[5], 70.0, 180.0)            # own_speed inside the TEST limits (sample.m:64 would have rejected others)
    tr, ln = oracle_propagate(tp, geo, 32, 17, 45, ("TEST", "RTCA228_A1"))
    put("terminal_traj_n12_T45_seed32_test_a1", dict(geo=geo, traj=tr, len=ln, first=np.int64(17)))
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **vec)
    for f in ("models.npz", "vectors.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
