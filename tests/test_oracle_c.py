"""The C restatement (oracle/oracle_c.c) against the Python restatement and the golden vectors."""
import numpy as np
import pytest

import cases
from oracle.c_oracle import COracle
from oracle.em_read import em_read


@pytest.mark.parametrize("name", sorted(cases.TRACK_CASES))
def test_c_oracle_matches_golden_tracks(model_paths, golden, name):
    c = cases.TRACK_CASES[name]
    ow = c.get("overwrite", ())
    p = em_read(model_paths[c["model"]], isOverwriteZeroBoundaries=bool(ow), idxZeroBoundaries=ow or (1, 2, 3))
    co = COracle(p, prior=c.get("prior", 0), start=c.get("start"), uncor=c["uncor"], isQuantize500=c.get("q500", False),
                 layers=c.get("layers"))
    res = co.sample_tracks(c["n"], c["T"], c["seed"], c.get("first", 0), threads=2)
    assert res["rc"] == 0
    bins, vals = co.dense_compact(res)
    g = golden[name]
    assert np.array_equal(res["init_bins"], g["init_bins"])
    assert np.array_equal(res["init_values"], g["init_values"])
    assert np.array_equal(res["attempts"], g["attempts"])
    assert np.array_equal(bins, g["bins"])
    assert np.array_equal(vals, g["values"])      # both fp64: identical


def test_c_oracle_matches_golden_initial_and_terminal(model_paths, golden):
    c = cases.INITIAL_CASES["glider_initial_n256_seed12_first77"]
    co = COracle(em_read(model_paths[c["model"]]))
    r = co.sample_initial(c["n"], c["seed"], c["first"])
    g = golden["glider_initial_n256_seed12_first77"]
    assert np.array_equal(r["bins"], g["bins"]) and np.array_equal(r["values"], g["values"])
    for name, c in cases.TERMINAL_CASES.items():
        p = em_read(model_paths[c["model"]])
        lo, hi = np.full(p.n_initial, -np.inf), np.full(p.n_initial, np.inf)
        for lab in ('"own_speed"', '"int_speed"'):
            lo[p.labels_initial.index(lab)], hi[p.labels_initial.index(lab)] = cases.GENERIC_VEL
        r = COracle(p, start=c["start"]).sample_initial(c["n"], c["seed"], 0, box=(lo, hi))
        g = golden[name]
        assert np.array_equal(r["bins"], g["bins"]) and np.array_equal(r["values"], g["values"])
        assert np.array_equal(r["attempts"], g["attempts"])


def test_c_oracle_thread_count_invariance(model_paths):
    p = em_read(model_paths["uncor_allcode_fwsingle_v1"])
    co = COracle(p, uncor=True)
    a = co.sample_tracks(300, 120, 5, threads=1)
    b = co.sample_tracks(300, 120, 5, threads=4)
    for k in ("init_bins", "init_values", "samples", "sample_bins", "n_events"):
        assert np.array_equal(a[k], b[k])
