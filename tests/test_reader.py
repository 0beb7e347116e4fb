"""Product reader/packer (libemb200.so, emb_model.cpp) against the oracle's independent restatement of
em_read.m, plus the exactness of the word-space threshold tables against the reference's fp64 rule."""
import os

import numpy as np
import pytest

from em_model_manned_bayes_b200 import _lib as L
from em_model_manned_bayes_b200.model import EncounterModel
from oracle import philox as px
from oracle import sampler as sp
from oracle.em_read import em_read

ALL = ["balloon_v1", "glider_v1", "paramotor_v1", "littoral_uncor_v1", "cor_v1", "uncor_1200code_v2p1",
       "uncor_allcode_fwsingle_v1", "uncor_1200only_fwse_v1p2", "terminal_v3_radar_encounter_model"]


@pytest.mark.parametrize("name", ALL)
def test_reader_fields_match_oracle(model_paths, name):
    m = EncounterModel(model_paths[name])
    p = em_read(model_paths[name])
    assert m.labels_initial == p.labels_initial and m.labels_transition == p.labels_transition
    assert m.n_initial == p.n_initial and m.n_transition == p.n_transition
    assert np.array_equal(m.G_initial, p.G_initial)
    assert np.array_equal(m.r_initial, p.r_initial)
    assert m.order_initial == p.order_initial
    for a, b in zip(m.N_initial, p.N_initial):
        assert np.array_equal(a, b)
    if p.n_transition:
        assert np.array_equal(m.G_transition, p.G_transition)
        assert np.array_equal(m.r_transition, p.r_transition)
        assert m.order_transition == p.order_transition
        assert np.array_equal(m.temporal_map, p.temporal_map)
        for a, b in zip(m.N_transition, p.N_transition):
            assert (a is None and b is None) or np.array_equal(a, b)
        dv = p.temporal_map[:, 1] - 1
        assert m.is_dynvar_depend == bool(p.G_transition[np.ix_(dv, dv)].any())
    for a, b in zip(m.boundaries, p.boundaries):
        assert np.array_equal(a, b)
    assert m.zero_bins == p.zero_bins
    assert np.array_equal(m.resample_rates, p.resample_rates)
    assert np.array_equal(m.bounds_initial, p.bounds_initial)
    for a, b in zip(m.cutpoints_initial, p.cutpoints_initial):
        assert np.array_equal(a, b)


def test_overwrite_zero_boundaries(model_paths):
    path = model_paths["uncor_1200code_v2p1"]
    m = EncounterModel(path, idxZeroBoundaries=(1, 2, 3), isOverwriteZeroBoundaries=True)
    p = em_read(path, isOverwriteZeroBoundaries=True, idxZeroBoundaries=(1, 2, 3))
    assert [len(b) for b in m.boundaries] == [len(b) for b in p.boundaries] == [0, 0, 0, 9, 6, 8, 8]
    assert np.array_equal(m.bounds_initial, p.bounds_initial)
    assert m.zero_bins == p.zero_bins


def _columns(model, which):
    """Yield (weights-free) packed columns: (var index, column index j, rp, packed uint32[rp])."""
    packed = model.packed(which)
    tabs = model.N_transition if which else model.N_initial
    off = 0
    for i, N in enumerate(tabs):
        if N is None:
            continue
        r, q = N.shape
        rp = (r + 3) & ~3
        yield i, N, packed[off: off + q * rp].reshape(q, rp)
        off += q * rp
    assert off == packed.size


def _bins_from_packed(col, ks):
    return int(col[-1]) + (ks[:, None] > col[None, :-1].astype(np.uint64)).sum(axis=1)


@pytest.mark.parametrize("name", ["balloon_v1", "glider_v1", "cor_v1", "uncor_1200code_v2p1", "terminal_v3_radar_encounter_model"])
def test_packed_thresholds_reproduce_fp64_rule(model_paths, name):
    """For sampled columns: bin from the packed word thresholds == select_random.m's
    `find(cumsum(w) >= sum(w)*u, 1)` for random words AND for the words adjacent to every threshold."""
    m = EncounterModel(model_paths[name])
    rng = np.random.default_rng(0)
    for which in (0, 1):
        if which and not m.n_transition:
            continue
        for i, N, cols in _columns(m, which):
            q = N.shape[1]
            js = np.unique(np.concatenate([[0, q - 1], rng.integers(0, q, size=min(q, 12))]))
            nz = np.nonzero(N.sum(axis=0) > 0)[0]
            if nz.size:
                js = np.unique(np.concatenate([js, nz[rng.integers(0, nz.size, size=min(nz.size, 12))]]))
            for j in js:
                col = cols[j]
                edge = col[:-1].astype(np.int64)
                edge = edge[edge != 0xFFFFFFFF]
                ks = np.concatenate([rng.integers(0, 2 ** 32, size=24), [0, 2 ** 32 - 1], edge, edge + 1, edge - 1])
                ks = np.unique(np.clip(ks, 0, 2 ** 32 - 1)).astype(np.uint64)
                got = _bins_from_packed(col, ks)
                want = np.array([sp.select_random_u(N[:, j], px.u01(int(k))) - 1 for k in ks])
                assert np.array_equal(got, want), (name, which, i, j)


def test_packed_thresholds_with_priors(model_paths):
    """dbe prior (fractional weights) and constant prior still give exact thresholds."""
    m = EncounterModel(model_paths["paramotor_v1"])
    rng = np.random.default_rng(1)
    for prior in ("dbe", 1, 0.25):
        m.prior = prior
        for which in (0, 1):
            for i, N, cols in _columns(m, which):
                r, q = N.shape
                alpha = np.full((r, q), 1.0 / (r * q)) if prior == "dbe" else np.full((r, q), float(prior))
                for j in rng.integers(0, q, size=min(q, 10)):
                    ks = rng.integers(0, 2 ** 32, size=32).astype(np.uint64)
                    got = _bins_from_packed(cols[j], ks)
                    want = np.array([sp.select_random_u(N[:, j] + alpha[:, j], px.u01(int(k))) - 1 for k in ks])
                    assert np.array_equal(got, want), (prior, which, i, j)


def test_reader_errors(tmp_path, model_paths):
    bad = tmp_path / "bad.txt"
    src = open(model_paths["balloon_v1"]).read()
    bad.write_text(src.replace("# resample_rates", "# resample_ratez"))
    with pytest.raises(L.EmbError, match="Unknown field: # resample_ratez") as ei:
        EncounterModel(str(bad))
    assert ei.value.code == L.EMB_E_PARSE
    cyc = tmp_path / "cyc.txt"
    cyc.write_text(src.replace("# G_initial\n0 1 \n0 0 ", "# G_initial\n0 1 \n1 0 "))
    with pytest.raises(L.EmbError, match="Network could not be hierarchically sorted"):
        EncounterModel(str(cyc))
    with pytest.raises(L.EmbError) as ei:
        EncounterModel(str(tmp_path / "missing.txt"))
    assert ei.value.code == L.EMB_E_IO
    short = tmp_path / "short.txt"
    short.write_text(src.replace("28313 5586", "5586"))
    with pytest.raises(L.EmbError, match="N_initial"):
        EncounterModel(str(short))


def test_crlf_and_blank_lines(tmp_path, model_paths):
    src = open(model_paths["balloon_v1"]).read()
    f = tmp_path / "crlf.txt"
    f.write_bytes(src.replace("\n", "\r\n\r\n").encode())
    a, b = EncounterModel(str(f)), EncounterModel(model_paths["balloon_v1"])
    assert np.array_equal(a.packed(0), b.packed(0)) and np.array_equal(a.packed(1), b.packed(1))


def test_uncor_get_dynamic_limits_matches_direct_conditioning(model_paths):
    """@UncorEncounterModel/getDynamicLimits.m: the strided-column arithmetic of :57-78 against the same
    conditional distributions taken from the count tables reshaped to their parent axes."""
    from em_model_manned_bayes_b200.model import UncorEncounterModel
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    N, r = m.N_initial, [int(x) for x in m.r_initial]
    Nv = N[3].reshape((r[3], r[0], r[1], r[2]), order="F")                       # v | G, A, L
    Ndh = N[5].reshape((r[5], r[0], r[1], r[2], r[3], r[4]), order="F")          # \dot h | G, A, L, v, \dot v
    combos = [(g, a, l, v) for g in range(1, r[0] + 1) for a in range(1, r[1] + 1) for l in range(1, r[2] + 1)
              for v in range(1, r[3] + 1) if Ndh[:, g - 1, a - 1, l - 1, v - 1, :].sum() > 0]
    assert (1, 4, 2, 4) in combos
    for (dG, dA, dL, dV) in [(1, 4, 2, 4), combos[0], combos[len(combos) // 2], combos[-1]]:
        init = [dG, dA, dL, dV, 1, 1, 1]
        got = m.getDynamicLimits(init, is_discretized=[True] * 7)
        v = Nv[:, dG - 1, dA - 1, dL - 1]
        dh = Ndh[:, dG - 1, dA - 1, dL - 1, dV - 1, :].sum(axis=1)

        def pct(w):
            cs = np.cumsum(100.0 * w / w.sum())
            return int(np.argmax(cs >= 1)) + 1, int(np.argmax(cs >= 99)) + 1
        klo, khi = pct(v)
        bV, bH = m.boundaries[3], m.boundaries[5]
        assert got["minVel_ft_s"] == max(bV[klo] * 1.68780972222222, 30.0)
        assert got["maxVel_ft_s"] == bV[khi] * 1.68780972222222
        klo, khi = pct(dh)
        assert got["maxVertRate_ft_s"] == max(abs(bH[klo]), abs(bH[khi])) / 60.0
    # continuous inputs: L and v come from the simulated track (results), G and A are bins already
    got = m.getDynamicLimits([1, 4, 1650.0, 120.0, 0, 0, 0], results=dict(up_ft=[1300.0, 2900.0], speed_ftps=[150.0, 260.0]))
    want = m.getDynamicLimits([1, 4, 2, 0, 0, 0, 0], is_discretized=[True, True, True, False, True, True, True],
                              results=dict(up_ft=[0.0], speed_ftps=[150.0, 260.0]))
    assert got == want and got["maxVel_ft_s"] > got["minVel_ft_s"] >= 30.0
    # a model without G/A in positions 1-2 falls back to the marginal tables (:85-86)
    g = UncorEncounterModel(model_paths["glider_v1"])
    lim = g.getDynamicLimits([1, 1, 1, 1, 1])
    assert lim["maxVel_ft_s"] > lim["minVel_ft_s"] and lim["maxVertRate_ft_s"] > 0


@pytest.mark.parametrize("name", ["uncor_1200code_v2p1", "uncor_allcode_fwsingle_v1", "uncor_1200only_fwse_v1p2", "glider_v1"])
def test_uncor_get_dynamic_limits_matches_the_oracle_restatement(model_paths, name):
    """Product (model.py, host-side table arithmetic) against oracle/dynlimits.py, the statement-by-statement restatement of
    @UncorEncounterModel/getDynamicLimits.m:1-130: every reachable (G, A, L, v) bin combination with discretised inputs, plus
    continuous L / v taken from a simulated track (`results`, :35-50), rotorcraft flag included."""
    from em_model_manned_bayes_b200.model import UncorEncounterModel, _find
    from oracle.dynlimits import get_dynamic_limits
    from oracle.em_read import em_read
    m = UncorEncounterModel(model_paths[name])
    p = em_read(model_paths[name])
    lab = p.labels_initial
    idx = dict(idx_G=_find(lab, '"G"'), idx_A=_find(lab, '"A"'), idx_L=_find(lab, '"L"'), idx_V=_find(lab, '"v"'),
               idx_DH=_find(lab, '"\\dot h"'))
    r = [int(x) for x in p.r_initial]
    disc = [True] * p.n_initial
    rs = np.random.RandomState(5)
    tried = 0
    for _ in range(400):
        init = [int(rs.randint(1, ri + 1)) for ri in r]
        try:
            want = get_dynamic_limits(p, init, is_discretized=disc, **idx)
        except (IndexError, ZeroDivisionError, ValueError, FloatingPointError):
            continue                                   # an unreachable combination (all-zero conditional table)
        if not np.isfinite(list(want.values())).all():
            continue
        got = m.getDynamicLimits(init, is_discretized=disc)
        assert got == want, (init, got, want)
        tried += 1
    assert tried >= 50
    if idx["idx_G"] == 1 and idx["idx_DH"] == 6:
        for rot in (False, True):
            m.isRotorcraft = rot
            init = [1, r[1], 1650.0, 120.0, 0, 0, 0][:p.n_initial]
            res = dict(up_ft=[700.0, 2900.0], speed_ftps=[90.0, 330.0])
            nd = [len(b) == 0 for b in p.boundaries]
            got = m.getDynamicLimits(init, results=res)
            want = get_dynamic_limits(p, init, results=res, is_discretized=nd, is_rotorcraft=rot, **idx)
            assert got == want
