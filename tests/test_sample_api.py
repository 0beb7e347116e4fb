"""The reference-shaped drivers of the host mirror: UncorEncounterModel.sample (UncorEncounterModel.m:192-313, all four
outputs, including the EncounterModelEvents controls and their unit conversions) and CorTerminalModel.sample
(@CorTerminalModel/sample.m:29-77), against the oracle's restatement of the same drivers.

CPU part: the product's vectorised event expansion (model.expand_events / controls_of) against the oracle's line-by-line
events2samples / events2controls on oracle-generated event lists.  GPU part: the calls themselves."""
import numpy as np
import pytest

from em_model_manned_bayes_b200 import model as M
from oracle import sampler as sp
from oracle.drivers import terminal_sample, uncor_sample
from oracle.em_read import em_read
from oracle.uniforms import KeyedPhilox


def _oracle(model_paths, name, n, T, seed, **kw):
    p = em_read(model_paths[name])
    return p, uncor_sample(p, n, T, KeyedPhilox(seed), **kw)


@pytest.mark.parametrize("name,n,T", [("uncor_1200code_v2p1", 12, 90), ("glider_v1", 10, 70), ("uncor_1200only_fwse_v1p2", 8, 64)])
def test_vectorised_expansion_equals_the_reference_loops(model_paths, name, n, T):
    """expand_events / controls_of take the rows of all tracks at once; the oracle replays one list at a time exactly as
    events2samples.m:9-27 and events2controls.m:9-31 do.  Same event lists in, identical matrices out (fp64)."""
    p, out = _oracle(model_paths, name, n, T, 3)
    init = np.stack([s.initial for s in out])
    rows = np.concatenate([s.events for s in out])
    off = np.concatenate([[0], np.cumsum([s.events.shape[0] for s in out])])
    dense = M.expand_events(init, rows[:, 0], rows[:, 1], rows[:, 2], off, T)
    tm = np.asarray(p.temporal_map)
    ctl = M.controls_of(dense, rows[:, 0], off, tm[:, 0] - 1)
    assert len(ctl) == n
    for k, s in enumerate(out):
        assert np.array_equal(dense[k], sp.events2samples(s.initial, s.events))
        assert np.array_equal(ctl[k], sp.events2controls(s.initial, s.events, tm))
        assert np.array_equal(M.events2samples(s.initial, s.events), dense[k])
        assert np.array_equal(M.events2controls(s.initial, s.events, tm), ctl[k])


def test_expansion_edge_cases():
    """A track with only the closing row; simultaneous rows (dt = 0) on the same variable: the last one is the visible one;
    a row that takes effect exactly at T is invisible in the T columns but still gives a control row."""
    init = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    rows = np.array([[5, 0, 0],                                    # track 0: nothing happens
                     [2, 2, 20.0], [0, 2, 21.0], [0, 3, 30.0], [3, 1, 10.0], [0, 0, 0]], dtype=np.float64)   # track 1
    off = np.array([0, 1, 6])
    d = M.expand_events(init, rows[:, 0], rows[:, 1], rows[:, 2], off, 5)
    assert np.array_equal(d[0], np.repeat(init[0][:, None], 5, axis=1))
    assert np.array_equal(d[1], sp.events2samples(init[1], rows[1:]))
    assert np.array_equal(d[1][1], [5, 5, 21, 21, 21]) and np.array_equal(d[1][0], [4, 4, 4, 4, 4])
    c = M.controls_of(d, rows[:, 0], off, [0, 1])
    assert np.array_equal(c[0], [[0, 1, 2]])
    assert np.array_equal(c[1], sp.events2controls(init[1], rows[1:], np.array([[1, 4], [2, 5]])))


def test_encounter_model_events_class():
    """@EncounterModelEvents/EncounterModelEvents.m:34-49: the event matrix has four columns and at least one row."""
    e = M.EncounterModelEvents(event=[[0, 1, 2, 3], [5, 6, 7, 8]])
    assert np.array_equal(e.time_s, [0, 5]) and np.array_equal(e.longitudeAccel_ftpss, [3, 8])
    assert np.array_equal(M.EncounterModelEvents(event=np.zeros((0, 4))).event, [[0, 0, 0, 0]])
    assert np.array_equal(M.EncounterModelEvents().event, [[0, 0, 0, 0]])
    with pytest.raises(Exception):
        M.EncounterModelEvents(event=[[1, 2, 3]])


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,T,kw", [
    ("uncor_1200code_v2p1", 40, 300, {}),                                            # BASELINE configs[0] shape (RUN_uncor.m)
    ("uncor_allcode_fwsingle_v1", 24, 210, dict(isQuantize500=True)),
    ("uncor_1200only_fwse_v1p2", 16, 120, {}),
    ("glider_v1", 24, 150, {}),                                                      # slow branch of dbn_sample.m
])
def test_uncor_sample_returns_the_references_four_outputs(model_paths, name, n, T, kw):
    """[out_inits, out_events, out_samples, out_EME] = mdl.sample(n, T, 'seed', s) (UncorEncounterModel.m:192): inits
    identical (fp64), event rows dt/var identical and values within 1e-6 (fp32 rows), the expanded n_initial x T matrices
    and the EncounterModelEvents matrices [t, dh/60, deg2rad(dpsi), dv*1.68780972222222] (:291-297) within 1e-6."""
    seed = 11
    p, want = _oracle(model_paths, name, n, T, seed, **kw)
    m = M.UncorEncounterModel(model_paths[name])
    out_inits, out_events, out_samples, out_EME = m.sample(n, T, seed=seed, **kw)
    assert out_inits.shape == (n, p.n_initial) and len(out_events) == len(out_samples) == len(out_EME) == n
    close = lambda a, b: np.all(np.abs(a - b) <= 1e-6 * np.abs(b))
    for k, s in enumerate(want):
        assert np.array_equal(out_inits[k], s.initial)
        assert out_events[k].shape == s.events.shape
        assert np.array_equal(out_events[k][:, :2], s.events[:, :2]) and close(out_events[k][:, 2], s.events[:, 2])
        assert out_samples[k].shape == (p.n_initial, T) and close(out_samples[k], s.samples)
        assert isinstance(out_EME[k], M.EncounterModelEvents)
        got = out_EME[k].event
        assert got.shape == s.controls.shape and np.array_equal(got[:, 0], s.controls[:, 0]) and close(got, s.controls)


@pytest.mark.gpu
def test_uncor_sample_layers_option(model_paths):
    """'layers' (UncorEncounterModel.m:259-263) needs L as a bin index: isOverwriteZeroBoundaries = true."""
    from oracle.em_read import em_read as rd
    path = model_paths["uncor_1200code_v2p1"]
    p = rd(path, isOverwriteZeroBoundaries=True, idxZeroBoundaries=[1, 2, 3])
    layers = np.array([[500.0, 1200.0], [1200.0, 3000.0], [3000.0, 5000.0], [5000.0, 18000.0]])
    want = uncor_sample(p, 20, 60, KeyedPhilox(4), layers=layers)
    m = M.UncorEncounterModel(path, isOverwriteZeroBoundaries=True)
    out_inits, out_events, out_samples, out_EME = m.sample(20, 60, seed=4, layers=layers)
    for k, s in enumerate(want):
        assert np.array_equal(out_inits[k], s.initial)
        assert np.all(np.abs(out_EME[k].event - s.controls) <= 1e-6 * np.abs(s.controls))


@pytest.mark.gpu
def test_terminal_sample_returns_the_references_two_outputs(model_paths):
    """[outInits, outSamples] = mdl.sample(n, 'seed', s) (@CorTerminalModel/sample.m:29-77): outInits identical (fp64),
    outSamples one struct per sample with the unquoted labels as field names (:58-61)."""
    name = "terminal_v3_radar_encounter_model"
    p = em_read(model_paths[name])
    n, seed = 64, 9
    want, want_bins, att = terminal_sample(p, n, KeyedPhilox(seed))
    m = M.CorTerminalModel(model_paths[name])
    out_inits, out_samples = m.sample(n, seed=seed)
    assert np.array_equal(out_inits, want)
    names = [l.replace('"', "") for l in p.labels_initial]
    assert len(out_samples) == n and list(out_samples[0].keys()) == names
    for k in range(n):
        assert [out_samples[k][f] for f in names] == list(want[k])


@pytest.mark.gpu
@pytest.mark.parametrize("name,prior", [("uncor_1200code_v2p1", 0), ("glider_v1", "dbe"), ("cor_v1", 0), ("balloon_v1", 2.5)])
def test_model_from_arrays_equals_the_file_model(model_paths, name, prior):
    """The call sequence of matlab/emb_handle.m -- the object's own arrays (G, r, N + dirichlet, temporal_map, boundaries,
    resample_rates; EncounterModel.m:5-70) through emb_model_from_arrays -- must give the same tracks, bit for bit, as the
    file-loaded model with the same prior set through emb_set_prior."""
    ref = M.EncounterModel(model_paths[name])
    if prior != 0:
        ref.prior = prior
    Ni, Nt = ref.N_initial, ref.N_transition
    alpha_i, alpha_t = sp.bn_dirichlet_prior(Ni, prior), sp.bn_dirichlet_prior(Nt, prior)
    m = M.EncounterModel.from_arrays(ref.G_initial, ref.r_initial, Ni, ref.G_transition, ref.r_transition, Nt,
                                     ref.temporal_map, ref.boundaries, ref.resample_rates, alpha_i, alpha_t)
    assert m.n_initial == ref.n_initial and m.order_initial == ref.order_initial
    assert np.array_equal(m.temporal_map, ref.temporal_map) and m.zero_bins == ref.zero_bins
    a = ref.sample_tracks(500, 120, seed=21, first_sample=9)
    b = m.sample_tracks(500, 120, seed=21, first_sample=9)
    assert np.array_equal(a.bins, b.bins) and np.array_equal(a.values, b.values)
    assert np.array_equal(a.init_bins, b.init_bins) and np.array_equal(a.init_values, b.init_values)
    ea, eb = ref.sample_events(100, 120, seed=22), m.sample_events(100, 120, seed=22)
    assert np.array_equal(ea.events, eb.events) and np.array_equal(ea.offsets, eb.offsets)


@pytest.mark.gpu
def test_stay_prior_through_the_weight_tables_equals_emb_set_prior(tmp_path):
    """createEncounter_b200.m hands setTransitionPriors' alpha to the library inside the weights; the product's own
    EMB_PRIOR_STAY must select the same thresholds (createEncounter.m:128-129)."""
    from em_model_manned_bayes_b200 import synthetic
    paths = synthetic.write_terminal_model_set(str(tmp_path), seed=3)
    path = sorted(paths.values())[0] if isinstance(paths, dict) else sorted(paths)[0]
    ref = M.EncounterModel(path)
    ref.set_transition_stay_prior(1.0)
    alpha_t = sp.set_transition_priors(ref.G_transition, ref.r_transition, ref.temporal_map, 1)
    m = M.EncounterModel.from_arrays(ref.G_initial, ref.r_initial, ref.N_initial, ref.G_transition, ref.r_transition,
                                     ref.N_transition, ref.temporal_map, ref.boundaries, ref.resample_rates, None, alpha_t)
    a, b = ref.sample_tracks(300, 40, seed=5), m.sample_tracks(300, 40, seed=5)
    assert np.array_equal(a.bins, b.bins) and np.array_equal(a.values, b.values)


@pytest.mark.parametrize("name,prior", [("uncor_1200code_v2p1", 0), ("glider_v1", "dbe"), ("cor_v1", 0), ("balloon_v1", 2.5),
                                        ("terminal_v3_radar_encounter_model", 0)])
def test_model_from_arrays_packs_the_same_thresholds_as_the_file_model(model_paths, name, prior):
    """No GPU needed: the word-space threshold tables uploaded to the device (emb_model_get_packed) are identical for the
    file-loaded model with emb_set_prior and for the array-built model whose weights already contain N + alpha."""
    ref = M.EncounterModel(model_paths[name])
    if prior != 0:
        ref.prior = prior
    Ni, Nt = ref.N_initial, ref.N_transition
    m = M.EncounterModel.from_arrays(ref.G_initial, ref.r_initial, Ni, ref.G_transition, ref.r_transition, Nt,
                                     ref.temporal_map, ref.boundaries, ref.resample_rates,
                                     sp.bn_dirichlet_prior(Ni, prior), sp.bn_dirichlet_prior(Nt, prior))
    assert np.array_equal(m.packed(0), ref.packed(0))
    if ref.n_transition:
        assert np.array_equal(m.packed(1), ref.packed(1))
    assert m.order_initial == ref.order_initial and m.order_transition == ref.order_transition
    assert m.zero_bins == ref.zero_bins and np.array_equal(m.resample_rates, ref.resample_rates)
    assert all(np.array_equal(a, b) for a, b in zip(m.boundaries, ref.boundaries))


# ---- @CorTerminalModel/InitStartTerminal.m and per-sample presets (RUN_terminal.m:28-44) ---------------------------------------
def test_init_start_terminal_matches_the_oracle_restatement(model_paths):
    """Combination order (class slowest, intruder intent fastest), n_enc_per_comb = ceil(n / n_combs), the row count of the
    grown cell, and the n_combs > nSamples case (InitStartTerminal.m:47-56)."""
    from oracle.drivers import init_start_terminal
    p = em_read(model_paths["terminal_v3_radar_encounter_model"])
    m = M.CorTerminalModel(model_paths["terminal_v3_radar_encounter_model"])
    for kw in (dict(nSamples=18), dict(nSamples=40), dict(nSamples=5),
               dict(nSamples=25, airspace_class=(True, False, False, True), own_intent=(False, True), int_intent=(True, False, True))):
        got = m.InitStartTerminal(**kw)
        okw = dict(kw)
        okw["n_samples"] = okw.pop("nSamples")
        want = init_start_terminal(p, **okw)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert [None if v is None else int(v) for v in a] == [None if (isinstance(v, list) and not v) else int(v) for v in b]
    rows = m.InitStartTerminal(nSamples=18)
    assert len(rows) == 18 and rows[0][:3] == [2, 1, 1] and rows[1][:3] == [2, 1, 2] and rows[17][:3] == [4, 2, 3]   # RUN_terminal.m:29-32
    with pytest.raises(Exception):
        m.InitStartTerminal(nSamples=10, own_intent=(True,))


def test_per_sample_presets_on_the_host_emulation(model_paths):
    """start_per_sample through the device code (host emulation): sample i uses row i of the InitStartTerminal cell -- equal
    to the oracle's bn_sample called with that row as `start` at global index i; a row that presets a variable whose parent
    is free is refused like bn_sample.m:46-47."""
    import helpers as H
    from em_model_manned_bayes_b200 import _lib as L
    from oracle.drivers import init_start_terminal, initial_sample
    path = model_paths["terminal_v3_radar_encounter_model"]
    p = em_read(path)
    rows = init_start_terminal(p, 18)
    em = H.EmuModel(path)
    sps = np.ascontiguousarray(np.array([[0 if (isinstance(v, list) and not v) else int(v) for v in r] for r in rows], dtype=np.int8).T)
    o = H.EmuModel.opts(p.n_initial)
    o.start_per_sample = sps.ctypes.data
    bins, vals, att = em.sample_initial(p.n_initial, 18, 4, 100, o)
    for i, r in enumerate(rows):
        S, V = initial_sample(p, 1, KeyedPhilox(4), first_sample=100 + i, start=r)
        assert np.array_equal(bins[i], S[0]) and np.array_equal(vals[i], V[0])
    bad = sps.copy()
    bad[0, 3] = 0                      # own_intent preset, its parent airspace_class free
    o.start_per_sample = bad.ctypes.data
    with pytest.raises(L.EmbError) as ei:
        em.sample_initial(p.n_initial, 18, 4, 100, o)
    assert ei.value.code == L.EMB_E_ARG


@pytest.mark.gpu
def test_run_terminal_start_combinations(model_paths):
    """RUN_terminal.m:28-44: the 18 start rows of InitStartTerminal, each used as mdl.start for mdl.sample(500, 'seed', 1)
    (the reference's loop), and all of them in ONE batched call with per-sample presets -- both against the oracle."""
    from oracle.drivers import init_start_terminal
    name = "terminal_v3_radar_encounter_model"
    p = em_read(model_paths[name])
    m = M.CorTerminalModel(model_paths[name])
    m.acType1 = "RTCA228_A1"                                                             # RUN_terminal.m:25
    starts = m.InitStartTerminal(nSamples=18)
    want_rows = init_start_terminal(p, 18)
    lim1 = M.CorTerminalModel.DYN_LIMITS["RTCA228_A1"]
    for ii in (0, 7, 17):
        m.start = starts[ii]
        out_inits, out_samples = m.sample(60, seed=1)
        want, _, _ = terminal_sample(p, 60, KeyedPhilox(1), start=want_rows[ii], dyn_limits1=lim1)
        assert np.array_equal(out_inits, want)
        assert np.all(out_inits[:, 0] == starts[ii][0]) and np.all(out_inits[:, 1] == starts[ii][1])
    m.start = [None] * m.n_initial
    rows = [starts[k % 18] for k in range(90)]
    for device in (None, "cuda:0"):
        vals, bins, att = m.sample_raw(90, seed=3, first_sample=11, device=device, start_per_sample=rows)
        vals = vals.cpu().numpy() if device else vals
        for k in (0, 5, 17, 18, 44, 89):
            want, _, _ = terminal_sample(p, 1, KeyedPhilox(3), first_sample=11 + k, start=want_rows[k % 18], dyn_limits1=lim1)
            assert np.array_equal(vals[k], want[0])


@pytest.mark.gpu
def test_per_sample_presets_on_tracks_and_bad_rows(model_paths):
    """emb_sample_tracks / emb_sample_track_events with start_per_sample: track i starts from row i (compared with one call
    per distinct row using the uniform `start`), and an invalid row is reported as EMB_E_ARG."""
    from em_model_manned_bayes_b200 import _lib as L
    m = M.UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    rows = [[1, 4, 2, None, None, None, None], [None] * 7, [2, 1, 1, 3, None, None, None]]
    n, T = 30, 50
    sps = [rows[k % 3] for k in range(n)]
    o = m.uncor_opts()
    o2 = m._opts(start_per_sample=np.array([[np.nan if v is None else v for v in r] for r in sps]))
    for f in ("reject_mode", "idx_v", "idx_dh", "idx_L"):
        setattr(o2, f, getattr(o, f))
    got = m.sample_tracks(n, T, seed=8, first_sample=5, opts=o2)
    for j, r in enumerate(rows):
        ref = m.sample_tracks(n, T, seed=8, first_sample=5, opts=m.uncor_opts(start=r))
        pick = np.arange(j, n, 3)
        assert np.array_equal(got.bins[pick], ref.bins[pick]) and np.array_equal(got.values[pick], ref.values[pick])
        assert np.array_equal(got.init_values[:, pick], ref.init_values[:, pick])
    bad = m._opts(start_per_sample=np.array([[np.nan, np.nan, np.nan, 2, np.nan, np.nan, np.nan]] * n))   # v preset, parents free
    with pytest.raises(L.EmbError) as ei:
        m.sample_tracks(n, T, seed=8, opts=bad)
    assert ei.value.code == L.EMB_E_ARG


@pytest.mark.gpu
def test_correct_dbn_option_on_the_gpu(model_paths):
    """emb_sample_opts::correct_dbn on the device (specialised per-step kernel for the 7-variable shape): dense tracks and
    event lists against the oracle run with the same option; default stays the reference's frozen-parent behaviour."""
    import helpers as H
    from em_model_manned_bayes_b200 import _lib as L
    p = em_read(model_paths["uncor_1200code_v2p1"])
    n, T = 48, 150
    want = uncor_sample(p, n, T, KeyedPhilox(6), correct_dbn=True)
    bins, vals, dyn, tv = H.oracle_dense(p, want)
    m = M.UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    got = m.sample_tracks(n, T, seed=6, opts=m.uncor_opts(correct_dbn=True), device="cuda:0")
    assert L.lib().emb_debug_last_kernel_fast() == 1
    assert np.array_equal(got.bins.cpu().numpy(), bins)
    assert np.all(np.abs(got.values.cpu().numpy().astype(np.float64) - vals) <= 1e-6 * np.abs(vals))
    ev = m.sample_events(n, T, seed=6, opts=m.uncor_opts(correct_dbn=True))
    for k, s in enumerate(want):
        assert np.array_equal(ev.track(k)[:, :2], s.events[:, :2])
    ref = m.sample_tracks(n, T, seed=6, opts=m.uncor_opts(), device="cuda:0")
    assert not np.array_equal(ref.bins.cpu().numpy(), bins)
