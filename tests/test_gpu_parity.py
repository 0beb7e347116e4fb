"""GPU suite (`-m gpu`, run on the B200): libemb200.so through the C ABI against the oracle's golden
vectors, against the oracle run live on small seeded inputs, and -- at BASELINE.json's full sizes --
through size-independent properties (shard invariance, determinism, histogram == dense counts,
chi-square of node marginals against the normalised count tables)."""
import numpy as np
import pytest

import cases
from em_model_manned_bayes_b200 import _lib as L
from em_model_manned_bayes_b200.model import CorTerminalModel, EncounterModel, UncorEncounterModel
from oracle.drivers import uncor_sample
from oracle.em_read import em_read
from oracle.uniforms import KeyedPhilox
from helpers import oracle_dense

pytestmark = pytest.mark.gpu


def _model(model_paths, c):
    path = model_paths[c["model"]]
    ow = c.get("overwrite", ())
    cls = UncorEncounterModel if c["uncor"] else EncounterModel
    m = cls(path, idxZeroBoundaries=ow or (1, 2, 3), isOverwriteZeroBoundaries=bool(ow))
    if c.get("prior") is not None:
        m.prior = c["prior"]
    return m


def _run_tracks(model_paths, c, device=None):
    m = _model(model_paths, c)
    if c["uncor"]:
        o = m.uncor_opts(isQuantize500=c.get("q500", False), layers=c.get("layers"), start=c.get("start"))
    else:
        o = m._opts(start=c.get("start"))
    res = m.sample_tracks(c["n"], c["T"], seed=c["seed"], first_sample=c.get("first", 0), opts=o, device=device)
    cpu = (lambda a: a.cpu().numpy()) if device is not None else (lambda a: a)
    return dict(bins=cpu(res.bins), values=cpu(res.values), init_bins=cpu(res.init_bins).T,
                init_values=cpu(res.init_values).T, attempts=cpu(res.attempts))


@pytest.mark.parametrize("generic", [0, 1], ids=["specialised", "generic"])
@pytest.mark.parametrize("name", sorted(cases.TRACK_CASES))
def test_tracks_match_golden(model_paths, golden, name, generic):
    """Both kernels (k_tracks_fast<...> and k_tracks_generic) against the oracle's golden vectors."""
    lib = L.lib()
    lib.emb_debug_force_generic(generic)
    try:
        got = _run_tracks(model_paths, cases.TRACK_CASES[name])
        was_fast = lib.emb_debug_last_kernel_fast()
    finally:
        lib.emb_debug_force_generic(0)
    cases.check_tracks(got, golden[name])
    assert was_fast == (0 if generic else 1)


def _opts_of(m, c):
    if c["uncor"]:
        return m.uncor_opts(isQuantize500=c.get("q500", False), layers=c.get("layers"), start=c.get("start"))
    return m._opts(start=c.get("start"))


@pytest.mark.parametrize("generic", [0, 1], ids=["specialised", "generic"])
@pytest.mark.parametrize("name", sorted(cases.TRACK_CASES))
def test_event_lists_match_golden(model_paths, golden, name, generic):
    """emb_sample_track_events (count pass, device prefix sum, write pass) against the oracle's out_events:
    same rows in the same order, dt / variable / bin identical, values within 1e-6."""
    lib = L.lib()
    c = cases.TRACK_CASES[name]
    m = _model(model_paths, c)
    lib.emb_debug_force_generic(generic)
    try:
        res = m.sample_events(c["n"], c["T"], seed=c["seed"], first_sample=c.get("first", 0), opts=_opts_of(m, c))
        assert lib.emb_debug_last_kernel_fast() == (0 if generic else 1)
    finally:
        lib.emb_debug_force_generic(0)
    cases.check_events(res.events, res.offsets, golden[name])
    assert np.array_equal(res.init_values.T, golden[name]["init_values"])
    assert np.array_equal(res.attempts.astype(np.int64), golden[name]["attempts"].astype(np.int64))


@pytest.mark.parametrize("name", sorted(cases.TRACK_CASES))
def test_packed_event_rows_match_golden(model_paths, golden, name):
    """emb_sample_track_events_packed: the 5-byte rows (what crosses PCIe) decode to the oracle's out_events -- dt, variable
    and bin identical, and the value, evaluated on the host in fp64 from (bin, 23-bit fraction) exactly as dediscretize.m:39
    does, equal to the oracle's fp64 value to 1e-12 (the 8-byte rows carry it rounded to fp32)."""
    c = cases.TRACK_CASES[name]
    m = _model(model_paths, c)
    res = m.sample_events_packed(c["n"], c["T"], seed=c["seed"], first_sample=c.get("first", 0), opts=_opts_of(m, c))
    legacy = m.sample_events(c["n"], c["T"], seed=c["seed"], first_sample=c.get("first", 0), opts=_opts_of(m, c))
    dt, var, b, val = res.decode()
    rows = np.asarray(legacy.events)
    assert np.array_equal(res.offsets, legacy.offsets) and res.total == legacy.total
    assert np.array_equal(dt, rows["dt"]) and np.array_equal(var, rows["var"]) and np.array_equal(b, rows["bin"])
    assert np.all(np.abs(val - rows["value"].astype(np.float64)) <= 1e-6 * np.abs(val))
    g = golden[name]
    want = g["events"] if "events" in g else None
    if want is not None:          # golden rows: [dt, var, value] per track, concatenated
        assert np.all(np.abs(val - want[:, 2]) <= 1e-12 * np.abs(want[:, 2]))
    assert np.array_equal(res.init_values.T, g["init_values"])


def test_packed_event_rows_on_device_and_limits(model_paths):
    """Device buffers give the same packed rows as host buffers (pipelined path); T > 1023 is refused by the packed entry
    point (10-bit dt) while the 8-byte entry point still works (it switches the device rows to 2-byte dts)."""
    m = UncorEncounterModel(model_paths["uncor_allcode_fwsingle_v1"])
    n, T = 20_000, 600
    h = m.sample_events_packed(n, T, seed=41, first_sample=7, opts=m.uncor_opts())
    d = m.sample_events_packed(n, T, seed=41, first_sample=7, opts=m.uncor_opts(), device="cuda:0")
    assert h.total == d.total and np.array_equal(h.offsets, d.offsets.cpu().numpy())
    assert np.array_equal(h.words, d.words.cpu().numpy().view(np.uint32)) and np.array_equal(h.dts, d.dts.cpu().numpy())
    with pytest.raises(L.EmbError) as ei:
        m.sample_events_packed(50, 1500, seed=1, opts=m.uncor_opts())
    assert ei.value.code == L.EMB_E_LIMIT
    long_ = m.sample_events(50, 1500, seed=1, opts=m.uncor_opts())
    off = np.asarray(long_.offsets)
    assert np.array_equal(np.add.reduceat(np.asarray(long_.events)["dt"].astype(np.int64), off[:-1]), np.full(50, 1500))


def test_event_buffer_too_small_is_reported_and_retried(model_paths):
    c = cases.TRACK_CASES["uncor_v2p1_n24_T300_seed1"]
    m = _model(model_paths, c)
    a = m.sample_events(c["n"], c["T"], seed=c["seed"], opts=_opts_of(m, c), capacity=10)      # forces the retry
    b = m.sample_events(c["n"], c["T"], seed=c["seed"], opts=_opts_of(m, c))
    assert a.total == b.total and np.array_equal(a.events, b.events) and np.array_equal(a.offsets, b.offsets)


def test_events_expand_to_the_dense_output_at_scale(model_paths):
    """200k tracks x 600 s on the device: events2samples(event list) == dense tiles, bit for bit (same fp32 values),
    and both kernels give the same rows."""
    import torch
    from em_model_manned_bayes_b200.model import events2samples
    lib = L.lib()
    n, T = 200_000, 600
    m = UncorEncounterModel(model_paths["uncor_allcode_fwsingle_v1"])
    ev = m.sample_events_uncor(n, T, seed=5, device="cuda:0")
    dense = m.sample_compact(n, T, seed=5, device="cuda:0")
    lib.emb_debug_force_generic(1)
    try:
        evg = m.sample_events_uncor(n, T, seed=5, device="cuda:0")
    finally:
        lib.emb_debug_force_generic(0)
    assert torch.equal(ev.offsets, evg.offsets)
    a = ev.events.cpu().numpy().view(L.EVENT_DTYPE)
    g = evg.events.cpu().numpy().view(L.EVENT_DTYPE)
    for f in ("dt", "var", "bin"):
        assert np.array_equal(a[f], g[f])
    assert np.all(np.abs(a["value"].astype(np.float64) - g["value"]) <= 1e-6 * np.abs(g["value"].astype(np.float64)))
    off = ev.offsets.cpu().numpy()
    assert off[-1] == ev.total and np.all(np.diff(off) >= 1)
    # every track: sum(dt) == T, closing row has var 0
    last = a[off[1:] - 1]
    assert np.all(last["var"] == 0)
    assert np.array_equal(np.add.reduceat(a["dt"].astype(np.int64), off[:-1]), np.full(n, T))
    iv = dense.init_values.cpu().numpy().T
    tv0 = [v - 1 for v in dense.tv_vars]
    vals = dense.values[:64].cpu().numpy()
    for k in range(64):
        rows = a[off[k]:off[k + 1]]
        ev3 = np.stack([rows["dt"].astype(np.float64), rows["var"].astype(np.float64), rows["value"].astype(np.float64)], 1)
        init32 = iv[k].copy()
        init32[tv0] = vals[k][:, 0]                       # dense values are fp32
        d = events2samples(init32, ev3)
        assert np.array_equal(d[tv0, :].astype(np.float32), vals[k])


def test_specialised_equals_generic_at_scale(model_paths):
    """100k tracks x 600 s: the two kernels must agree bit-for-bit on every bin and, because the
    specialised kernel de-discretises in fp32 and the generic one in fp64, within 1e-6 relative on values."""
    lib = L.lib()
    m = UncorEncounterModel(model_paths["uncor_allcode_fwsingle_v1"])
    a = m.sample_compact(100_000, 600, seed=99, device="cuda:0")
    lib.emb_debug_force_generic(1)
    try:
        b = m.sample_compact(100_000, 600, seed=99, device="cuda:0")
    finally:
        lib.emb_debug_force_generic(0)
    import torch
    assert torch.equal(a.bins_tiled, b.bins_tiled)
    av, bv = a.values_tiled.double(), b.values_tiled.double()
    assert bool(((av - bv).abs() <= 1e-6 * bv.abs()).all())
    assert bool(((av == 0) == (bv == 0)).all())
    assert torch.equal(a.init_values, b.init_values)


@pytest.mark.parametrize("model,n", [("glider_v1", 1_000_003), ("uncor_allcode_fwsingle_v1", 400_000),
                                     ("terminal_v3_radar_encounter_model", 100_000)])
def test_initial_specialised_equals_generic_at_scale(model_paths, model, n):
    """k_initial_fast (4 samples per thread) against k_initial on every output byte, ragged n included."""
    import torch
    lib = L.lib()
    m = EncounterModel(model_paths[model])
    a = m.sample_initial(n, seed=17, first_sample=3, device="cuda:0")
    assert lib.emb_debug_last_kernel_fast() == 1
    lib.emb_debug_force_generic(1)
    try:
        b = m.sample_initial(n, seed=17, first_sample=3, device="cuda:0")
        assert lib.emb_debug_last_kernel_fast() == 0
    finally:
        lib.emb_debug_force_generic(0)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_initial_fp32_values_match_fp64(model_paths):
    """emb_sample_initial_f32 (the compact 1 + 4 bytes per variable of SURVEY 8d): same bins, values within 1e-6 relative of
    the fp64 values, specialised and generic kernel, aligned and unaligned first_sample."""
    lib = L.lib()
    for name in ("glider_v1", "uncor_1200code_v2p1", "terminal_v3_radar_encounter_model"):
        m = EncounterModel(model_paths[name])
        for first in (0, 7):
            b64, v64, _ = m.sample_initial(100_003, seed=3, first_sample=first, device="cuda:0")
            b32, v32, _ = m.sample_initial(100_003, seed=3, first_sample=first, device="cuda:0", values_fp32=True)
            assert lib.emb_debug_last_kernel_fast() == 1
            assert b32.equal(b64)
            a, b = v32.double().cpu().numpy(), v64.cpu().numpy()
            assert np.all(np.abs(a - b) <= 1e-6 * np.abs(b))
            lib.emb_debug_force_generic(1)
            try:
                bg, vg, _ = m.sample_initial(100_003, seed=3, first_sample=first, device="cuda:0", values_fp32=True)
            finally:
                lib.emb_debug_force_generic(0)
            assert bg.equal(b64) and np.all(np.abs(vg.double().cpu().numpy() - b) <= 1e-6 * np.abs(b))


@pytest.mark.parametrize("n_devices", [1, 2, 0], ids=["one", "two", "all"])
def test_in_library_multi_gpu_equals_single_device(model_paths, n_devices):
    """emb_sample_tracks_multi: one process, one host thread per device, shards by global sample index, ONE collective
    (ncclAllReduce of the verification histograms inside the library).  The shards concatenated equal the single-device call
    bit for bit, and the reduced histograms equal the single-device histograms.  (Two-device case: skipped on a 1-GPU box.)"""
    from em_model_manned_bayes_b200.shard import sample_tracks_multi
    lib = L.lib()
    have = lib.emb_device_count()
    if n_devices > have:
        pytest.skip("needs %d GPUs" % n_devices)
    assert lib.emb_nccl_available() == 1
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    n, T = 5003, 96
    res, hi, ht = sample_tracks_multi(m, n, T, seed=19, first_sample=100, opts=m.uncor_opts(), n_devices=n_devices)
    hi1 = np.zeros((m.n_initial, 64), dtype=np.uint64)
    ht1 = np.zeros((m.n_dyn, 64), dtype=np.uint64)
    one = m.sample_tracks(n, T, seed=19, first_sample=100, opts=m.uncor_opts(), hist_initial=hi1, hist_transition=ht1)
    assert sum(r.n for r in res) == n and len(res) == (n_devices or have)
    assert np.array_equal(np.concatenate([r.bins for r in res]), one.bins)
    assert np.array_equal(np.concatenate([r.values for r in res]), one.values)
    assert np.array_equal(np.concatenate([r.init_values for r in res], axis=1), one.init_values)
    assert np.array_equal(hi, hi1) and np.array_equal(ht, ht1)
    assert int(hi.sum()) == n * m.n_initial and int(ht.sum()) == n * (T - 1) * m.n_dyn


def test_tracks_device_buffers_match_host_buffers(model_paths, golden):
    name = "uncor_v2p1_n24_T300_seed1"
    got = _run_tracks(model_paths, cases.TRACK_CASES[name], device="cuda:0")
    cases.check_tracks(got, golden[name])


@pytest.mark.parametrize("name", sorted(cases.INITIAL_CASES))
def test_initial_matches_golden(model_paths, golden, name):
    c = cases.INITIAL_CASES[name]
    m = EncounterModel(model_paths[c["model"]])
    bins, vals, att = m.sample_initial(c["n"], seed=c["seed"], first_sample=c["first"])
    assert np.array_equal(bins, golden[name]["bins"])
    assert np.array_equal(vals, golden[name]["values"])
    assert np.all(att == 1)


@pytest.mark.parametrize("name", sorted(cases.TERMINAL_CASES))
def test_terminal_geometry_matches_golden(model_paths, golden, name):
    c = cases.TERMINAL_CASES[name]
    m = CorTerminalModel(model_paths[c["model"]])
    if c["start"] is not None:
        m.start = c["start"]
    out_inits, bins, att = m.sample_raw(c["n"], seed=c["seed"])
    assert np.array_equal(bins, golden[name]["bins"])
    assert np.array_equal(out_inits, golden[name]["values"])
    assert np.array_equal(att.astype(np.int64), golden[name]["attempts"].astype(np.int64))


@pytest.mark.parametrize("model,n,T,seed", [("uncor_1200code_v2p1", 40, 97, 21), ("glider_v1", 33, 61, 22)])
def test_tracks_match_live_oracle(model_paths, model, n, T, seed):
    """Fresh seeds (not in the golden file): oracle computed here, on the box's CPU."""
    p = em_read(model_paths[model])
    bins, vals, dyn, tv = oracle_dense(p, uncor_sample(p, n, T, KeyedPhilox(seed), first_sample=5))
    m = UncorEncounterModel(model_paths[model])
    res = m.sample_compact(n, T, seed=seed, first_sample=5)
    assert res.dyn_vars == dyn and res.tv_vars == tv
    assert np.array_equal(res.bins, bins)
    assert np.all(np.abs(res.values.astype(np.float64) - vals) <= 1e-6 * np.abs(vals))


SWEEP = [("uncor_1200code_v2p1", True, 4000, 300), ("uncor_allcode_fwsingle_v1", True, 2000, 600),
         ("uncor_1200only_fwse_v1p2", True, 3000, 200), ("glider_v1", True, 3000, 200), ("paramotor_v1", True, 3000, 150),
         ("littoral_uncor_v1", True, 3000, 120), ("cor_v1", False, 3000, 60), ("balloon_v1", False, 5000, 100),
         # one model per remaining compiled shape of k_tracks_fast (emb_fast.cuh: EMB_FAST_SHAPES)
         ("fai1_v1", True, 3000, 150), ("uncor_1200code_v1", True, 2000, 200), ("blimp_v1", True, 3000, 150),
         ("dueregard_v1", True, 2000, 200), ("haa_v1", True, 2000, 200), ("littoral_cor_v1", False, 2000, 120),
         ("weatherballoon_v1", False, 4000, 100)]


@pytest.mark.parametrize("model,uncor,n,T", SWEEP)
def test_tracks_match_c_oracle_mid_scale(model_paths, model, uncor, n, T):
    """Every packed model at a few 10^5..10^6 track-seconds against the plain-C restatement (oracle/oracle_c.c, the size the
    oracle finishes in seconds): initial bins, attempts and every dense bin bit-exact, values within 1e-6 relative."""
    from oracle.c_oracle import COracle
    p = em_read(model_paths[model])
    ref = COracle(p, uncor=uncor).sample_tracks(n, T, seed=91, first_sample=10 ** 12, threads=0)
    assert ref["rc"] == 0
    m = (UncorEncounterModel if uncor else EncounterModel)(model_paths[model])
    got = m.sample_compact(n, T, seed=91, first_sample=10 ** 12) if uncor else m.sample_tracks(n, T, seed=91, first_sample=10 ** 12)
    assert L.lib().emb_debug_last_kernel_fast() == 1, "no specialised kernel for this model shape"
    dyn = [int(v) - 1 for v in np.asarray(p.temporal_map)[:, 0]]
    tv = [v - 1 for v in got.tv_vars]
    assert np.array_equal(np.asarray(got.init_bins).T, ref["init_bins"])
    assert np.array_equal(np.asarray(got.attempts).astype(np.int64), ref["attempts"].astype(np.int64))
    assert np.array_equal(np.asarray(got.init_values).T, ref["init_values"])
    assert np.array_equal(np.asarray(got.bins), ref["sample_bins"][:, dyn, :])
    want = ref["samples"][:, tv, :]
    gv = np.asarray(got.values, dtype=np.float64)
    assert np.all(np.abs(gv - want) <= 1e-6 * np.abs(want))


def test_shard_invariance_and_determinism(model_paths):
    """Global-index keying: any split of [0, n) gives identical results, twice."""
    m = UncorEncounterModel(model_paths["uncor_allcode_fwsingle_v1"])
    n, T = 5000, 600
    whole = m.sample_compact(n, T, seed=7)
    again = m.sample_compact(n, T, seed=7)
    assert np.array_equal(whole.bins_tiled, again.bins_tiled) and np.array_equal(whole.values_tiled, again.values_tiled)
    parts = [m.sample_compact(k, T, seed=7, first_sample=f) for f, k in ((0, 1250), (1250, 1250), (2500, 2499), (4999, 1))]
    assert np.array_equal(np.concatenate([p.bins for p in parts]), whole.bins)
    assert np.array_equal(np.concatenate([p.values for p in parts]), whole.values)
    assert np.array_equal(np.concatenate([p.init_values for p in parts], axis=1), whole.init_values)
    other = m.sample_compact(64, T, seed=8)
    assert not np.array_equal(other.bins, whole.bins[:64])


def test_ragged_and_tiny_sizes(model_paths):
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    ref = m.sample_compact(131, 49, seed=3)
    for n, T in [(1, 1), (1, 2), (3, 15), (3, 16), (3, 17), (129, 49)]:
        r = m.sample_compact(n, T, seed=3)
        assert r.bins.shape == (n, 3, T) and r.values.shape == (n, 4, T)
        assert np.array_equal(r.bins, ref.bins[:n, :, :T])          # prefix property in n and T
        assert np.array_equal(r.values, ref.values[:n, :, :T])
    z = m.sample_compact(0, 10, seed=3)
    assert z.bins.shape[0] == 0


def test_full_size_properties_config3(model_paths):
    """uncor_allcode_fwsingle_v1 at BASELINE configs[2]'s per-GPU size, 1.25 M tracks x 600 s, outputs resident in HBM
    (14.3 GB): histograms == dense counts, initial-node marginals pass a chi-square test against the normalised count tables."""
    import torch
    m = UncorEncounterModel(model_paths["uncor_allcode_fwsingle_v1"])
    n, T = 1_250_000, 600
    hi = torch.zeros((m.n_initial, 64), dtype=torch.int64, device="cuda:0")
    ht = torch.zeros((m.n_dyn, 64), dtype=torch.int64, device="cuda:0")
    res = m.sample_compact(n, T, seed=11, device="cuda:0", hist_initial=hi, hist_transition=ht)
    torch.cuda.synchronize()
    bins = res.bins
    assert int(bins.min()) >= 1
    for d in range(m.n_dyn):
        cnt = torch.zeros(64, dtype=torch.int64, device="cuda:0")
        for lo in range(0, n, 250_000):      # bincount wants int64: a quarter of a million tracks at a time
            cnt += torch.bincount(bins[lo:lo + 250_000, d, 1:].reshape(-1).to(torch.int64) - 1, minlength=64)
        assert torch.equal(cnt, ht[d])
    assert hi.sum(dim=1).tolist() == [n] * m.n_initial
    # root variable G: exact marginal from the table
    N = m.N_initial
    pG = N[0][:, 0] / N[0][:, 0].sum()
    obs = hi[0, : len(pG)].cpu().numpy().astype(np.float64)
    # rejection (v*1.68781 > |dh|/60) removes a negligible fraction; chi-square with generous bound
    chi2 = ((obs - n * pG) ** 2 / np.maximum(n * pG, 1e-9)).sum()
    assert chi2 < 40.0, chi2
    # transition bins must stay inside [1, r]
    for d, v in enumerate(res.dyn_vars):
        assert int(bins[:, d, :].max()) <= int(m.r_initial[v - 1])
    vals = res.values
    assert bool(torch.isfinite(vals).all())


def test_initial_full_size_chi_square_config2(model_paths):
    """glider_v1 initial network at BASELINE configs[1]'s size, 1e8 samples on the device: every node marginal against exact
    enumeration (chi-square at the 1e8 scale resolves a relative bias of 1e-4 in a bin of probability 0.1)."""
    import torch
    m = EncounterModel(model_paths["glider_v1"])
    n = 100_000_000
    bins, vals, _ = m.sample_initial(n, seed=5, device="cuda:0", want_attempts=False, values_fp32=True)
    torch.cuda.synchronize()
    # exact joint by enumeration of the 5-node network (4*8*5*7*7 = 7840 states)
    N, G, r = m.N_initial, m.G_initial, m.r_initial
    dims = [int(x) for x in r]
    grids = np.indices(dims)
    joint = np.ones(dims)
    for i0 in range(len(dims)):
        par = np.nonzero(G[:, i0])[0]
        j = np.zeros(dims, dtype=np.int64)
        stride = 1
        for p_ in par:                                   # asub2ind.m:13 strides
            j += stride * grids[p_]
            stride *= dims[p_]
        tab = N[i0] / np.maximum(N[i0].sum(axis=0, keepdims=True), 1e-300)
        joint = joint * tab[grids[i0], j]
    assert abs(joint.sum() - 1.0) < 1e-9
    for i in range(len(r)):
        marg = joint.sum(axis=tuple(k for k in range(len(r)) if k != i))
        obs = torch.bincount(bins[:, i].to(torch.int64) - 1, minlength=int(r[i])).cpu().numpy().astype(np.float64)
        keep = marg > 0
        assert obs[~keep].sum() == 0
        chi2 = ((obs[keep] - n * marg[keep]) ** 2 / (n * marg[keep])).sum()
        assert chi2 < 50.0, (i, chi2)


def test_step_word_statistics_spec_v5(model_paths):
    """Stream spec v5 derives the transition select, the resample gate and the de-discretisation of a variable in
    a second from ONE Philox word.  With every initial variable preset the frozen-parent columns are known, so:
    transition bins ~ the normalised count column (chi-square), rows per variable = gates (rate) + changes
    (1 - stay probability) within 5 sigma, and the de-discretised values of the gate rows are uniform in their bin."""
    import torch
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    start = [1, 4, 2, 4, 3, 4, 4]          # columns 3918 / 1960 / 558 of the three dynamic variables
    n, T = 40_000, 250
    ht = torch.zeros((m.n_dyn, 64), dtype=torch.int64, device="cuda:0")
    o = m.uncor_opts(start=start)
    m.sample_tracks(n, T, seed=77, opts=o, device="cuda:0", hist_transition=ht, want_values=False, want_init=False)
    ev = m.sample_events(n, T, seed=77, opts=m.uncor_opts(start=start), device="cuda:0", want_init=False)
    torch.cuda.synchronize()
    rows = ev.events.cpu().numpy().view(L.EVENT_DTYPE)
    G, r, Nt = m.G_transition, m.r_transition, m.N_transition
    x = np.array(start) - 1
    steps = n * (T - 1)
    for d, (vt, vt1) in enumerate(m.temporal_map):
        par = np.nonzero(G[:, vt1 - 1])[0]
        j, stride = 0, 1
        for p_ in par:                                       # asub2ind.m:13-14
            j += stride * x[p_]
            stride *= int(r[p_])
        col = Nt[vt1 - 1][:, j]
        pcol = col / col.sum()
        obs = ht[d, : len(pcol)].cpu().numpy().astype(np.float64)
        keep = pcol > 0
        assert obs[~keep].sum() == 0
        chi2 = ((obs[keep] - steps * pcol[keep]) ** 2 / (steps * pcol[keep])).sum()
        assert chi2 < 45.0, (d, chi2)
        # rows of this variable: fired gates (n*T draws at `rate`) + bin changes (i.i.d. draws: P(new != old))
        rate = float(m.resample_rates[vt - 1])
        p_change = 1.0 - float((pcol ** 2).sum())            # consecutive i.i.d. draws differ (first step: vs the preset bin)
        got = int((rows["var"] == vt).sum())
        want = n * T * rate + steps * p_change
        sd = np.sqrt(n * T * rate * (1 - rate) + 3.0 * steps * p_change)
        assert abs(got - want) < 6 * sd + 0.002 * want, (vt, got, want, sd)
    # static resampled variable v (4): only gate rows, value uniform in its bin [edges of bin 4]
    v = 4
    sel = rows[rows["var"] == v]
    rate = float(m.resample_rates[v - 1])
    assert abs(len(sel) - n * T * rate) < 6 * np.sqrt(n * T * rate)
    a, b = m.boundaries[v - 1][start[v - 1] - 1], m.boundaries[v - 1][start[v - 1]]
    u = (sel["value"].astype(np.float64) - a) / (b - a)
    assert u.min() >= 0.0 and u.max() <= 1.0
    h = np.histogram(u, bins=20, range=(0, 1))[0].astype(np.float64)
    chi2 = ((h - len(u) / 20) ** 2 / (len(u) / 20)).sum()
    assert chi2 < 60.0, chi2


@pytest.mark.gpu
def _row_table(ev, n, T):
    """rows of an EventResult on the device -> numpy (track, second, var, bin, value) with the closing rows dropped"""
    rows = ev.events.cpu().numpy().view(L.EVENT_DTYPE)
    off = ev.offsets.cpu().numpy()
    trk = np.repeat(np.arange(n), np.diff(off))
    cum = np.cumsum(rows["dt"].astype(np.int64))
    before = np.where(off[:-1] > 0, cum[np.maximum(off[:-1], 1) - 1], 0)
    sec = cum - np.repeat(before, np.diff(off))
    keep = rows["var"] > 0
    return trk[keep], sec[keep], rows["var"][keep].astype(np.int64), rows["bin"][keep].astype(np.int64), rows["value"][keep].astype(np.float64)


def _uniformity(u, cells=16):
    assert u.min() >= 0.0 and u.max() <= 1.0
    h = np.histogram(u, bins=cells, range=(0, 1))[0].astype(np.float64)
    return float(((h - len(u) / cells) ** 2 / (len(u) / cells)).sum())


def test_joint_law_of_select_gate_and_value_fast_branch(model_paths):
    """The reference draws the transition select, the resample gate and the de-discretisation uniform of a variable
    independently; stream spec v5 derives them from the variable's word of the second (select on k, gate on k*A, value on
    k*B + k').  Joint check on the frozen-parent branch with every initial variable preset (the column of each dynamic
    variable is then known exactly), 2e7 track-seconds:
      * chi-square of the 2 x r table (gate fired) x (new bin) against  p(bin) * rate  -- including the bins of probability
        < 1e-3 of columns 3918 / 1960 / 558;
      * the value of every row is uniform inside its bin GIVEN the kind of row: gate only, transition only, gate and
        transition in the same second, and transition into a bin of probability < 1e-2."""
    import torch
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    start = [1, 4, 2, 4, 3, 4, 4]
    n, T = 80_000, 250
    dense = m.sample_tracks(n, T, seed=123, opts=m.uncor_opts(start=start), device="cuda:0", want_values=False, want_init=False)
    ev = m.sample_events(n, T, seed=123, opts=m.uncor_opts(start=start), device="cuda:0", want_init=False)
    torch.cuda.synchronize()
    bins = dense.bins.cpu().numpy()                      # (n, n_dyn, T) 1-based; column c = second c + 1 (state after step e = c)
    trk, sec, var, rbin, val = _row_table(ev, n, T)
    G, r, Nt = m.G_transition, m.r_transition, m.N_transition
    x = np.array(start) - 1
    rare_seen = 0
    for d, (vt, vt1) in enumerate(m.temporal_map):
        par = np.nonzero(G[:, vt1 - 1])[0]
        j, stride = 0, 1
        for p_ in par:
            j += stride * x[p_]
            stride *= int(r[p_])
        col = Nt[vt1 - 1][:, j]
        pcol = col / col.sum()
        rate = float(m.resample_rates[vt - 1])
        b = bins[:, d, :].astype(np.int64)
        prev, new = b[:, :-1], b[:, 1:]                  # seconds e = 1 .. T-1
        sel = (var == vt) & (sec >= 1) & (sec <= T - 1)
        t_, e_, rb_, v_ = trk[sel], sec[sel], rbin[sel], val[sel]
        is_gate = rb_ == prev[t_, e_ - 1]                # a gate row re-emits the PRE-transition bin (resample_events.m:26-29)
        fired = np.zeros((n, T - 1), dtype=bool)
        fired[t_[is_gate], e_[is_gate] - 1] = True
        changed = new != prev
        # every transition row is there and carries the new bin
        tr = ~is_gate
        assert tr.sum() == changed.sum() and np.array_equal(rb_[tr], new[t_[tr], e_[tr] - 1])
        steps = n * (T - 1)
        obs = np.zeros((2, len(pcol)))
        for f in (0, 1):
            obs[f] = np.bincount(new[fired == bool(f)] - 1, minlength=len(pcol))
        exp = np.outer([1 - rate, rate], pcol) * steps
        keep = exp > 0
        assert obs[~keep].sum() == 0
        chi2 = float(((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum())
        assert chi2 < 60.0, (vt, chi2, obs, exp)          # <= 2r - 1 = 13 degrees of freedom
        rare_seen += int(obs[:, (pcol > 0) & (pcol < 1e-3)].sum())
        # values: u = (value - a) / (b - a) inside the row's bin
        edges = np.asarray(m.boundaries[vt - 1])
        zb = m.zero_bins[vt - 1][0] if m.zero_bins[vt - 1] else 0
        nz = rb_ != zb
        u = (v_ - edges[rb_ - 1]) / (edges[rb_] - edges[rb_ - 1])
        assert np.all(v_[~nz] == 0.0)
        both = np.zeros(len(rb_), dtype=bool)
        both[tr] = fired[t_[tr], e_[tr] - 1]              # transition rows of seconds in which the gate fired too
        into_rare = tr & np.isin(rb_ - 1, np.nonzero(pcol < 1e-2)[0])
        for name, pick in (("gate", is_gate), ("transition", tr), ("gate+transition", both), ("rare transition", into_rare)):
            pick = pick & nz
            if pick.sum() >= 800:
                c2 = _uniformity(u[pick])
                assert c2 < 55.0, (vt, name, int(pick.sum()), c2)
    assert rare_seen > 1000      # the tables really contained bins of probability < 1e-3 (columns 3918 and 558)


@pytest.mark.gpu
def test_joint_law_of_gate_and_value_slow_branch(model_paths):
    """The same on the slow branch (glider_v1: dynamic -> dynamic edges, parents re-evaluated every second), where the
    transition columns vary along the track: the gate frequency GIVEN that the variable changes in that second equals its
    resample rate, and the values of gate rows, transition rows and same-second rows are uniform inside their bins."""
    import torch
    m = UncorEncounterModel(model_paths["glider_v1"])
    n, T = 60_000, 200
    dense = m.sample_compact(n, T, seed=321, device="cuda:0", want_values=False, want_init=False)
    ev = m.sample_events_uncor(n, T, seed=321, device="cuda:0", want_init=False)
    torch.cuda.synchronize()
    bins = dense.bins.cpu().numpy()
    trk, sec, var, rbin, val = _row_table(ev, n, T)
    for d, (vt, vt1) in enumerate(m.temporal_map):
        rate = float(m.resample_rates[vt - 1])
        b = bins[:, d, :].astype(np.int64)
        prev, new = b[:, :-1], b[:, 1:]
        sel = (var == vt) & (sec >= 1) & (sec <= T - 1)
        t_, e_, rb_, v_ = trk[sel], sec[sel], rbin[sel], val[sel]
        is_gate = rb_ == prev[t_, e_ - 1]
        fired = np.zeros((n, T - 1), dtype=bool)
        fired[t_[is_gate], e_[is_gate] - 1] = True
        changed = new != prev
        k, tot = int((fired & changed).sum()), int(changed.sum())
        assert tot > 20_000
        assert abs(k - tot * rate) < 5.0 * np.sqrt(tot * rate * (1 - rate)) + 1.0, (vt, k, tot, rate)
        k0, tot0 = int((fired & ~changed).sum()), int((~changed).sum())
        assert abs(k0 - tot0 * rate) < 5.0 * np.sqrt(tot0 * rate * (1 - rate)) + 1.0, (vt, k0, tot0, rate)
        edges = np.asarray(m.boundaries[vt - 1])
        zb = m.zero_bins[vt - 1][0] if m.zero_bins[vt - 1] else 0
        nz = rb_ != zb
        u = (v_ - edges[rb_ - 1]) / (edges[rb_] - edges[rb_ - 1])
        tr = ~is_gate
        both = np.zeros(len(rb_), dtype=bool)
        both[tr] = fired[t_[tr], e_[tr] - 1]
        for name, pick in (("gate", is_gate), ("transition", tr), ("gate+transition", both)):
            pick = pick & nz
            if pick.sum() >= 800:
                c2 = _uniformity(u[pick])
                assert c2 < 55.0, (vt, name, int(pick.sum()), c2)


def test_em_sample_files_match_oracle(model_paths, tmp_path):
    """em_sample.m:59-100: the two legacy text files.  Headers, ids and seconds are identical to the
    oracle's; numbers are compared after parsing (the dense values are fp32, %g prints 6 digits)."""
    from em_model_manned_bayes_b200.em_sample import em_sample
    from oracle.drivers import em_sample_text
    for model, n, T in (("cor_v1", 12, 60), ("uncor_1200code_v2p1", 9, 35)):
        fi, ft = str(tmp_path / (model + "_initial.txt")), str(tmp_path / (model + "_transition.txt"))
        em_sample(model_paths[model], fi, ft, num_initial_samples=n, num_transition_samples=T, rng_seed=42)
        p = em_read(model_paths[model])
        want_i, want_t = em_sample_text(p, n, T, KeyedPhilox(42))
        for got, want in ((open(fi).read(), want_i), (open(ft).read(), want_t)):
            g, w = got.splitlines(), want.splitlines()
            assert len(g) == len(w) and g[0] == w[0]
            a = np.array([[float(x) for x in line.split()] for line in g[1:]])
            b = np.array([[float(x) for x in line.split()] for line in w[1:]])
            assert a.shape == b.shape
            assert np.all(np.abs(a - b) <= 2e-5 * np.abs(b))
        assert open(fi).read() == want_i          # initial values are fp64 on both sides: identical text


def test_errors_through_the_abi(model_paths):
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    with pytest.raises(L.EmbError, match="Attempt to preset a dependent variable"):
        m.sample_tracks(4, 10, seed=1, start=[None, 2, None, None, None, None, None])
    b = UncorEncounterModel(model_paths["balloon_v1"])
    with pytest.raises(L.EmbError, match="dynvar:empty"):
        b.sample_compact(4, 10, seed=1)
    t = EncounterModel(model_paths["terminal_v3_radar_encounter_model"])
    with pytest.raises(L.EmbError, match="dynvar:empty"):
        t.sample_tracks(4, 10, seed=1)


def test_kernels_actually_launched(model_paths):
    lib = L.lib()
    before = lib.emb_launch_count()
    UncorEncounterModel(model_paths["uncor_1200code_v2p1"]).sample_compact(8, 8, seed=1)
    assert lib.emb_launch_count() == before + 1


def test_temporaries_are_pooled_and_can_be_trimmed(model_paths):
    """Host-buffer calls stage through the device memory pool (emb_api.cpp: tmp_alloc): results do not depend on whether the
    pool is warm, and emb_trim_device_memory gives the memory back."""
    import torch
    lib = L.lib()
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    a = m.sample_events(3000, 120, seed=5, opts=m.uncor_opts())
    free_warm = torch.cuda.mem_get_info()[0]
    assert lib.emb_trim_device_memory(-1) == 0
    assert torch.cuda.mem_get_info()[0] >= free_warm
    b = m.sample_events(3000, 120, seed=5, opts=m.uncor_opts())
    assert np.array_equal(np.asarray(a.offsets), np.asarray(b.offsets))
    assert np.array_equal(np.asarray(a.events).view(np.uint8), np.asarray(b.events).view(np.uint8))


def test_enqueue_only_passes_equal_synchronous_passes(model_paths):
    """EMB_MEM_ASYNC: emb_sample_tracks only enqueues; the outputs equal those of the synchronous call, emb_async_status
    reports nothing for a healthy pass and EMB_E_REJECT after a pass whose rejection loop was cut to one attempt."""
    import torch
    from em_model_manned_bayes_b200.model import async_status
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    a = m.sample_compact(5000, 90, seed=77, device="cuda:0")
    b = m.sample_compact(5000, 90, seed=78, device="cuda:0")
    m.sample_compact(5000, 90, seed=77, device="cuda:0", out=b, enqueue_only=True)
    async_status(0)
    assert torch.equal(a.bins_tiled, b.bins_tiled) and torch.equal(a.values_tiled, b.values_tiled)
    assert torch.equal(a.init_values, b.init_values)
    # a rejection loop that may not retry: some of 200 000 samples are rejected by v*1.68781 > |dh|/60 on their first attempt
    o = m.uncor_opts(max_attempts=1)
    o.max_attempts = 1
    big = m.sample_tracks(200_000, 8, seed=3, opts=m.uncor_opts(), device="cuda:0")
    if int((big.attempts.to(torch.int32) > 2).sum()) > 0:
        m.sample_tracks(200_000, 8, seed=3, opts=o, device="cuda:0", out=big, enqueue_only=True)
        with pytest.raises(L.EmbError):
            async_status(0)
        async_status(0)      # the flag is cleared by the report
    with pytest.raises(L.EmbError):
        m.sample_events(10, 10, seed=1, opts=_async_host_opts(m))


def _async_host_opts(m):
    o = m.uncor_opts()
    o.mem = L.EMB_MEM_HOST | L.EMB_MEM_ASYNC
    return o


def test_enqueue_only_initial_equals_synchronous(model_paths):
    import torch
    from em_model_manned_bayes_b200.model import async_status
    g = EncounterModel(model_paths["glider_v1"])
    a = g.sample_initial(100_003, seed=9, device="cuda:0")
    b = g.sample_initial(100_003, seed=10, device="cuda:0")
    g.sample_initial(100_003, seed=9, device="cuda:0", out=b, enqueue_only=True)
    async_status(0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])


def test_pipelined_host_events_equal_the_single_pass(model_paths):
    """Host buffers and >= 8192 tracks take the chunked path (count -> scan with carry -> write -> copy per chunk):
    same offsets, rows and per-track outputs as the device-resident single pass; a too small buffer still reports the
    number of rows needed."""
    import torch
    n, T = 50_003, 150
    m = UncorEncounterModel(model_paths["uncor_1200code_v2p1"])
    h = m.sample_events(n, T, seed=41, first_sample=7, opts=m.uncor_opts())
    d = m.sample_events(n, T, seed=41, first_sample=7, opts=m.uncor_opts(), device="cuda:0")
    assert h.total == d.total
    assert np.array_equal(np.asarray(h.offsets), d.offsets.cpu().numpy())
    rows_h = np.asarray(h.events)[:h.total].view(np.int64)
    assert np.array_equal(rows_h, d.events[:d.total].cpu().numpy())
    assert np.array_equal(np.asarray(h.init_values), d.init_values.cpu().numpy())
    assert np.array_equal(np.asarray(h.init_bins), d.init_bins.cpu().numpy())
    assert np.array_equal(np.asarray(h.attempts).astype(np.int64), d.attempts.cpu().numpy().astype(np.int64))
    small = m.sample_events(n, T, seed=41, first_sample=7, opts=m.uncor_opts(), capacity=1000)   # LIMIT -> retried by the wrapper
    assert small.total == h.total and np.array_equal(np.asarray(small.events)[:small.total].view(np.int64), rows_h)
