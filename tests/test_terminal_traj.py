"""Terminal trajectory chains: @CorTerminalModel/createEncounter.m:93-329.

CPU suite: the host emulation of the device code (tests/emu, same emb_terminal.cuh) against the Python oracle
(oracle/terminal.py) on the synthetic trajectory DBNs, fed the same keyed uniforms, and against the committed golden
vectors.  GPU suite: the C ABI (emb_terminal_propagate) against the same goldens plus size-independent properties."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from em_model_manned_bayes_b200 import _lib as L
from em_model_manned_bayes_b200.synthetic import TRAJECTORY_STEMS, write_terminal_model_set

SLOTS = (("ownship_landing_model", "own_fwd", 0), ("ownship_landing_model_reverse", "own_bck", 0),
         ("ownship_takeoff_model", "own_fwd", 1), ("ownship_takeoff_model_reverse", "own_bck", 1),
         ("intruder_landing_model", "int_fwd", 0), ("intruder_landing_model_reverse", "int_bck", 0),
         ("intruder_takeoff_model", "int_fwd", 1), ("intruder_takeoff_model_reverse", "int_bck", 1),
         ("intruder_transit_model", "int_fwd", 2), ("intruder_transit_model_reverse", "int_bck", 2))   # TermParams::m order

LIMITS = {"GENERIC": (50.0, 506.0, 12.0, 5000.0, 6000 / 60), "RTCA228_A1": (169.0, 491.0, 1.5, 5000.0, 2500 / 60),
          "RTCA228_A3": (68.0, 186.0, 7.0, 5000.0, 500 / 60), "TEST": (68.0, 186.0, 7.0, 1200.0, 500 / 60)}


@pytest.fixture(scope="session")
def traj_paths(tmp_path_factory):
    return write_terminal_model_set(str(tmp_path_factory.mktemp("traj_models")))


def geo_from_golden(golden, n):
    """(12, n) sample_geo rows from the golden terminal-geometry samples (labels of terminal_v3_radar_encounter_model)."""
    labels = ["airspace_class", "own_intent", "int_intent", "int_type", "int_runway", "own_distance", "own_bearing", "own_alt",
              "own_speed", "own_heading", "int_distance", "int_bearing", "int_alt", "int_speed", "int_heading"]
    v = golden["terminal_geo_n64_seed13"]["values"][:n]
    fields = ("own_intent", "own_distance", "own_bearing", "own_alt", "own_heading", "own_speed",
              "int_intent", "int_distance", "int_bearing", "int_alt", "int_heading", "int_speed")
    return np.ascontiguousarray(np.stack([v[:, labels.index(f)] for f in fields]))


def emu_propagate(traj_paths, geo, seed, first, tmax_s, types=("GENERIC", "GENERIC"), max_attempts=0):
    lib = H.emu_lib()
    models = [H.EmuModel(traj_paths[stem]) for stem, _, _ in SLOTS]
    for m in models:
        m.set_prior(1, L.EMB_PRIOR_STAY, 1.0)
    arr = (C.c_void_p * 10)(*[m.h.value for m in models])
    n = geo.shape[1]
    tmax = int(tmax_s)
    traj = np.zeros((5, 2, 2 * tmax + 1, n), dtype=np.float32)
    ln = np.zeros((4, n), dtype=np.int16)
    lim = (L.DynLimits * 2)(*[L.DynLimits(*LIMITS[t]) for t in types])
    rows = (C.c_int32 * 12)(*range(12))
    rc = lib.emu_terminal_propagate(arr, seed, first, n, geo.ctypes.data, n, rows, float(tmax_s), lim, max_attempts,
                                    traj.ctypes.data, ln.ctypes.data)
    return rc, traj, ln


_ORACLE_MODELS = {}


def oracle_propagate(traj_paths, geo, seed, first, tmax_s, types=("GENERIC", "GENERIC")):
    from oracle import terminal as T
    from oracle.em_read import em_read
    ms = {}
    for stem, group, k in SLOTS:
        if traj_paths[stem] not in _ORACLE_MODELS:
            _ORACLE_MODELS[traj_paths[stem]] = em_read(traj_paths[stem])
        ms[(group, k)] = _ORACLE_MODELS[traj_paths[stem]]
    return T.create_encounters(ms, geo, seed, first, tmax_s, types)


def check_traj(got, got_len, want, want_len, rtol=1e-6):
    assert np.array_equal(np.asarray(got_len), np.asarray(want_len)), "chain lengths differ"
    got = np.asarray(got, dtype=np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "occupied slots differ"
    ok = ~np.isnan(want)
    err = np.abs(got[ok] - want[ok])
    assert np.all(err <= rtol * np.abs(want[ok]) + 1e-12), "trajectory values differ by more than 1e-6 relative (max %g)" % err.max()


# ---------------------------------------------------------------------------------------------------
def test_oracle_helpers_known_answers():
    from oracle import terminal as T
    assert T.cosd(90.0) == 0.0 and T.sind(180.0) == 0.0 and T.cosd(-270.0) == 0.0 and T.sind(-90.0) == -1.0
    assert T.cosd(360.0) == 1.0 and T.sind(450.0) == 1.0
    assert T.wrap_to_360(0.0) == 0.0 and T.wrap_to_360(360.0) == 360.0 and T.wrap_to_360(720.0) == 360.0
    assert T.wrap_to_360(-10.0) == 350.0 and T.wrap_to_360(370.0) == 10.0
    assert T.round2(1.004) == 1.0 and T.round2(-2.674) == -2.67 and T.round2(2.676) == 2.68 and T.round2(0.125) == 0.13 and T.round2(-0.125) == -0.13
    from oracle.sampler import discretize_bayes
    cut = [0.5, 1, 2, 3, 4, 5]                     # distance cutpoints (terminal_v3_radar_encounter_model.txt:29)
    assert [discretize_bayes(x, cut) for x in (0.0, 0.49, 0.5, 4.99, 5.0, 9.0)] == [1, 1, 2, 6, 7, 7]


def test_oracle_chain_invariants(traj_paths, golden):
    """createEncounter.m semantics on the oracle itself: termination rules, rate limits, merge order."""
    geo = geo_from_golden(golden, 6)
    traj, ln = oracle_propagate(traj_paths, geo, seed=21, first=0, tmax_s=60)
    tmax = 60
    assert ln.min() >= 1 and ln.max() <= tmax + 1
    for s in range(geo.shape[1]):
        for ac in range(2):
            occ = ~np.isnan(traj[0, ac, :, s])
            lo, hi = tmax - (ln[2 * ac + 1, s] - 1), tmax + ln[2 * ac, s] - 1
            assert occ[lo:hi + 1].all() and occ.sum() == hi - lo + 1                   # contiguous in time (:74-84)
            z = traj[2, ac, lo:hi + 1, s]
            assert np.all(np.abs(np.diff(z[tmax - lo:])) <= 100.0 + 1e-9)              # maxVertRate (:180-184) forward part
            v = traj[4, ac, lo:hi + 1, s]
            assert np.all(v[1:-1] <= 506.0 + 1e-9)
    # the t = 0 state is the encounter geometry (:45-49)
    assert np.allclose(traj[2, 0, tmax, :], geo[3]) and np.allclose(traj[2, 1, tmax, :], geo[9])
    assert np.allclose(np.hypot(traj[0, 0, tmax, :], traj[1, 0, tmax, :]), geo[1])


@pytest.mark.parametrize("types,tmax_s,seed,first", [(("GENERIC", "GENERIC"), 120, 21, 0),
                                                     (("TEST", "RTCA228_A1"), 45, 22, 1234567890123),
                                                     (("RTCA228_A3", "GENERIC"), 0, 23, 5),
                                                     (("GENERIC", "TEST"), 30.5, 24, 0)])
def test_emu_chains_match_oracle(traj_paths, golden, types, tmax_s, seed, first):
    geo = geo_from_golden(golden, 10)
    if types[0] != "GENERIC":
        geo[5] = np.clip(geo[5], 70.0, 180.0)
    rc, traj, ln = emu_propagate(traj_paths, geo, seed, first, tmax_s, types)
    assert rc == 0
    want, want_len = oracle_propagate(traj_paths, geo, seed, first, tmax_s, types)
    check_traj(traj, ln, want, want_len)


def test_speed_edges_on_the_dynamic_limits(tmp_path, golden, monkeypatch):
    """createEncounter.m:221-226 clamps a sampled speed to minVel / maxVel; the next state records norm(v) (:168) and
    discretises it (:290) after v was rotated.  With speed bin edges ON those limits the cell depends on the last bit of
    norm(R v): the device code must take norm(v) from (vx, vy) whenever v changed, as the reference does, instead of carrying
    the clamped value.  The last bit of sind/cosd decides these cells (MATLAB's own are closed source), so for this test the
    oracle is given the chain kernel's sind/cosd (checked against libm in
    test_sind_cosd_and_constant_division_of_the_chain_kernel): everything else must then agree exactly."""
    from oracle import terminal as T
    lib = H.emu_lib()

    def _sc(x):
        a, s1, c1 = np.array([float(x)]), np.zeros(1), np.zeros(1)
        lib.emu_sincosd(1, a.ctypes.data, s1.ctypes.data, c1.ctypes.data)
        return float(s1[0]), float(c1[0])

    monkeypatch.setattr(T, "sind", lambda x: _sc(x)[0])
    monkeypatch.setattr(T, "cosd", lambda x: _sc(x)[1])
    edges = [0, 50, 68, 100, 169, 186, 338, 491, 506, 600]
    paths = write_terminal_model_set(str(tmp_path / "edge_models"), seed=7, speed_edges=edges)
    geo = geo_from_golden(golden, 24)
    for types, seed in ((("RTCA228_A3", "TEST"), 41), (("RTCA228_A1", "GENERIC"), 42), (("TEST", "RTCA228_A3"), 43)):
        g = geo.copy()
        g[5] = np.clip(g[5], 70.0, 180.0)
        g[11] = np.clip(g[11], 70.0, 180.0)
        rc, traj, ln = emu_propagate(paths, g, seed, 0, 90, types)
        assert rc == 0
        want, want_len = oracle_propagate(paths, g, seed, 0, 90, types)
        check_traj(traj, ln, want, want_len)
        spd = traj[4][~np.isnan(traj[4])]
        assert any(np.any(np.abs(spd - e) < 1e-3) for e in (68.0, 169.0, 186.0, 338.0)), "no speed was clamped to a limit"


def test_emu_chains_match_golden(traj_paths, golden):
    g = golden["terminal_traj_n16_T120_seed31"]
    rc, traj, ln = emu_propagate(traj_paths, np.ascontiguousarray(g["geo"]), 31, int(g["first"]), 120)
    assert rc == 0
    check_traj(traj, ln, g["traj"], g["len"])


def _oracle_screen(traj32, ln, tmax, thres_dist, thres_low):
    """Oracle screening of the fp32 trajectories the product wrote (a pure function of traj)."""
    from oracle import terminal as T
    n = traj32.shape[3]
    out = dict(hmd=[], vmd=[], tcpa=[], io=[], ii=[], enc=[], rw=[])
    for s in range(n):
        pair = []
        for ac in range(2):
            lo, hi = tmax - (int(ln[2 * ac + 1, s]) - 1), tmax + int(ln[2 * ac, s]) - 1
            pair.append(dict(t_s=np.arange(lo - tmax, hi - tmax + 1, dtype=np.float64), x_nm=traj32[0, ac, lo:hi + 1, s],
                             y_nm=traj32[1, ac, lo:hi + 1, s], z_ft=traj32[2, ac, lo:hi + 1, s]))
        h, v, t, io, ii, enc = T.get_generated_miss_distance(pair)
        c1, l1 = T.check_runway_proximity(pair[0], thres_dist, thres_low)
        c2, l2 = T.check_runway_proximity(pair[1], thres_dist, thres_low)
        for k, x in zip(("hmd", "vmd", "tcpa", "io", "ii", "enc", "rw"), (h, v, t, io, ii, enc, c1 + 2 * l1 + 4 * c2 + 8 * l2)):
            out[k].append(x)
    return {k: np.asarray(v) for k, v in out.items()}


def test_emu_screening_matches_oracle(traj_paths, golden):
    g = golden["terminal_traj_n16_T120_seed31"]
    rc, traj, ln = emu_propagate(traj_paths, np.ascontiguousarray(g["geo"]), 31, int(g["first"]), 120)
    assert rc == 0
    n = traj.shape[3]
    hmd, vmd = np.zeros(n), np.zeros(n)
    tcpa, enc, rw = np.zeros((3, n), dtype=np.int16), np.zeros(n, dtype=np.int16), np.zeros(n, dtype=np.uint8)
    thres_dist, thres_low = 3.0, 1500.0          # in the reference's own d_nm * 1.68781 "feet" (CorTerminalModel.m:197)
    H.emu_lib().emu_terminal_screen(traj.ctypes.data, ln.ctypes.data, n, 120.0, thres_dist, thres_low, hmd.ctypes.data,
                                    vmd.ctypes.data, tcpa.ctypes.data, enc.ctypes.data, rw.ctypes.data)
    want = _oracle_screen(traj, ln, 120, thres_dist, thres_low)
    assert np.array_equal(hmd, want["hmd"]) and np.array_equal(vmd, want["vmd"])          # fp64 on the same fp32 inputs
    assert np.array_equal(tcpa[0], want["tcpa"]) and np.array_equal(tcpa[1], want["io"]) and np.array_equal(tcpa[2], want["ii"])
    assert np.array_equal(enc, want["enc"]) and np.array_equal(rw, want["rw"])
    assert 0 < rw.astype(int).sum() and len(set(rw.tolist())) > 1                          # thresholds exercise both outcomes


def test_emu_unknown_intent_is_an_error(traj_paths, golden):
    geo = geo_from_golden(golden, 4)
    geo[0, 2] = 3.0                                   # own_intent = 3 has no ownship model (createEncounter.m:21-22)
    rc, _, _ = emu_propagate(traj_paths, geo, 1, 0, 10)
    assert rc == L.EMB_E_ARG


def test_bearing_cell_from_pseudo_angle_equals_atan2_route(traj_paths):
    """The chain kernel takes the bearing cell from a pseudo-angle instead of wrapTo360(atan2d(y, x)) (createEncounter.m:277,
    :293).  Both routes must give the same cell: 2e6 random points at all scales, the axes and diagonals, and points a hair to
    either side of every cutpoint direction; points closer than 1e-9 deg to a cutpoint are where last-bit rounding of either
    route decides, so only there may the cells differ (by one)."""
    lib = H.emu_lib()
    stem = SLOTS[0][0]
    m = H.EmuModel(traj_paths[stem])
    m.set_prior(1, L.EMB_PRIOR_STAY, 1.0)
    from oracle.em_read import em_read
    p = em_read(traj_paths[stem])
    ib = [k for k, lab in enumerate(p.labels_initial) if lab == '"bearing"'][0]
    cuts = np.asarray(p.boundaries[ib], dtype=np.float64)[1:-1]
    rng = np.random.default_rng(5)
    n = 2_000_000
    r = 10.0 ** rng.uniform(-6, 3, n)
    th = rng.uniform(0.0, 360.0, n)
    x, y = r * np.cos(np.radians(th)), r * np.sin(np.radians(th))
    ax = np.array([[1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1], [0, 0], [1, -0.0], [-1, -0.0],
                   [1e-300, -1e-320], [3, -1e-17]], dtype=np.float64)
    eps = np.concatenate([cuts - 1e-7, cuts + 1e-7, cuts - 1e-11, cuts + 1e-11])
    xe, ye = 2.5 * np.cos(np.radians(eps)), 2.5 * np.sin(np.radians(eps))
    x = np.ascontiguousarray(np.concatenate([x, ax[:, 0], xe]))
    y = np.ascontiguousarray(np.concatenate([y, ax[:, 1], ye]))
    got = np.zeros(x.size, dtype=np.int32)
    ref = np.zeros(x.size, dtype=np.int32)
    assert lib.emu_bearing_cells(m.h, x.size, x.ctypes.data, y.ctypes.data, got.ctypes.data, ref.ctypes.data, None, None, None) == 0
    bearing = np.mod(np.degrees(np.arctan2(y, x)), 360.0)
    near = np.min(np.abs(bearing[:, None] - cuts[None, :]), axis=1) < 1e-9
    assert np.array_equal(got[~near], ref[~near])
    assert np.all(np.abs(got[near] - ref[near]) <= 1)
    assert near.sum() < 200 and len(np.unique(got)) == len(cuts) + 1
    # and the reference route agrees with NumPy's own digitize on the same angles
    assert np.array_equal(ref[~near], np.searchsorted(cuts, bearing[~near], side="right"))


def test_distance_cell_from_the_squared_norm_is_exact(traj_paths):
    """The chain kernel takes the cell of d_nm = norm([x y]) (createEncounter.m:277) and the tests d_nm > dist_max, d_nm <= 0.25
    (:310-312) on x*x + y*y against thresholds min{s : sqrt(s) >= c}: identical to the square-root route on every point,
    including points placed within a few ulps of every cutpoint."""
    lib = H.emu_lib()
    m = H.EmuModel(traj_paths["intruder_transit_model"])
    m.set_prior(1, L.EMB_PRIOR_STAY, 1.0)
    from oracle.em_read import em_read
    p = em_read(traj_paths["intruder_transit_model"])
    idist = [k for k, lab in enumerate(p.labels_initial) if lab == '"distance"'][0]
    cuts = np.asarray(p.boundaries[idist], dtype=np.float64)[1:-1]
    rng = np.random.default_rng(6)
    n = 1_000_000
    r = np.concatenate([10.0 ** rng.uniform(-3, 1.2, n), rng.uniform(0.0, 9.0, n)])
    th = rng.uniform(0.0, 2 * np.pi, r.size)
    x, y = r * np.cos(th), r * np.sin(th)
    # points whose norm lands within a few ulps of a cutpoint (and of 0.25 and the upper bound), on axes and off them
    edge = np.concatenate([cuts, [0.25, float(p.boundaries[idist][-1])]])
    xe, ye = [], []
    for c in edge:
        for k in range(-6, 7):
            ck = c
            for _ in range(abs(k)):
                ck = np.nextafter(ck, np.inf if k > 0 else -np.inf)
            xe += [ck, 0.0, ck * 0.6, ck * np.cos(1.0)]
            ye += [0.0, -ck, ck * 0.8, ck * np.sin(1.0)]
    x = np.ascontiguousarray(np.concatenate([x, xe, [0.0]]))
    y = np.ascontiguousarray(np.concatenate([y, ye, [0.0]]))
    pc = np.zeros(x.size, dtype=np.int32)
    pr = np.zeros(x.size, dtype=np.int32)
    d2 = np.zeros(x.size, dtype=np.int32)
    dr = np.zeros(x.size, dtype=np.int32)
    near = np.zeros(2)
    assert lib.emu_bearing_cells(m.h, x.size, x.ctypes.data, y.ctypes.data, pc.ctypes.data, pr.ctypes.data, d2.ctypes.data,
                                 dr.ctypes.data, near.ctypes.data) == 0
    assert np.array_equal(d2, dr)
    d = np.sqrt(x * x + y * y)
    assert np.array_equal(dr, np.searchsorted(cuts, d, side="right"))
    s = x * x + y * y
    assert np.array_equal(s >= near[0], d > float(p.boundaries[idist][-1]))
    assert np.array_equal(s < near[1], d <= 0.25)
    assert len(np.unique(d2)) == len(cuts) + 1


def test_sind_cosd_and_constant_division_of_the_chain_kernel():
    """sincosd (degree reduction + fdlibm kernels) against libm within 2 ulp, exact at multiples of 90 degrees; div_const
    bit-identical to IEEE division."""
    lib = H.emu_lib()
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-360.0, 360.0, 2_000_000), rng.uniform(-1e-3, 1e-3, 1000), np.arange(-720.0, 721.0, 45.0),
                        [1e-300, -1e-300, 359.99999999999994, -359.99999999999994, 44.99999999999999, 45.00000000000001, 1000.5]])
    x = np.ascontiguousarray(x)
    s, c = np.zeros_like(x), np.zeros_like(x)
    lib.emu_sincosd(x.size, x.ctypes.data, s.ctypes.data, c.ctypes.data)
    r = np.fmod(x, 360.0)
    sr, cr = np.sin(np.radians(r)), np.cos(np.radians(r))
    # reference error budget: the radian conversion of the unreduced angle alone moves sin/cos by |a| * 2^-53 ~ 7e-16
    assert np.max(np.abs(s - sr)) < 1.5e-15 and np.max(np.abs(c - cr)) < 1.5e-15
    small = np.abs(r) < 1.0
    assert np.max(np.abs(s[small] - sr[small]) / np.maximum(np.abs(sr[small]), 1e-300)) < 4.5e-16
    m90 = np.fmod(r, 90.0) == 0.0
    q = np.round(r[m90] / 90.0).astype(int) % 4
    assert np.array_equal(s[m90], np.array([0.0, 1.0, 0.0, -1.0])[q]) and np.array_equal(c[m90], np.array([1.0, 0.0, -1.0, 0.0])[q])
    assert not np.any(np.signbit(s[m90] * 0.0 + s[m90]) & (s[m90] == 0.0)) and not np.any(np.signbit(c[m90]) & (c[m90] == 0.0))
    assert np.max(np.abs(s * s + c * c - 1.0)) < 5e-16
    a = np.concatenate([rng.uniform(-1e4, 1e4, 2_000_000), 10.0 ** rng.uniform(-12, 9, 1_000_000), np.arange(-36000.0, 36001.0), [0.0]])
    a = np.ascontiguousarray(a)
    qf, qh = np.zeros_like(a), np.zeros_like(a)
    lib.emu_div_const(a.size, a.ctypes.data, qf.ctypes.data, qh.ctypes.data)
    assert np.array_equal(qf, a / 6076.1154855643) and np.array_equal(qh, a / 100.0)


def test_stay_prior_is_required(traj_paths):
    lib = H.emu_lib()
    models = [H.EmuModel(traj_paths[stem]) for stem, _, _ in SLOTS]
    arr = (C.c_void_p * 10)(*[m.h.value for m in models])
    geo = np.zeros((12, 1))
    lim = (L.DynLimits * 2)(*[L.DynLimits(*LIMITS["GENERIC"])] * 2)
    rows = (C.c_int32 * 12)(*range(12))
    rc = lib.emu_terminal_propagate(arr, 1, 0, 1, geo.ctypes.data, 1, rows, 10.0, lim, 0, None, None)
    assert rc == L.EMB_E_ARG and b"stay prior" in lib.emu_last_error()


def test_abi_rejects_bad_arguments_without_gpu(traj_paths):
    lib = L.lib()
    lim = (L.DynLimits * 2)()
    assert lib.emb_dyn_limits_named(b"generic", C.byref(lim[0])) == 0 and lim[0].maxVel_ft_s == 506.0
    assert lib.emb_dyn_limits_named(b"RTCA228_A2", C.byref(lim[1])) == 0 and lim[1].maxTurnRate_deg_s == 3.0
    assert lib.emb_dyn_limits_named(b"nope", C.byref(lim[1])) == L.EMB_E_ARG
    assert lib.emb_terminal_traj_len(10, 120.0) == 5 * 2 * 241 * 10
    assert lib.emb_terminal_traj_len(10, -1.0) == 0


# ---------------------------------------------------------------------------------------------------
def _product_model(model_paths, traj_paths, types=("GENERIC", "GENERIC")):
    from em_model_manned_bayes_b200.model import CorTerminalModel
    m = CorTerminalModel(model_paths["terminal_v3_radar_encounter_model"], acType1=types[0], acType2=types[1])
    m.load_trajectory_models(os.path.dirname(traj_paths[TRAJECTORY_STEMS[0]]))
    return m


@pytest.mark.gpu
def test_gpu_chains_match_golden(model_paths, traj_paths, golden):
    g = golden["terminal_traj_n16_T120_seed31"]
    m = _product_model(model_paths, traj_paths)
    res = m.create_encounters(np.ascontiguousarray(g["geo"]), 120, seed=31, first_sample=int(g["first"]), geo_rows=range(12))
    check_traj(res.traj, res.len, g["traj"], g["len"])
    enc = res.encounter(3)
    assert enc[0]["t_s"][0] == -(int(g["len"][1, 3]) - 1) and enc[0]["t_s"][-1] == int(g["len"][0, 3]) - 1


@pytest.mark.gpu
def test_gpu_chains_match_live_oracle_mixed_limits(model_paths, traj_paths, golden):
    """More encounters than the goldens hold, oracle run live (about 0.15 s per encounter): all 64 golden geometry
    samples under GENERIC limits and 48 under (RTCA228_A2-like narrow turn / TEST altitude) limits, 64-bit encounter ids."""
    geo = geo_from_golden(golden, 64)
    m = _product_model(model_paths, traj_paths)
    res = m.create_encounters(geo, 120, seed=41, first_sample=2 ** 40 + 3, geo_rows=range(12))
    want, want_len = oracle_propagate(traj_paths, geo, 41, 2 ** 40 + 3, 120)
    check_traj(res.traj, res.len, want, want_len)
    geo2 = geo_from_golden(golden, 48)
    geo2[5] = np.clip(geo2[5], 70.0, 180.0)
    geo2[11] = np.clip(geo2[11], 70.0, 180.0)
    m2 = _product_model(model_paths, traj_paths, ("RTCA228_A3", "TEST"))
    res2 = m2.create_encounters(geo2, 75, seed=42, geo_rows=range(12))
    want2, want_len2 = oracle_propagate(traj_paths, geo2, 42, 0, 75, ("RTCA228_A3", "TEST"))
    check_traj(res2.traj, res2.len, want2, want_len2)


@pytest.mark.gpu
def test_gpu_chains_equal_the_host_emulation_bit_for_bit(model_paths, traj_paths, tmp_path):
    """The chain code is one source for the device and the host emulation, and everything in it is now exactly specified
    arithmetic (own sind/cosd, correctly rounded division and square root, no library transcendental): trajectories, NaN
    pattern and chain lengths must be IDENTICAL, on 3 000 sampled encounters and on the models whose speed edges sit on the
    dynamic limits (the cells that depend on the last bit of norm(R v), test_speed_edges_on_the_dynamic_limits)."""
    from em_model_manned_bayes_b200.model import CorTerminalModel
    m = _product_model(model_paths, traj_paths, ("RTCA228_A3", "GENERIC"))
    vals, _, _ = m.sample_raw(3000, seed=4)
    labels = [l.strip('"') for l in m.labels_initial]
    fields = ("own_intent", "own_distance", "own_bearing", "own_alt", "own_heading", "own_speed",
              "int_intent", "int_distance", "int_bearing", "int_alt", "int_heading", "int_speed")
    geo = np.ascontiguousarray(np.stack([np.asarray(vals)[:, labels.index(f)] for f in fields]))
    geo[5] = np.clip(geo[5], 70.0, 180.0)
    for paths, mm, tmax in ((traj_paths, m, 120),):
        res = mm.create_encounters(geo, tmax, seed=9, first_sample=2 ** 33 + 1, geo_rows=range(12))
        rc, traj, ln = emu_propagate(paths, geo, 9, 2 ** 33 + 1, tmax, ("RTCA228_A3", "GENERIC"))
        assert rc == 0
        assert np.array_equal(np.asarray(res.len), ln)
        assert np.array_equal(np.asarray(res.traj).view(np.uint32), traj.view(np.uint32))
    edges = [0, 50, 68, 100, 169, 186, 338, 491, 506, 600]
    epaths = write_terminal_model_set(str(tmp_path / "edge_models"), seed=7, speed_edges=edges)
    me = CorTerminalModel(model_paths["terminal_v3_radar_encounter_model"], acType1="RTCA228_A3", acType2="TEST")
    me.load_trajectory_models(os.path.dirname(epaths[TRAJECTORY_STEMS[0]]))
    g2 = geo[:, :600].copy()
    g2[11] = np.clip(g2[11], 70.0, 180.0)
    res = me.create_encounters(g2, 90, seed=41, geo_rows=range(12))
    rc, traj, ln = emu_propagate(epaths, g2, 41, 0, 90, ("RTCA228_A3", "TEST"))
    assert rc == 0
    assert np.array_equal(np.asarray(res.len), ln)
    assert np.array_equal(np.asarray(res.traj).view(np.uint32), traj.view(np.uint32))


@pytest.mark.gpu
def test_gpu_limits_golden(model_paths, traj_paths, golden):
    g = golden["terminal_traj_n12_T45_seed32_test_a1"]
    m = _product_model(model_paths, traj_paths, ("TEST", "RTCA228_A1"))
    res = m.create_encounters(np.ascontiguousarray(g["geo"]), 45, seed=32, first_sample=int(g["first"]), geo_rows=range(12))
    check_traj(res.traj, res.len, g["traj"], g["len"])


@pytest.mark.gpu
def test_gpu_pipeline_properties_and_shard_invariance(model_paths, traj_paths):
    """geometry sampling -> chains, all on the device; 20 000 encounters; any shard reproduces the same numbers."""
    import torch
    m = _product_model(model_paths, traj_paths)
    n, tmax = 20000, 120
    vals, _, _ = m.sample_raw(n, seed=5, device="cuda:0")              # (n, 15) view of a (15, n) device tensor
    geo = vals.T.contiguous()
    res = m.create_encounters(geo, tmax, seed=6, device="cuda:0")
    traj, ln = res.traj.cpu().numpy(), res.len.cpu().numpy()
    assert ln.min() >= 1 and ln.max() <= tmax + 1
    occ = ~np.isnan(traj[0])                                            # (2, S, n)
    for ac in range(2):
        assert np.array_equal(occ[ac].sum(axis=0), ln[2 * ac] + ln[2 * ac + 1] - 1)
        assert np.array_equal(np.isnan(traj[:, ac]), np.broadcast_to(~occ[ac], traj[:, ac].shape))
    with np.errstate(invalid="ignore"):
        dz = np.abs(np.diff(traj[2], axis=1))
        assert np.nanmax(dz) <= 100.0 + 2e-3                            # maxVertRate_ft_s, GENERIC (fp32 outputs near 1e4 ft)
        assert np.nanmax(traj[4]) <= max(506.0, float(geo[8].max()), float(geo[13].max())) * (1 + 1e-6)
        hd = traj[3]
        assert np.nanmin(hd) >= 0.0 and np.nanmax(hd) <= 360.0
        assert np.nanmax(np.hypot(traj[0], traj[1])) <= 8.0 + 506.0 / 6076.0 + 1e-3   # one step beyond bounds_initial(dist)
    # shard [7000, 7000+500) alone
    part = m.create_encounters(geo[:, 7000:7500].contiguous(), tmax, seed=6, first_sample=7000, device="cuda:0")
    assert torch.equal(part.len, res.len[:, 7000:7500])
    a, b = part.traj.cpu().numpy(), traj[:, :, :, 7000:7500]
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    # host-memory call gives the same bytes as the device-memory call
    host = m.create_encounters(geo[:, :300].cpu().numpy(), tmax, seed=6)
    assert np.array_equal(np.asarray(host.len), ln[:, :300])
    assert np.array_equal(np.nan_to_num(np.asarray(host.traj), nan=-1.0), np.nan_to_num(traj[:, :, :, :300], nan=-1.0))


@pytest.mark.gpu
def test_gpu_screening_matches_oracle(model_paths, traj_paths, golden):
    import torch
    m = _product_model(model_paths, traj_paths)
    n, tmax = 3000, 120
    vals, _, _ = m.sample_raw(n, seed=15, device="cuda:0")
    res = m.create_encounters(vals.T.contiguous(), tmax, seed=16, device="cuda:0")
    sc = m.screen_encounters(res, thresDist_ft=3.0, thresAltLow_ft=1500.0, device="cuda:0")
    torch.cuda.synchronize()
    traj, ln = res.traj.cpu().numpy(), res.len.cpu().numpy()
    k = 400                                                   # oracle on the first 400 encounters
    want = _oracle_screen(traj[:, :, :, :k], ln[:, :k], tmax, 3.0, 1500.0)
    assert np.array_equal(sc["hmd_ft"].cpu().numpy()[:k], want["hmd"]) and np.array_equal(sc["vmd_ft"].cpu().numpy()[:k], want["vmd"])
    assert np.array_equal(sc["tcpa_s"].cpu().numpy()[:k], want["tcpa"])
    assert np.array_equal(sc["tcpa_index_own"].cpu().numpy()[:k], want["io"]) and np.array_equal(sc["tcpa_index_int"].cpu().numpy()[:k], want["ii"])
    assert np.array_equal(sc["enc_time_s"].cpu().numpy()[:k], want["enc"])
    rw = (sc["is_close1"].to(torch.uint8) + 2 * sc["is_low1"].to(torch.uint8) + 4 * sc["is_close2"].to(torch.uint8)
          + 8 * sc["is_low2"].to(torch.uint8)).cpu().numpy()
    assert np.array_equal(rw[:k], want["rw"])
    # size-independent properties on all encounters: the CPA is a common time, hmd is the minimum over them
    t = sc["tcpa_s"].cpu().numpy().astype(int)
    own_lo, own_hi = -(ln[1].astype(int) - 1), ln[0].astype(int) - 1
    int_lo, int_hi = -(ln[3].astype(int) - 1), ln[2].astype(int) - 1
    assert np.all(t >= np.maximum(own_lo, int_lo)) and np.all(t <= np.minimum(own_hi, int_hi))
    assert np.array_equal(sc["enc_time_s"].cpu().numpy(), np.minimum(own_hi, int_hi) - np.maximum(own_lo, int_lo) + 1)
    d0 = np.hypot(traj[0, 0, tmax] - traj[0, 1, tmax], traj[1, 0, tmax] - traj[1, 1, tmax]).astype(np.float64) * 6076.1154855643
    assert np.all(sc["hmd_ft"].cpu().numpy() <= d0 * (1 + 1e-6) + 1e-3)
    # host-memory call == device-memory call
    host = type(res)(n=64, tmax=tmax, traj=np.ascontiguousarray(traj[:, :, :, :64]), len=np.ascontiguousarray(ln[:, :64]))
    sh = m.screen_encounters(host, thresDist_ft=3.0, thresAltLow_ft=1500.0)
    assert np.array_equal(sh["hmd_ft"], sc["hmd_ft"].cpu().numpy()[:64]) and np.array_equal(sh["tcpa_s"], t[:64])


@pytest.mark.gpu
def test_gpu_unknown_intent_and_missing_models(model_paths, traj_paths):
    from em_model_manned_bayes_b200.model import CorTerminalModel
    m = _product_model(model_paths, traj_paths)
    geo = np.array([[3.0], [3.0], [200.0], [1200.0], [45.0], [180.0], [1.0], [4.0], [100.0], [2200.0], [300.0], [250.0]])
    with pytest.raises(L.EmbError) as e:
        m.create_encounters(geo, 10, seed=1, geo_rows=range(12))
    assert e.value.code == L.EMB_E_ARG and "Unknown int_intent" in e.value.message
    bare = CorTerminalModel(model_paths["terminal_v3_radar_encounter_model"])
    with pytest.raises(L.EmbError):
        bare.create_encounters(geo, 10, seed=1, geo_rows=range(12))
