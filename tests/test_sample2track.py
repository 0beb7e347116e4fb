"""First-order track integration: sample2track.m:188-244.

CPU: host emulation of the device routine (emb_integrate.cuh) against the oracle restatement (oracle/track.py) on
oracle-sampled tracks.  GPU: the C ABI on GPU-sampled tracks, device- and host-memory, and the file-to-file driver
em_sample -> sample2track against the oracle's own text pipeline (CSV rows are integer feet)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import helpers as H
from em_model_manned_bayes_b200 import _lib as L
from em_model_manned_bayes_b200.sample2track import FT_PER_NM, make_valid_name
from oracle import track as OT
from oracle.drivers import dbn_tracks, em_sample_text
from oracle.em_read import em_read
from oracle.uniforms import KeyedPhilox


def tiles_from_samples(vals):
    """(n, n_tv, T) -> [ceil(T/4)][ceil(n/128)][n_tv][128][4] float32 (the dense layout of emb_sample_tracks), written
    element by element from the layout's definition (emb200.h: emb_track_out)."""
    n, ntv, T = vals.shape
    nch, ntile = (T + 3) // 4, (n + 127) // 128
    out = np.zeros((nch, ntile, ntv, 128, 4), dtype=np.float32)
    s = np.arange(n)
    for g in range(ntv):
        for t in range(T):
            out[t // 4, s // 128, g, s % 128, t % 4] = vals[:, g, t]
    return out.ravel()


def oracle_tracks(p, out, T, from_fp32=True):
    """Oracle integration of oracle samples; the rates pass through fp32 like the dense output they are read from."""
    lab = [make_valid_name(l) for l in p.labels_initial]
    iL, iv, ia, ih, it = (lab.index(k) for k in ("L", "v", "dotV", "dotH", "dotPsi"))
    b = p.boundaries[iv]
    res = []
    for s in out:
        f = (lambda a: a.astype(np.float32).astype(np.float64)) if from_fp32 else (lambda a: a)
        res.append(OT.integrate(s.samples[iL, 0], s.samples[iv, 0], f(s.samples[ia, :T]), f(s.samples[ih, :T]),
                                f(s.samples[it, :T]), b[0], b[-1]))
    return res, (iL, iv, ia, ih, it)


def check_xyz(xyz, good, ref):
    for k, (t, x, y, z, ok) in enumerate(ref):
        for f, w in enumerate((x, y, z)):
            g = np.asarray(xyz[f, :, k], dtype=np.float64)
            assert np.all(np.abs(g - w) <= 2e-6 * np.abs(w) + 0.02), (k, f, np.abs(g - w).max())
        assert bool(good[k]) == ok


def test_make_valid_name_matches_the_reference_defaults():
    assert [make_valid_name(l) for l in ('"G"', '"\\dot v"', '"\\dot h"', '"\\dot \\psi"')] == ["G", "dotV", "dotH", "dotPsi"]
    assert make_valid_name('"\\dot v(t+1)"') == "dotV_t_1_" and make_valid_name('"\\dot \\psi(t+1)"') == "dotPsi_t_1_"   # :39-41
    assert abs(FT_PER_NM - 6076.1154855643) < 1e-9


def test_oracle_integration_known_answer():
    # straight and level: 100 kt for 3 s -> 168.78 ft per second along x
    t, x, y, z, ok = OT.integrate(1000.0, 100.0, [0, 0, 0], [0, 0, 0], [0, 0, 0], 50.0, 300.0)
    assert np.allclose(x, np.arange(4) * 100 * FT_PER_NM / 3600) and np.all(y == 0) and np.all(z == 1000.0) and ok
    # 90 deg/s turn: second step goes along +y (cosd(90) == 0 exactly); descent of 60000 ft/min hits the ground -> CFIT
    t, x, y, z, ok = OT.integrate(1000.0, 100.0, [0, 0], [0, -60000.0], [90.0, 0], 50.0, 300.0)
    assert x[2] == x[1] and y[2] > 0 and z[2] == 0.0 and ok
    assert not OT.integrate(1000.0, 100.0, [0, 0], [0, -60001.0], [0, 0], 50.0, 300.0)[4]
    assert not OT.integrate(1000.0, 100.0, [250.0], [0], [0], 50.0, 300.0)[4]          # speed >= max -> rejected (:240)


@pytest.mark.parametrize("model,n,T", [("uncor_1200code_v2p1", 24, 120), ("glider_v1", 16, 61)])
def test_emu_integration_matches_oracle(model_paths, model, n, T):
    p = em_read(model_paths[model])
    out = dbn_tracks(p, n, T, KeyedPhilox(55))
    ref, (iL, iv, ia, ih, it) = oracle_tracks(p, out, T)
    tm = [int(v) for v in np.asarray(p.temporal_map)[:, 0]]
    rates = np.asarray(p.resample_rates)
    tv = sorted(set(tm) | {i + 1 for i in range(p.n_initial) if rates[i] > 0})
    vals = np.stack([s.samples[[v - 1 for v in tv], :T] for s in out])
    tiles = tiles_from_samples(vals)
    init = np.ascontiguousarray(np.stack([s.samples[:, 0] for s in out]).T)
    xyz = np.zeros((3, T + 1, n), dtype=np.float32)
    good = np.zeros(n, dtype=np.uint8)
    b = p.boundaries[iv]
    ur = FT_PER_NM / 3600.0
    rc = H.emu_lib().emu_tracks_integrate(n, T, iL, iv, tv.index(ia + 1), tv.index(ih + 1), tv.index(it + 1), len(tv), ur, 1.0 / 60.0, 1.0,
                                          float(b[0]) * ur, float(b[-1]) * ur, init.ctypes.data, tiles.ctypes.data,
                                          xyz.ctypes.data, good.ctypes.data)
    assert rc == 0
    check_xyz(xyz, good, ref)


def test_abi_rejects_models_without_the_variables(model_paths):
    from em_model_manned_bayes_b200.model import EncounterModel
    m = EncounterModel(model_paths["balloon_v1"])
    o = L.IntegrateOpts()
    o.idx_altitude, o.idx_speed, o.idx_acceleration, o.idx_vertrate, o.idx_turnrate = 1, 2, 1, 2, 2   # variable 1 never varies in time
    dummy = np.zeros(64)
    rc = L.lib().emb_tracks_integrate(m._h, 1, 4, dummy.ctypes.data, dummy.ctypes.data, C.byref(o), None, None)
    assert rc == L.EMB_E_ARG


# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_integration_matches_oracle_and_host_equals_device(model_paths):
    import torch
    from em_model_manned_bayes_b200.model import EncounterModel
    from em_model_manned_bayes_b200.sample2track import integrate_tracks
    model, n, T = "uncor_1200code_v2p1", 40, 150
    p = em_read(model_paths[model])
    out = dbn_tracks(p, n, T, KeyedPhilox(56))
    ref, _ = oracle_tracks(p, out, T, from_fp32=False)
    m = EncounterModel(model_paths[model])
    res = m.sample_tracks(n, T, seed=56, device="cuda:0")
    xyz, good = integrate_tracks(m, res, device="cuda:0")
    torch.cuda.synchronize()
    check_xyz(xyz.cpu().numpy(), good.cpu().numpy(), ref)
    host = m.sample_tracks(n, T, seed=56)
    xyz_h, good_h = integrate_tracks(m, host)
    assert np.array_equal(xyz_h, xyz.cpu().numpy()) and np.array_equal(good_h, good.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("model,uncor,n,T", [("uncor_1200code_v2p1", True, 3000, 150), ("uncor_allcode_fwsingle_v1", True, 2000, 601),
                                             ("glider_v1", True, 1500, 90), ("cor_v1", False, 0, 0)])
def test_gpu_fused_xyz_equals_sample_then_integrate(model_paths, model, uncor, n, T):
    """emb_sample_tracks_xyz (the Euler loop of sample2track.m:199-244 fused into the track kernel) against emb_sample_tracks followed
    by emb_tracks_integrate: points, is_good and every dense output identical, on the fast branch, the per-step branch, with the
    rejection loop of UncorEncounterModel.sample, with host buffers, against the oracle, and through the two-pass fallback of a
    model shape without a specialised kernel (forced generic)."""
    import torch
    from em_model_manned_bayes_b200.model import EncounterModel, UncorEncounterModel
    from em_model_manned_bayes_b200.sample2track import integrate_tracks, sample_tracks_xyz
    if not uncor:
        pytest.skip("cor_v1 has no altitude-layer / airspeed pair to integrate")
    m = UncorEncounterModel(model_paths[model])
    opts = lambda: m.uncor_opts()
    res = m.sample_tracks(n, T, seed=77, first_sample=123, opts=opts(), device="cuda:0")
    xyz, good = integrate_tracks(m, res, device="cuda:0")
    fx, fg = sample_tracks_xyz(m, n, T, seed=77, first_sample=123, sample_opts=opts(), device="cuda:0")
    torch.cuda.synchronize()
    assert L.lib().emb_debug_last_kernel_fast() == 1
    assert torch.equal(fx, xyz) and torch.equal(fg, good)
    assert 0 < int(good.sum()) < n
    # with the dense outputs requested as well
    dense = m.sample_tracks(1, T, seed=0, opts=opts(), device="cuda:0")      # a TrackResult to size the buffers from
    dense = m.sample_tracks(n, T, seed=1, opts=opts(), device="cuda:0")
    fx2, fg2 = sample_tracks_xyz(m, n, T, seed=77, first_sample=123, sample_opts=opts(), device="cuda:0", dense=dense)
    torch.cuda.synchronize()
    assert torch.equal(fx2, xyz) and torch.equal(fg2, good)
    for a, b in ((dense.bins_tiled, res.bins_tiled), (dense.values_tiled, res.values_tiled), (dense.init_bins, res.init_bins),
                 (dense.init_values, res.init_values), (dense.attempts, res.attempts)):
        assert torch.equal(a, b)
    # host buffers
    hx, hg = sample_tracks_xyz(m, 257, T, seed=77, first_sample=123, sample_opts=opts())
    assert np.array_equal(hx, xyz[:, :, :257].cpu().numpy()) and np.array_equal(hg, good[:257].cpu().numpy())
    # is_good only
    _, g3 = sample_tracks_xyz(m, n, T, seed=77, first_sample=123, sample_opts=opts(), device="cuda:0", want_xyz=False)
    assert torch.equal(g3, good)
    # a shape without a specialised kernel: the library runs the two passes itself (the generic kernel's fp32 values may differ
    # from the specialised kernel's in the last bit, so the reference is the generic two-pass route)
    L.lib().emb_debug_force_generic(1)
    try:
        gx, gg = sample_tracks_xyz(m, 500, T, seed=77, first_sample=123, sample_opts=opts(), device="cuda:0")
        rx, rg = integrate_tracks(m, m.sample_tracks(500, T, seed=77, first_sample=123, opts=opts(), device="cuda:0"), device="cuda:0")
        torch.cuda.synchronize()
        assert L.lib().emb_debug_last_kernel_fast() == 0
    finally:
        L.lib().emb_debug_force_generic(0)
    assert torch.equal(gx, rx) and torch.equal(gg, rg)


@pytest.mark.gpu
def test_gpu_fused_xyz_matches_oracle(model_paths):
    from em_model_manned_bayes_b200.model import EncounterModel
    from em_model_manned_bayes_b200.sample2track import sample_tracks_xyz
    model, n, T = "uncor_1200code_v2p1", 40, 150
    p = em_read(model_paths[model])
    out = dbn_tracks(p, n, T, KeyedPhilox(56))
    ref, _ = oracle_tracks(p, out, T, from_fp32=False)
    m = EncounterModel(model_paths[model])
    xyz, good = sample_tracks_xyz(m, n, T, seed=56)
    check_xyz(xyz, good, ref)


@pytest.mark.gpu
def test_gpu_integration_full_size_properties(model_paths):
    """1e5 tracks x 600 s in HBM: integration invariants that do not need the oracle."""
    import torch
    from em_model_manned_bayes_b200.model import UncorEncounterModel
    from em_model_manned_bayes_b200.sample2track import integrate_opts, integrate_tracks
    m = UncorEncounterModel(model_paths["uncor_allcode_fwsingle_v1"])
    n, T = 100_000, 600
    res = m.sample_compact(n, T, seed=3, device="cuda:0")
    o = integrate_opts(m)
    xyz, good = integrate_tracks(m, res, opts=o, device="cuda:0")
    torch.cuda.synchronize()
    assert bool(torch.isfinite(xyz).all())
    assert bool((xyz[0, 0] == 0).all()) and bool((xyz[1, 0] == 0).all())
    iv = res.init_values
    assert torch.equal(xyz[2, 0], iv[o.idx_altitude - 1].to(torch.float32))
    # z(t) - z(0) == sum of vertical rates / 60 (fp64 accumulation in the kernel, fp32 outputs)
    tvk = res.tv_vars.index(o.idx_vertrate)
    dz = res.values[:, tvk, :].to(torch.float64).sum(dim=1) / 60.0
    got = (xyz[2, T].to(torch.float64) - iv[o.idx_altitude - 1])
    assert float((got - dz).abs().max()) < 0.05
    # step length == speed of the previous second: |dxy| <= max_speed for good tracks
    step = torch.hypot(xyz[0, 1:] - xyz[0, :-1], xyz[1, 1:] - xyz[1, :-1])
    g = good.bool()
    assert float(step[:, g].max()) < o.max_speed + 0.1 and float(step[:, g].min()) > o.min_speed - 0.1
    assert bool((xyz[2][:, g] >= 0).all())
    assert 0.2 < float(g.float().mean()) <= 1.0


@pytest.mark.gpu
def test_file_pipeline_em_sample_to_sample2track(model_paths, tmp_path):
    """RUN_1_emsample -> RUN_2_sample2track: the reference's text files in, CSV tracks out, against the oracle fed its own
    %g-rounded text (rows are integer feet; a last-digit rounding tie may differ by one foot)."""
    from em_model_manned_bayes_b200.em_sample import em_sample
    from em_model_manned_bayes_b200.sample2track import sample2track
    model, n, T = "uncor_1200code_v2p1", 10, 80
    fi, ft = str(tmp_path / "initial.txt"), str(tmp_path / "transition.txt")
    em_sample(model_paths[model], fi, ft, num_initial_samples=n, num_transition_samples=T, rng_seed=42)
    out_dir = str(tmp_path / "tracks")
    is_good, Tc = sample2track(model_paths[model], fi, ft, out_dir_parent=out_dir)
    p = em_read(model_paths[model])
    ti, tt = em_sample_text(p, n, T, KeyedPhilox(42))
    A = np.array([[float(x) for x in l.split()] for l in ti.splitlines()[1:]])
    B = np.array([[float(x) for x in l.split()] for l in tt.splitlines()[1:]])
    lab = [make_valid_name(l) for l in p.labels_initial]
    iL, iv = lab.index("L"), lab.index("v")
    tm = [int(v) for v in np.asarray(p.temporal_map)[:, 0]]
    col = {name: 2 + tm.index(lab.index(name) + 1) for name in ("dotV", "dotH", "dotPsi")}
    b = p.boundaries[iv]
    files = glob.glob(os.path.join(out_dir, "**", "*.csv"), recursive=True)
    assert len(files) == int(is_good.sum())
    for k in range(n):
        rows = B[B[:, 0] == k + 1]
        t, x, y, z, ok = OT.integrate(A[k, 1 + iL], A[k, 1 + iv], rows[:, col["dotV"]], rows[:, col["dotH"]], rows[:, col["dotPsi"]],
                                      b[0], b[-1])
        assert ok == bool(is_good[k])
        if not ok:
            continue
        f = [q for q in files if "_id%d_" % (k + 1) in os.path.basename(q)]
        assert len(f) == 1 and os.path.basename(f[0]).startswith("BAYES_t%d_id%d_alt%d_" % (T, k + 1, round(z[0])))
        assert "G%d" % int(A[k, 1]) in f[0] and "A%d" % int(A[k, 2]) in f[0]
        got = np.loadtxt(f[0], delimiter=",", skiprows=1)
        assert open(f[0]).readline().strip() == "time_s,x_ft,y_ft,z_ft"
        assert np.array_equal(got[:, 0], t)
        assert np.all(np.abs(got[:, 1:] - np.stack([x, y, z], axis=1)) <= 1.0 + 2e-5 * np.abs(np.stack([x, y, z], axis=1)))
