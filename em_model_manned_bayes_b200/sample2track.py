"""sample2track.m:1-300 -- first-order tracks from sampled initial conditions and per-second rates, on top of the C ABI.

  integrate_tracks(model, res)                  the per-track loop of sample2track.m:188-244 on a TrackResult that is
                                                still in HBM (or in host memory): (xyz (3, T+1, n) float32, is_good (n,))
  sample2track(parameters_filename, initial_filename, transition_filename, ...)
                                                the reference's file-to-file driver: reads the two text files em_sample
                                                wrote, integrates on the GPU, writes BAYES_t*_id*_alt*_speed*.csv into
                                                the same G*/A*/<alt>ft directory tree (:143-178, :246-290)

Differences from the reference, all outside the integration itself: `randperm` (when the initial file holds more than
num_max_tracks tracks, :77-79) uses numpy's generator seeded with rng_seed, and plotting (:220-232) is not offered."""
from __future__ import annotations

import ctypes as C
import math
import os
import re
from typing import Optional

import numpy as np

from . import _lib as L
from .model import EncounterModel, TrackResult, _ptr, tile

FT_PER_NM = 1852.0 / 0.3048           # unitsratio('ft', 'nm')


def make_valid_name(label: str) -> str:
    """matlab.lang.makeValidName(erase(label, {'"', '\\'})) for the labels of the shipped models
    ('"\\dot v"' -> 'dotV', '"\\dot h(t+1)"' -> 'dotH_t_1_')."""
    s = label.replace('"', "").replace("\\", "")
    s = re.sub(r"\s+([a-zA-Z])", lambda m: m.group(1).upper(), s.strip())   # whitespace removed, next letter upper-cased
    s = re.sub(r"\s+", "", s)
    s = re.sub(r"[^A-Za-z0-9_]", "_", s)
    if not re.match(r"[A-Za-z]", s):
        s = "x" + s
    return s


def _find(names, name):
    return names.index(name) + 1 if name in names else 0


def integrate_opts(model: EncounterModel, label_initial_altitude="L", label_initial_speed="v",
                   label_initial_acceleration="dotV", label_initial_vertrate="dotH", label_initial_turnrate="dotPsi"):
    """sample2track.m:81-141: column lookup, unit ratios and speed bounds."""
    names = [make_valid_name(l) for l in model.labels_initial]
    o = L.IntegrateOpts()
    o.idx_altitude, o.idx_speed = _find(names, label_initial_altitude), _find(names, label_initial_speed)
    o.idx_acceleration = _find(names, label_initial_acceleration)
    o.idx_vertrate, o.idx_turnrate = _find(names, label_initial_vertrate), _find(names, label_initial_turnrate)
    if not all((o.idx_altitude, o.idx_speed, o.idx_acceleration, o.idx_vertrate, o.idx_turnrate)):
        raise L.EmbError(L.EMB_E_ARG, "sample2track: the model lacks one of the variables L, v, dotV, dotH, dotPsi")
    o.ur_speed, o.ur_vertrate, o.ur_heading = FT_PER_NM / 3600.0, 1.0 / 60.0, 1.0             # :108-125
    b = model.boundaries[o.idx_speed - 1]
    o.min_speed, o.max_speed = float(b[0]) * o.ur_speed, float(b[-1]) * o.ur_speed           # :98-99, :140-141
    o.mem, o.device, o.stream = L.EMB_MEM_HOST, -1, None
    return o


def integrate_tracks(model: EncounterModel, res: TrackResult, opts: Optional[L.IntegrateOpts] = None, device=None,
                     want_xyz: bool = True):
    """sample2track.m:188-244 for every track of `res` (dense compact output of sample_tracks).
    Returns (xyz (3, T+1, n) float32: x_ft, y_ft, z_ft at time_s = 0..T, is_good (n,) uint8)."""
    o = opts if opts is not None else integrate_opts(model)
    n, T = res.n, res.T
    if device is not None:
        import torch
        dev = torch.device(device)
        o.mem, o.device = L.EMB_MEM_DEVICE, dev.index if dev.index is not None else torch.cuda.current_device()
        o.stream = torch.cuda.current_stream(dev).cuda_stream
        xyz = torch.empty((3, T + 1, n), dtype=torch.float32, device=dev) if want_xyz else None
        good = torch.empty((n,), dtype=torch.uint8, device=dev)
    else:
        xyz = np.empty((3, T + 1, n), dtype=np.float32) if want_xyz else None
        good = np.empty((n,), dtype=np.uint8)
    L.check(L.lib().emb_tracks_integrate(model._h, n, T, _ptr(res.init_values), _ptr(res.values_tiled), C.byref(o),
                                         _ptr(xyz), _ptr(good)))
    return xyz, good


def sample_tracks_xyz(model: EncounterModel, n: int, T: int, seed: int = 0, first_sample: int = 0, sample_opts=None,
                      opts: Optional[L.IntegrateOpts] = None, device=None, want_xyz: bool = True, dense: Optional[TrackResult] = None):
    """Sampling and the Euler loop of sample2track.m:188-244 in ONE kernel pass (emb200.h: emb_sample_tracks_xyz): the same
    (xyz, is_good) as `integrate_tracks(model, model.sample_tracks(...))` without the dense tiles' round trip through HBM.
    `sample_opts`: e.g. `UncorEncounterModel.uncor_opts()`; `dense`: a TrackResult whose buffers are filled as well."""
    io = opts if opts is not None else integrate_opts(model)
    so = sample_opts if sample_opts is not None else model._opts()
    if device is not None:
        import torch
        dev = torch.device(device)
        so.mem, so.device = L.EMB_MEM_DEVICE, dev.index if dev.index is not None else torch.cuda.current_device()
        so.stream = torch.cuda.current_stream(dev).cuda_stream
        xyz = torch.empty((3, T + 1, n), dtype=torch.float32, device=dev) if want_xyz else None
        good = torch.empty((n,), dtype=torch.uint8, device=dev)
    else:
        xyz = np.empty((3, T + 1, n), dtype=np.float32) if want_xyz else None
        good = np.empty((n,), dtype=np.uint8)
    to = None
    if dense is not None:
        to = L.TrackOut(_ptr(dense.bins_tiled), _ptr(dense.values_tiled), _ptr(dense.init_bins), _ptr(dense.init_values),
                        _ptr(dense.attempts), None, None)
    rng = L.Rng(int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_sample))
    L.check(L.lib().emb_sample_tracks_xyz(model._h, C.byref(rng), n, T, C.byref(so), C.byref(io),
                                          C.byref(to) if to is not None else None, _ptr(xyz), _ptr(good)))
    return xyz, good


def _read_table(path):
    with open(path, encoding="utf-8") as f:
        f.readline()
        return np.loadtxt(f, ndmin=2)


def sample2track(parameters_filename: str, initial_filename: str, transition_filename: str, num_max_tracks: int = 10000,
                 out_dir_parent: str = "output/tracks", isOverwriteZeroBoundaries: bool = False, idxZeroBoundaries=(1, 2, 3),
                 rng_seed: int = 42, label_initial_geographic="G", label_initial_airspace="A", **labels):
    """-> (is_good (num_tracks,) bool, T_initial (num_tracks, 1 + n_initial) with the unit conversions of :128-131)."""
    model = EncounterModel(parameters_filename, idxZeroBoundaries=idxZeroBoundaries,
                           isOverwriteZeroBoundaries=isOverwriteZeroBoundaries)
    o = integrate_opts(model, **labels)
    names_init = [make_valid_name(l) for l in model.labels_initial]
    dyn_t = [int(v) for v in model.temporal_map[:, 0]]
    Ti = _read_table(initial_filename)                                    # id, initial variables
    Tt = _read_table(transition_filename)                                 # id, t, dynamic variables (temporal_map order)
    if Ti.shape[0] > num_max_tracks:                                      # :77-79
        keep = np.sort(np.random.default_rng(rng_seed).permutation(Ti.shape[0])[:num_max_tracks])
        Ti = Ti[keep]
    n = Ti.shape[0]
    ids = Ti[:, 0].astype(np.int64)
    order = np.argsort(Tt[:, 0], kind="stable")
    Tt = Tt[order]
    first = np.searchsorted(Tt[:, 0], ids, side="left")
    last = np.searchsorted(Tt[:, 0], ids, side="right")
    T = int(last[0] - first[0])
    if T < 1 or np.any(last - first != T):
        raise L.EmbError(L.EMB_E_ARG, "sample2track: every track needs the same number of transition rows")
    rows = (first[:, None] + np.arange(T)[None, :]).ravel()
    upd = Tt[rows, 2:].reshape(n, T, len(dyn_t))                          # (n, T, n_dyn)
    tv = list(model.timevarying_vars)
    dense = np.zeros((n, len(tv), T), dtype=np.float32)
    for k, v in enumerate(dyn_t):
        dense[:, tv.index(v), :] = upd[:, :, k]
    res = TrackResult(n=n, T=T, dyn_vars=dyn_t, tv_vars=tv, bins_tiled=None, values_tiled=tile(dense), init_bins=None,
                      init_values=np.ascontiguousarray(Ti[:, 1:].T), attempts=None)
    xyz, good = integrate_tracks(model, res, opts=o)
    is_good = good.astype(bool)
    # ---- output tree and files (:143-178, :246-290)
    b_alt = model.boundaries[o.idx_altitude - 1]
    min_alt, max_alt = float(b_alt[0]), float(b_alt[-1])
    Ls = np.arange(math.floor(min_alt - 50) if min_alt % 100 else min_alt, max_alt + 200 + 1e-9, 100.0)      # :151-158
    if Ls[0] < 0:
        Ls[0] = 0
    iG, iA = _find(names_init, label_initial_geographic), _find(names_init, label_initial_airspace)
    speed0 = Ti[:, o.idx_speed] * o.ur_speed
    for k in np.nonzero(is_good)[0]:
        z0 = float(xyz[2, 0, k])
        parts = []
        if iG:
            parts.append("G%d" % int(Ti[k, iG]))
        if iA:
            parts.append("A%d" % int(Ti[k, iA]))
        j = int(np.searchsorted(Ls, z0, side="right")) - 1                # discretize(z_ft(1), L) (:266)
        parts.append("%dft" % int(Ls[min(max(j, 0), len(Ls) - 2)]))
        d = os.path.join(out_dir_parent, *parts)
        os.makedirs(d, exist_ok=True)
        name = "BAYES_t%d_id%d_alt%d_speed%d.csv" % (T, k + 1, int(math.floor(z0 + 0.5)), int(math.floor(speed0[k] + 0.5)))
        with open(os.path.join(d, name), "w", encoding="utf-8") as f:
            f.write("time_s,x_ft,y_ft,z_ft\n")
            f.write("".join("%d,%.0f,%.0f,%.0f\n" % (t, xyz[0, t, k], xyz[1, t, k], xyz[2, t, k]) for t in range(T + 1)))
    Tc = Ti.copy()
    Tc[:, o.idx_speed] *= o.ur_speed
    Tc[:, o.idx_acceleration] *= o.ur_speed
    Tc[:, o.idx_vertrate] *= o.ur_vertrate
    return is_good, Tc
