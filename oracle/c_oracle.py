"""TEST INFRASTRUCTURE (oracle) -- ctypes front end of oracle/oracle_c.c.

The model arrays are produced by the oracle's own reader (oracle/em_read.py) and the priors by
oracle/sampler.py; nothing from the product (libemb200.so) is used."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import philox as px
from . import sampler as sp

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle_c.so")


class _Model(C.Structure):
    _fields_ = [
        ("n_initial", C.c_int32), ("n_transition", C.c_int32), ("n_dyn", C.c_int32), ("n_gated", C.c_int32),
        ("is_dynvar_depend", C.c_int32),
        ("G_initial", C.c_void_p), ("G_transition", C.c_void_p), ("r", C.c_void_p),
        ("W_initial", C.c_void_p), ("off_initial", C.c_void_p), ("W_transition", C.c_void_p), ("off_transition", C.c_void_p),
        ("order_initial", C.c_void_p), ("order_transition", C.c_void_p), ("temporal_map", C.c_void_p),
        ("boundaries", C.c_void_p), ("boundaries_off", C.c_void_p), ("zero_bins", C.c_void_p), ("rates", C.c_void_p),
        ("gated", C.c_void_p), ("gate_G", C.c_void_p),
        ("start", C.c_void_p),
        ("reject_uncor", C.c_int32), ("idx_v", C.c_int32), ("idx_dh", C.c_int32), ("idx_L", C.c_int32),
        ("is_quantize500", C.c_int32), ("n_layers", C.c_int32),
        ("layers", C.c_void_p),
        ("max_attempts", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "oracle_c.c")
        if not os.path.exists(SO) or os.path.getmtime(src) > os.path.getmtime(SO):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(SO)
        _lib.oc_sample_tracks.restype = C.c_int
        _lib.oc_sample_tracks.argtypes = [C.POINTER(_Model), C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 6
        _lib.oc_sample_initial.restype = C.c_int
        _lib.oc_sample_initial.argtypes = [C.POINTER(_Model), C.c_uint64, C.c_uint64, C.c_int64, C.c_int32] + [C.c_void_p] * 5
        _lib.oc_num_threads.restype = C.c_int
    return _lib


def _label(labels, name):
    return labels.index(name) + 1 if name in labels else 0


class COracle:
    """Binds one model (oracle reader output) + priors + driver options to the C restatement."""

    def __init__(self, parms, prior=0, start=None, uncor=False, isQuantize500=False, layers=None, max_attempts=65535):
        self.p = parms
        n, nt = parms.n_initial, parms.n_transition
        self._keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a.ctypes.data

        m = _Model()
        m.n_initial, m.n_transition = n, nt
        if prior == "stay":
            a_i = sp.bn_dirichlet_prior(parms.N_initial, 0)
            a_t = sp.set_transition_priors(parms.G_transition, parms.r_transition, parms.temporal_map, 1)
        else:
            a_i = sp.bn_dirichlet_prior(parms.N_initial, prior)
            a_t = sp.bn_dirichlet_prior(parms.N_transition, prior) if nt else []
        off, flat, o = [], [], 0
        for N, a in zip(parms.N_initial, a_i):
            off.append(o)
            w = (N + a).ravel(order="F")
            flat.append(w)
            o += w.size
        m.W_initial, m.off_initial = arr(np.concatenate(flat), np.float64), arr(off, np.int64)
        m.G_initial = arr(parms.G_initial, np.uint8)
        m.order_initial = arr(parms.order_initial, np.int32)
        if nt:
            tm = np.asarray(parms.temporal_map)
            m.n_dyn = tm.shape[0]
            dv = tm[:, 1] - 1
            m.is_dynvar_depend = int(parms.G_transition[np.ix_(dv, dv)].any())
            off, flat, o = [], [], 0
            for N, a in zip(parms.N_transition, a_t):
                if N is None:
                    off.append(-1)
                    continue
                if a is None:
                    a = np.zeros_like(N)
                off.append(o)
                w = (N + a).ravel(order="F")
                flat.append(w)
                o += w.size
            m.W_transition, m.off_transition = arr(np.concatenate(flat), np.float64), arr(off, np.int64)
            m.G_transition = arr(parms.G_transition, np.uint8)
            m.r = arr(parms.r_transition, np.int32)
            m.order_transition = arr(parms.order_transition, np.int32)
            m.temporal_map = arr(tm, np.int32)
        else:
            m.r = arr(parms.r_initial, np.int32)
        m.boundaries = arr(np.concatenate(list(parms.boundaries) + [np.zeros(1)]), np.float64)
        m.boundaries_off = arr(np.concatenate([[0], np.cumsum([len(b) for b in parms.boundaries])]), np.int32)
        m.zero_bins = arr([z[0] if z else 0 for z in parms.zero_bins], np.int32)
        rates = np.asarray(parms.resample_rates, dtype=np.float64)
        m.rates = arr(rates, np.float64)
        dynt = {int(v) for v in tm[:, 0]} if nt else set()
        gated = [i + 1 for i in range(n) if rates[i] > 0 or (i + 1) in dynt] if nt else []   # stream spec v5
        m.n_gated = len(gated)
        m.gated = arr(gated + [0], np.int32)
        m.gate_G = arr([px.gate_threshold(rates[g - 1]) for g in gated] + [0], np.uint64)
        st = parms.start if start is None else start
        m.start = arr([0 if s is None else int(s) for s in st], np.int32)
        if uncor:
            lab = parms.labels_initial
            m.reject_uncor = 1
            m.idx_v, m.idx_dh, m.idx_L = _label(lab, '"v"'), _label(lab, '"\\dot h"'), _label(lab, '"L"')
            m.is_quantize500 = int(bool(isQuantize500))
            if layers is not None and len(layers):
                layers = np.asarray(layers, dtype=np.float64).reshape(-1, 2)
                m.n_layers = layers.shape[0]
                m.layers = arr(layers, np.float64)
        m.max_attempts = max_attempts
        self.m = m

    def sample_tracks(self, n, T, seed, first_sample=0, threads=0, want_dense=True):
        ni = self.p.n_initial
        ib = np.zeros((n, ni), dtype=np.int8)
        iv = np.zeros((n, ni), dtype=np.float64)
        att = np.zeros(n, dtype=np.int32)
        sm = np.zeros((n, ni, T), dtype=np.float64) if want_dense else None
        sb = np.zeros((n, ni, T), dtype=np.int8) if want_dense else None
        ne = np.zeros(n, dtype=np.int32)
        rc = lib().oc_sample_tracks(C.byref(self.m), seed, first_sample, n, T, threads, ib.ctypes.data, iv.ctypes.data,
                                    att.ctypes.data, sm.ctypes.data if want_dense else None,
                                    sb.ctypes.data if want_dense else None, ne.ctypes.data)
        return dict(rc=rc, init_bins=ib, init_values=iv, attempts=att, samples=sm, sample_bins=sb, n_events=ne)

    def sample_initial(self, n, seed, first_sample=0, threads=0, box=None):
        ni = self.p.n_initial
        bins = np.zeros((n, ni), dtype=np.int8)
        vals = np.zeros((n, ni), dtype=np.float64)
        att = np.zeros(n, dtype=np.int32)
        lo = hi = None
        if box is not None:
            lo = np.ascontiguousarray(box[0], dtype=np.float64)
            hi = np.ascontiguousarray(box[1], dtype=np.float64)
        rc = lib().oc_sample_initial(C.byref(self.m), seed, first_sample, n, threads, lo.ctypes.data if box is not None else None,
                                     hi.ctypes.data if box is not None else None, bins.ctypes.data, vals.ctypes.data,
                                     att.ctypes.data)
        return dict(rc=rc, bins=bins, values=vals, attempts=att)

    def dense_compact(self, res):
        """(bins (n, n_dyn, T), values (n, n_tv, T)) in the layout of emb_sample_tracks."""
        tm = np.asarray(self.p.temporal_map)
        dyn = [int(v) - 1 for v in tm[:, 0]]
        rates = np.asarray(self.p.resample_rates)
        tv = sorted(set(dyn) | {i for i in range(self.p.n_initial) if rates[i] > 0})
        return res["sample_bins"][:, dyn, :], res["samples"][:, tv, :]
