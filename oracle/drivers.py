"""TEST INFRASTRUCTURE (oracle) -- not product code.

Restatement of the reference's sampling drivers:
  @UncorEncounterModel/UncorEncounterModel.m:192-313  (sample)
  @CorTerminalModel/sample.m:1-82                      (terminal encounter geometry)
  bn_sample.m:39 batch loop                            (initial-network-only sampling, config 2)
  em_sample.m:41-104                                   (legacy file writer)

`sample ids` are 0-based global indices (they key the Philox stream); everything else is 1-based
like MATLAB.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import sampler as sp


@dataclass
class UncorSample:
    initial: np.ndarray            # out_inits(ii,:)  continuous (bin index for '*' variables)
    events: np.ndarray             # out_events{ii}   k x 3 [dt var value]
    samples: np.ndarray            # out_samples{ii}  n_initial x T
    controls: np.ndarray           # EncounterModelEvents.event  [t dh dpsi dv] in EME units
    initial_bins: np.ndarray       # (oracle extra) sampled bins of the accepted attempt
    sample_bins: np.ndarray        # (oracle extra) n_initial x T bins, same expansion as samples
    event_bins: np.ndarray         # (oracle extra) k bins aligned with events rows
    prov: list = field(default_factory=list)
    attempts: int = 1


def _find_label(labels, name):
    for i, l in enumerate(labels):
        if l == name:
            return i + 1
    return None


def round500(num):
    """UncorEncounterModel.m:196"""
    return 500.0 * (np.floor(num / 500.0) + (np.mod(num, 500.0) > 250.0))


def uncor_sample(parms, n_samples, sample_time, U, *, first_sample=0, prior=0, isQuantize500=False,
                 layers=None, start=None, strict_quirks=False, max_attempts=65535, correct_dbn=False) -> List[UncorSample]:
    """UncorEncounterModel.m:192-313 (rng seeding is the provider's business)."""
    labels = parms.labels_initial
    idxL = _find_label(labels, '"L"')
    idxV = _find_label(labels, '"v"')
    idxDV = _find_label(labels, '"\\dot v"')
    idxDH = _find_label(labels, '"\\dot h"')
    idxDPsi = _find_label(labels, '"\\dot \\psi"')
    if idxDV is None or idxDH is None or idxDPsi is None:
        raise sp.OracleError("dynvar:empty")                                    # :231-234
    alpha_i = sp.bn_dirichlet_prior(parms.N_initial, prior)
    alpha_t = sp.bn_dirichlet_prior(parms.N_transition, prior)
    start = parms.start if start is None else start
    U.bind(parms.n_initial, parms.temporal_map, parms.resample_rates)
    out = []
    tm = np.asarray(parms.temporal_map)
    for ii in range(n_samples):
        attempt = 0
        while True:                                                              # :248
            U.begin(first_sample + ii, attempt)
            initial, events, prov, ibins, ebins = sp.dbn_hierarchical_sample(
                parms, alpha_i, alpha_t, sample_time, parms.boundaries, parms.zero_bins,
                parms.resample_rates, start, U, strict_quirks, correct_dbn)
            if layers is not None and len(layers):                               # :259-263
                L = int(initial[idxL - 1])
                h_ft = layers[L - 1][0] + U.layer() * (layers[L - 1][1] - layers[L - 1][0])
            else:
                h_ft = initial[idxL - 1]
            if initial[idxDH - 1] == 0 and isQuantize500:                        # :266-268
                h_ft = round500(h_ft)
            if (layers is not None and len(layers)) or isQuantize500:           # :270-272
                initial[idxL - 1] = h_ft
            if initial[idxV - 1] * 1.68781 > abs(initial[idxDH - 1]) / 60:       # :275
                break
            attempt += 1
            if attempt > max_attempts:
                raise sp.OracleError("rejection loop exceeded max_attempts")
        samples = sp.events2samples(initial, events)                             # :283
        controls = sp.events2controls(initial, events, tm)                       # :286
        dynt = list(tm[:, 0])
        idxEME = [dynt.index(idxDH) + 1, dynt.index(idxDPsi) + 1, dynt.index(idxDV) + 1]   # :291 (+1 -> col)
        controls = controls[:, [0] + idxEME]
        controls[:, 1] = controls[:, 1] / 60                                     # :295
        controls[:, 2] = np.deg2rad(controls[:, 2])                              # :296
        controls[:, 3] = controls[:, 3] * 1.68780972222222                       # :297
        bin_events = [[e[0], e[1], b] for e, b in zip(events, ebins)]
        sample_bins = sp.events2samples(ibins, bin_events)
        out.append(UncorSample(initial=np.array(initial), events=np.asarray(events, dtype=np.float64).reshape(-1, 3),
                               samples=samples, controls=controls, initial_bins=np.array(ibins),
                               sample_bins=sample_bins, event_bins=np.asarray(ebins, dtype=np.float64),
                               prov=prov, attempts=attempt + 1))
    return out


def initial_sample(parms, num_samples, U, *, first_sample=0, prior=0, start=None, dediscretize=True):
    """Config 2: bn_sample.m:39-58 over `num_samples`, then the initial-vector half of
    dbn_hierarchical_sample.m:25-31.  Returns (bins num_samples x n, values num_samples x n)."""
    alpha_i = sp.bn_dirichlet_prior(parms.N_initial, prior)
    start = parms.start if start is None else start
    U.bind(parms.n_initial, parms.temporal_map, parms.resample_rates)
    ids = [first_sample + k for k in range(num_samples)]
    S = sp.bn_sample(parms.G_initial, parms.r_initial, parms.N_initial, alpha_i, num_samples, start,
                     parms.order_initial, U, sample_ids=ids)
    V = S.copy()
    if dediscretize:
        for k in range(num_samples):
            U.begin(ids[k], 0)
            for ii in range(1, parms.n_initial + 1):
                V[k, ii - 1] = sp.dediscretize_u(S[k, ii - 1], parms.boundaries[ii - 1], parms.zero_bins[ii - 1],
                                                 lambda v=ii: U.dedisc_init(v))
    return S, V


def terminal_sample(parms, n_samples, U, *, first_sample=0, prior=0, start=None, bounds_sample=None,
                    dyn_limits1=(50.0, 506.0), dyn_limits2=(50.0, 506.0), max_attempts=65535):
    """@CorTerminalModel/sample.m:29-77.  dyn_limits = (minVel_ft_s, maxVel_ft_s) of the GENERIC
    aircraft type (@CorTerminalModel/getDynamicLimits.m:15-17).  Returns (outInits, bins, attempts)."""
    alpha_i = sp.bn_dirichlet_prior(parms.N_initial, prior)
    start = parms.start if start is None else start
    U.bind(parms.n_initial, None, None)
    i_own = _find_label(parms.labels_initial, '"own_speed"')
    i_int = _find_label(parms.labels_initial, '"int_speed"')
    n = parms.n_initial
    out_inits = np.zeros((n_samples, n))
    out_bins = np.zeros((n_samples, n))
    attempts = np.zeros(n_samples, dtype=np.int64)
    for ii in range(n_samples):
        attempt = 0
        while True:
            U.begin(first_sample + ii, attempt)
            bins = sp.bn_sample(parms.G_initial, parms.r_initial, parms.N_initial, alpha_i, 1, start,
                                parms.order_initial, U)[0]                        # sample.m:34
            initial = bins.copy()
            for kk in range(1, n + 1):                                            # sample.m:37-42
                if len(parms.boundaries[kk - 1]):
                    initial[kk - 1] = sp.dediscretize_u(initial[kk - 1], parms.boundaries[kk - 1],
                                                        parms.zero_bins[kk - 1], lambda v=kk: U.dedisc_init(v))
            good = True
            if bounds_sample is not None and len(bounds_sample):                  # sample.m:45-53
                good = bool(np.all((initial >= bounds_sample[:, 0]) & (initial <= bounds_sample[:, 1])))
            if good:                                                              # sample.m:57-71
                s1 = dyn_limits1[0] <= initial[i_own - 1] <= dyn_limits1[1]
                s2 = dyn_limits2[0] <= initial[i_int - 1] <= dyn_limits2[1]
                good = s1 and s2
            if good:
                break
            attempt += 1
            if attempt > max_attempts:
                raise sp.OracleError("rejection loop exceeded max_attempts")
        out_inits[ii] = initial
        out_bins[ii] = bins
        attempts[ii] = attempt + 1
    return out_inits, out_bins, attempts


def init_start_terminal(parms, n_samples=1000000, airspace_class=(False, True, True, True), own_intent=(True, True),
                        int_intent=(True, True, True)):
    """@CorTerminalModel/InitStartTerminal.m:43-92 -> out_start as a list of rows (cells: [] = free).
    The cell is preallocated with n_samples rows (:59) and filled in blocks of n_enc_per_comb rows (:75-80), so it ends up with
    n_combs * ceil(n_samples / n_combs) rows."""
    assert parms.labels_initial[0] == '"airspace_class"'                                             # :31
    assert parms.labels_initial[1] == '"own_intent"'
    assert parms.labels_initial[2] == '"int_intent"'
    assert len(airspace_class) == parms.r_initial[0] and len(own_intent) == parms.r_initial[1]       # :36-38
    assert len(int_intent) == parms.r_initial[2]
    idx_class = [k + 1 for k, f in enumerate(airspace_class) if f]                                   # :42-44
    idx_ownint = [k + 1 for k, f in enumerate(own_intent) if f]
    idx_intint = [k + 1 for k, f in enumerate(int_intent) if f]
    n_combs = len(idx_class) * len(idx_ownint) * len(idx_intint)                                     # :47
    if n_combs > n_samples:                                                                          # :50-53
        n_samples = n_combs
    n_enc_per_comb = int(np.ceil(n_samples / n_combs))                                               # :56
    out_start = [[[] for _ in range(parms.n_initial)] for _ in range(n_samples)]                     # :59
    s = 1                                                                                            # :63
    for ii in idx_class:                                                                             # :67
        for jj in idx_ownint:
            for kk in idx_intint:
                e = s + n_enc_per_comb - 1                                                           # :75
                while len(out_start) < e:                                                            # MATLAB grows the cell on assignment
                    out_start.append([[] for _ in range(parms.n_initial)])
                for row in range(s, e + 1):                                                          # :78-80
                    out_start[row - 1][0] = ii
                    out_start[row - 1][1] = jj
                    out_start[row - 1][2] = kk
                s = e + 1                                                                            # :88
    return out_start


def dbn_tracks(parms, n_samples, sample_time, U, *, first_sample=0, prior=0, start=None, strict_quirks=False):
    """The sampling loop of em_sample.m:78-85 (dbn_hierarchical_sample + events2samples, no
    rejection) for any model with a transition network, e.g. the correlated model cor_v1.txt.
    `prior` follows bn_dirichlet_prior (number or 'dbe'); pass prior='stay' for the terminal
    trajectory priors of createEncounter.m:128-129."""
    if prior == "stay":
        alpha_i = sp.bn_dirichlet_prior(parms.N_initial, 0)
        alpha_t = sp.set_transition_priors(parms.G_transition, parms.r_transition, parms.temporal_map, 1)
    else:
        alpha_i = sp.bn_dirichlet_prior(parms.N_initial, prior)
        alpha_t = sp.bn_dirichlet_prior(parms.N_transition, prior)
    start = parms.start if start is None else start
    U.bind(parms.n_initial, parms.temporal_map, parms.resample_rates)
    out = []
    for ii in range(n_samples):
        U.begin(first_sample + ii, 0)
        initial, events, prov, ibins, ebins = sp.dbn_hierarchical_sample(
            parms, alpha_i, alpha_t, sample_time, parms.boundaries, parms.zero_bins, parms.resample_rates, start, U,
            strict_quirks)
        samples = sp.events2samples(initial, events)
        bin_events = [[e[0], e[1], b] for e, b in zip(events, ebins)]
        out.append(UncorSample(initial=np.array(initial), events=np.asarray(events, dtype=np.float64).reshape(-1, 3),
                               samples=samples, controls=sp.events2controls(initial, events, parms.temporal_map),
                               initial_bins=np.array(ibins), sample_bins=sp.events2samples(ibins, bin_events),
                               event_bins=np.asarray(ebins, dtype=np.float64), prov=prov, attempts=1))
    return out


def em_sample_text(parms, num_initial_samples, num_transition_samples, U, *, start=None):
    """em_sample.m:59-100: the text of initial.txt and transition.txt (fprintf %g formatting)."""
    out = dbn_tracks(parms, num_initial_samples, num_transition_samples, U, start=start)
    tm = np.asarray(parms.temporal_map)
    ini = ["id " + "".join("%s " % l for l in parms.labels_initial) + "\n"]
    tra = ["initial_id t " + "".join("%s " % parms.labels_transition[int(k) - 1] for k in tm[:, 1]) + "\n"]
    for ii, s in enumerate(out, start=1):
        ini.append("%d " % ii + "".join("%g " % x for x in s.samples[:-1, 0]) + "%g" % s.samples[-1, 0] + "\n")
        for j in range(1, num_transition_samples + 1):
            col = s.samples[[int(v) - 1 for v in tm[:, 0]], j - 1]
            tra.append("%g %g " % (ii, j - 1) + "".join("%g " % x for x in col[:-1]) + "%g" % col[-1] + "\n")
    return "".join(ini), "".join(tra)
