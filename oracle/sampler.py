"""TEST INFRASTRUCTURE (oracle) -- not product code.

Line-by-line CPU restatement of the reference's sampler core (all paths relative to
/root/reference/code/matlab/):

  asub2ind.m, select_random.m, bn_sample.m (+ /bn_sample.m NaN variant), dbn_sample.m (both
  branches, quirks included), dbn_hierarchical_sample.m, resample_events.m, dediscretize.m,
  bn_dirichlet_prior.m, setTransitionPriors.m, events2samples.m, events2controls.m,
  discretize_bayes.m

Bins, variable ids and `t` are 1-based values exactly as in MATLAB.  Every `rand` of the reference
is a call on a uniform provider (oracle/uniforms.py) that receives the context of the draw.

PARITY UNPINNED (no reference tests, no MATLAB here): pinned by SURVEY.md A.8 known answers only.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


class OracleError(Exception):
    pass


# ------------------------------------------------------------------------------------------------
def asub2ind(siz, x) -> int:
    """asub2ind.m:13-14 : k = [1 cumprod(siz(1:end-1))]; ndx = k*(x-1)+1"""
    siz = np.asarray(siz, dtype=np.float64).ravel()
    x = np.asarray(x, dtype=np.float64).ravel()
    k = np.concatenate(([1.0], np.cumprod(siz[:-1])))
    return int(k @ (x - 1.0) + 1.0)


def select_random_u(weights, r: float) -> int:
    """select_random.m:17-20 with the uniform `r` supplied: s=cumsum(w); first s >= s(end)*r."""
    s = np.cumsum(np.asarray(weights, dtype=np.float64))  # sequential fp64, like MATLAB cumsum
    sthres = s[-1] * r
    x = s >= sthres
    idx = np.nonzero(x)[0]
    if idx.size == 0:
        raise OracleError("select_random: empty find (NaN weights?)")
    return int(idx[0]) + 1


def _is_free(v) -> bool:
    # bn_sample.m:45 `~isempty(start{i})`; root /bn_sample.m:45 adds `& ~isnan(start{i})`
    if v is None:
        return True
    if isinstance(v, (list, tuple, np.ndarray)) and len(v) == 0:
        return True
    try:
        return bool(np.isnan(v))
    except TypeError:
        return False


def bn_sample(G, r, N, alpha, num_samples, start, order, U, sample_ids=None) -> np.ndarray:
    """bn_sample.m:25-58.  Returns num_samples x n matrix of 1-based bins (float64 like MATLAB)."""
    n = len(N)
    assert len(start) == n and len(order) == n  # bn_sample.m:32-34
    G = np.asarray(G, dtype=bool)
    S = np.zeros((num_samples, n))
    for sample_index in range(num_samples):
        if sample_ids is not None:
            U.begin(sample_ids[sample_index], 0)
        for i in order:  # 1-based
            parents = G[:, i - 1]   # bn_sample.m:42 G(parent, child)
            j = 1
            if not _is_free(start[i - 1]):
                n_preset_parents = sum(1 for p in np.nonzero(parents)[0] if not _is_free(start[p]))
                if parents.any() and n_preset_parents < int(parents.sum()):
                    raise OracleError("Attempt to preset a dependent variable")  # bn_sample.m:47
                S[sample_index, i - 1] = start[i - 1]
            else:
                if parents.any():
                    j = asub2ind(np.asarray(r)[:n][parents], S[sample_index, :][parents])
                w = N[i - 1][:, j - 1] + alpha[i - 1][:, j - 1]
                S[sample_index, i - 1] = select_random_u(w, U.select_init(i))
    return S


# ------------------------------------------------------------------------------------------------
def bn_dirichlet_prior(N, prior=0):
    """bn_dirichlet_prior.m:17-38"""
    alpha = []
    for Ni in N:
        if Ni is None:
            alpha.append(None)
            continue
        r, q = Ni.shape
        if isinstance(prior, str):
            if prior.lower() == "dbe":
                alpha.append(np.full((r, q), 1.0 / (r * q)))
            else:
                raise OracleError("prior:notdbe")
        else:
            alpha.append(np.full((r, q), float(prior)))
    return alpha


def set_transition_priors(G, r, temporal_map, prior):
    """setTransitionPriors.m:12-33 : alpha(kk, block kk of the LAST parent) = prior."""
    from .em_read import bn_sort
    G = np.asarray(G, dtype=bool)
    n = G.shape[0]
    alpha: List[Optional[np.ndarray]] = [None] * n
    dyn = [int(v) for v in temporal_map[:, 1]]
    for ii in bn_sort(G):
        if ii in dyn:
            parents = G[:, ii - 1]
            if parents.any():
                nq = int(np.prod(np.asarray(r)[parents]))
                jj = int(temporal_map[dyn.index(ii), 0])
                a = np.zeros((int(r[jj - 1]), nq))
                nb = nq // int(r[jj - 1])
                for kk in range(1, int(r[jj - 1]) + 1):
                    a[kk - 1, nb * (kk - 1): nb * kk] = prior
                alpha[ii - 1] = a
    return alpha


# ------------------------------------------------------------------------------------------------
def dbn_sample(parms, dirichlet_initial, dirichlet_transition, t_max, start, U, strict_quirks=False, correct_dbn=False):
    """dbn_sample.m:1-166.  Returns (initial bins 1 x n_initial, events list of [dt, var, bin],
    provenance list of ('trans', second))."""
    G_transition = np.asarray(parms.G_transition, dtype=bool)
    temporal_map = np.asarray(parms.temporal_map)
    r_transition = np.asarray(parms.r_transition)
    n_initial = parms.n_initial
    N_transition = parms.N_transition
    order_transition = list(parms.order_transition)

    # dbn_sample.m:36 (note: r_transition is passed as r)
    initial = bn_sample(parms.G_initial, r_transition, parms.N_initial, dirichlet_initial, 1, start,
                        parms.order_initial, U)[0]

    dynamic_variables = [int(v) for v in temporal_map[:, 1]]          # :39
    x = np.concatenate([initial, np.zeros(len(dynamic_variables))])   # :40
    delta_t = 0
    # :47 [~, ia] = intersect(order_transition, dynamic_variables, 'stable') -> positions in order
    ia = [pos + 1 for pos, v in enumerate(order_transition) if v in dynamic_variables]
    dv0 = np.asarray(dynamic_variables) - 1
    is_dynvar_depend = bool(G_transition[np.ix_(dv0, dv0)].any())     # :55

    events: List[List[float]] = []
    prov: List[tuple] = []

    # correct_dbn (not in the reference): take the per-step branch also for models without a dynamic -> dynamic
    # edge, i.e. re-evaluate the parent configuration every second instead of freezing it at t = 1 (dbn_sample.m:110-135)
    if is_dynvar_depend or correct_dbn:                               # :65-93 "slow" branch
        for t in range(2, t_max + 1):
            delta_t += 1
            x_old = x.copy()
            for i in order_transition:
                if i in dynamic_variables:
                    parents = G_transition[:, i - 1]
                    j = 1
                    if parents.any():
                        j = asub2ind(r_transition[parents], x[parents])
                    w = N_transition[i - 1][:, j - 1] + dirichlet_transition[i - 1][:, j - 1]
                    x[i - 1] = select_random_u(w, U.select_trans(t, i))
            x[temporal_map[:, 0] - 1] = x[temporal_map[:, 1] - 1]    # :82 map back
            if np.any(x[:n_initial] != x_old[:n_initial]):
                for i in range(1, n_initial + 1):
                    if x[i - 1] != x_old[i - 1]:
                        events.append([delta_t, i, x[i - 1]])
                        prov.append(("trans", t - 1))
                        delta_t = 0
    else:                                                             # :95-166 "fast" branch
        if strict_quirks and len(dynamic_variables) != 3:
            # :61 events preallocated with numel(dynamic_variables) columns, rows are 3 wide (:156)
            raise OracleError("dbn_sample fast branch: MATLAB size mismatch unless 3 dynamic variables")
        s = {}
        sthres = {}
        for ii in order_transition:                                   # :110
            if ii in dynamic_variables:
                parents = G_transition[:, ii - 1]
                if parents.any():
                    j = asub2ind(r_transition[parents], x[parents])   # frozen at t = 1 (F6)
                else:
                    j = 1
                w = N_transition[ii - 1][:, j - 1] + dirichlet_transition[ii - 1][:, j - 1]
                s[ii] = np.cumsum(w)
                sthres[ii] = s[ii][-1] * U.trans_column(ii, t_max)    # :133 (row 1 never used)
        for t in range(2, t_max + 1):                                 # :138
            delta_t += 1
            x_old = x.copy()
            for ii in ia:  # :143 positions in order_transition used as variable ids (quirk)
                if ii not in s:
                    raise OracleError("dbn_sample fast branch: order_transition is not identity at "
                                      "the dynamic variables (reference would index an empty cell)")
                hit = np.nonzero(s[ii] >= sthres[ii][t - 1])[0]
                x[ii - 1] = int(hit[0]) + 1
            x[temporal_map[:, 0] - 1] = x[temporal_map[:, 1] - 1]    # :149
            if np.any(x[:n_initial] != x_old[:n_initial]):
                for ii in range(1, n_initial + 1):
                    if x[ii - 1] != x_old[ii - 1]:
                        events.append([delta_t, ii, x[ii - 1]])
                        prov.append(("trans", t - 1))
                        delta_t = 0
    return initial, events, prov


def resample_events(initial, events, prov, rates, U):
    """resample_events.m:11-37 with provenance.  `second` counts the held seconds 1..T."""
    rates = np.asarray(rates, dtype=np.float64)
    newevents: List[List[float]] = []
    newprov: List[tuple] = []
    x = np.array(initial, dtype=np.float64)
    second = 0
    for row, pv in zip(events, prov):
        holdtime = int(row[0])
        if holdtime == 0:
            newevents.append(list(row))
            newprov.append(pv)
        else:
            delta_t = 0
            for _j in range(holdtime):
                second += 1
                changes = np.nonzero(U.gates(second) < rates)[0] + 1   # :24
                delta_t += 1
                if changes.size:
                    for c_i, c in enumerate(changes):
                        newevents.append([delta_t if c_i == 0 else 0, int(c), x[c - 1]])
                        newprov.append(("gate", second))
                    delta_t = 0
            newevents.append([delta_t, row[1], row[2]])
            newprov.append(pv)
        if row[1] > 0:
            x[int(row[1]) - 1] = row[2]
    return newevents, newprov


def dediscretize_u(d, parameters, zero_bins, u_fn):
    """dediscretize.m:1-41 for scalar d; `u_fn()` is called iff the reference calls rand (:39).
    (`wrap` is never passed by any caller and is omitted.)"""
    if parameters is None or len(parameters) == 0:
        return d                                                     # :7-10
    if zero_bins is not None and any(z == d for z in zero_bins):
        return 0.0                                                   # :24-25
    dd = int(d)
    a = float(parameters[dd - 1])
    b = float(parameters[dd])
    return a + (b - a) * u_fn()                                      # :39


def dbn_hierarchical_sample(parms, dirichlet_initial, dirichlet_transition, sample_time,
                            dediscretize_parameters, zero_bins, resample_rates, start, U,
                            strict_quirks=False, correct_dbn=False):
    """dbn_hierarchical_sample.m:9-37.  Returns (initial continuous, events [[dt,var,value]],
    provenance, initial bins)."""
    initial, events, prov = dbn_sample(parms, dirichlet_initial, dirichlet_transition, sample_time,
                                       start, U, strict_quirks, correct_dbn)
    initial_bins = initial.copy()
    total = sum(e[0] for e in events)
    events = events + [[sample_time - total, 0, 0]]                   # :15-19
    prov = prov + [("end", sample_time)]
    events, prov = resample_events(initial, events, prov, resample_rates, U)   # :22
    initial = initial.copy()
    for ii in range(1, len(initial) + 1):                             # :25-31
        if len(dediscretize_parameters[ii - 1]) == parms.N_initial[ii - 1].shape[0] - 2:
            pass
        else:
            initial[ii - 1] = dediscretize_u(initial[ii - 1], dediscretize_parameters[ii - 1],
                                             zero_bins[ii - 1], lambda v=ii: U.dedisc_init(v))
    event_bins = [e[2] for e in events]
    for k in range(len(events) - 1):                                  # :33-37
        var = int(events[k][1])
        kind, second = prov[k]
        events[k][2] = dediscretize_u(events[k][2], dediscretize_parameters[var - 1], zero_bins[var - 1],
                                      lambda kind=kind, second=second, var=var: U.dedisc_event(kind, second, var))
    return initial, events, prov, initial_bins, event_bins


# ------------------------------------------------------------------------------------------------
def events2samples(initial, events) -> np.ndarray:
    """events2samples.m:9-27 -> n x sum(dt) matrix."""
    n = len(initial)
    T = int(sum(e[0] for e in events))
    d = np.zeros((n, T))
    x = np.array(initial, dtype=np.float64)
    t = 0
    for (delta_t, var, val) in events:
        delta_t = int(delta_t)
        if var == 0:
            t = t + 1
            d[:, t - 1: t - 1 + delta_t] = x[:, None]
        else:
            if delta_t > 0:
                d[:, t: t + delta_t] = x[:, None]
                t = t + delta_t
            x[int(var) - 1] = val
    return d


def events2controls(initial, events, temporal_map) -> np.ndarray:
    """events2controls.m:9-31 -> rows [t, x(vars)] per event with dt > 0."""
    vars_ = np.asarray(temporal_map)[:, 0] - 1
    x = np.array(initial, dtype=np.float64)
    rows = []
    t = 0
    for (delta_t, var, val) in events:
        if delta_t > 0:
            rows.append(np.concatenate(([t], x[vars_])))
            t = t + delta_t
        if var > 0:
            x[int(var) - 1] = val
    return np.asarray(rows).reshape(-1, 1 + len(vars_))


def discretize_bayes(x, thresholds) -> int:
    """discretize_bayes.m:14-22 (scalar)."""
    thresholds = np.asarray(thresholds, dtype=np.float64)
    if x >= thresholds[-1]:
        return thresholds.size + 1
    return int(np.nonzero(x < thresholds)[0][0]) + 1
