"""TEST INFRASTRUCTURE (oracle) -- not product code.

Restatement of @UncorEncounterModel/getDynamicLimits.m:1-130 (speed and vertical-rate limits from the 1st / 99th percentiles of
the count tables, conditioned on G, A, L and v when the model has them in positions 1, 2, 3, 4, 6), statement by statement with
MATLAB's own indexing idiom `N(:, d:r:end)` kept as 1-based strided column picks.
"""
from __future__ import annotations

import numpy as np


def discretize_bayes(x, cutpoints):
    """discretize_bayes.m:14-22 (1-based bins)"""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    c = np.asarray(cutpoints, dtype=np.float64)
    d = np.zeros(x.shape, dtype=np.int64)
    for i, v in enumerate(x):
        k = np.nonzero(v < c)[0]
        d[i] = (k[0] + 1) if k.size else c.size + 1
    return d


def _cols(N, d, r):
    """N(:, d:r:end) with 1-based d"""
    return N[:, (d - 1)::r]


def get_dynamic_limits(parms, initial, results=None, idx_G=None, idx_A=None, idx_L=None, idx_V=None, idx_DH=None,
                       is_discretized=None, is_rotorcraft=False):
    prct_low, prct_high = 1, 99                                                                  # :10-11
    N, r = parms.N_initial, [int(v) for v in parms.r_initial]
    empty = lambda v: v is None or v == 0
    is_idx = not any(empty(v) for v in (idx_G, idx_A, idx_L, idx_V, idx_DH))                     # :14
    if is_idx and (idx_G == 1 and idx_A == 2 and idx_L == 3 and idx_V == 4 and idx_DH == 6):     # :17
        if is_discretized[idx_G - 1]:                                                            # :20-24
            dG = int(initial[idx_G - 1])
        else:
            dG = int(discretize_bayes(initial[idx_G - 1], parms.cutpoints_initial[idx_G - 1])[0])
        if is_discretized[idx_A - 1]:                                                            # :27-31
            dA = int(initial[idx_A - 1])
        else:
            dA = int(discretize_bayes(initial[idx_A - 1], parms.cutpoints_initial[idx_A - 1])[0])
        if is_discretized[idx_L - 1]:                                                            # :34-39
            dL = [int(initial[idx_L - 1])]
        else:
            d = discretize_bayes([np.min(results["up_ft"]), np.max(results["up_ft"])], parms.cutpoints_initial[idx_L - 1])
            dL = list(range(int(d.min()), int(d.max()) + 1))
        if is_discretized[idx_V - 1]:                                                            # :45-51
            dV = [int(initial[idx_V - 1])]
        else:
            kts = np.asarray(results["speed_ftps"], dtype=np.float64) * 0.592484
            d = discretize_bayes([kts.min(), kts.max()], parms.cutpoints_initial[idx_V - 1])
            dV = list(range(int(d.min()), int(d.max()) + 1))
        v_G = _cols(N[idx_V - 1], dG, r[idx_G - 1])                                              # :57
        dh_G = _cols(N[idx_DH - 1], dG, r[idx_G - 1])                                            # :58
        v_GA = _cols(v_G, dA, r[idx_A - 1])                                                      # :61
        dh_GA = _cols(dh_G, dA, r[idx_A - 1])                                                    # :62
        v_GAL = v_GA[:, [k - 1 for k in sorted(set(dL))]]                                        # :66
        dh_GAL = np.zeros((r[idx_DH - 1], 1))                                                    # :69-72
        for di in sorted(set(dL)):
            dh_GAL = dh_GAL + _cols(dh_GA, di, r[idx_L - 1])
        dh_GALV = np.zeros((r[idx_DH - 1], 1))                                                   # :75-78
        for di in sorted(set(dV)):
            dh_GALV = dh_GALV + _cols(dh_GAL, di, r[idx_V - 1])
        assert dh_GALV.shape[1] == r[4]                                                          # :79
        v_initial = v_GAL.sum(axis=1)                                                            # :82-83
        dh_initial = dh_GALV.sum(axis=1)
    else:
        v_initial = N[idx_V - 1].sum(axis=1)                                                     # :85-86
        dh_initial = N[idx_DH - 1].sum(axis=1)
    prob_V = 100 * v_initial / v_initial.sum()                                                   # :93-96
    cs_V = np.cumsum(prob_V)
    k_min_V = int(np.nonzero(cs_V >= prct_low)[0][0]) + 1
    k_max_V = int(np.nonzero(cs_V >= prct_high)[0][0]) + 1
    bV = np.asarray(parms.boundaries[idx_V - 1], dtype=np.float64)
    min_speed = bV[k_min_V + 1 - 1] * 1.68780972222222                                           # :99-100
    max_speed = bV[k_max_V + 1 - 1] * 1.68780972222222
    if is_rotorcraft and max_speed > 304:                                                        # :104-109
        max_speed = 304.0
    if (not is_rotorcraft) and min_speed < 30:
        min_speed = 30.0
    prob_DH = 100 * dh_initial / dh_initial.sum()                                                # :115-118
    cs_DH = np.cumsum(prob_DH)
    k_min_DH = int(np.nonzero(cs_DH >= prct_low)[0][0]) + 1
    k_max_DH = int(np.nonzero(cs_DH >= prct_high)[0][0]) + 1
    bH = np.asarray(parms.boundaries[idx_DH - 1], dtype=np.float64)
    vr = float(np.max(np.abs(np.array([bH[k_min_DH], bH[k_max_DH]]) / 60)))                      # :120
    if np.isnan(vr):
        vr = 0.0
    return dict(minVel_ft_s=float(min_speed), maxVel_ft_s=float(max_speed), maxVertRate_ft_s=vr)
