"""TEST INFRASTRUCTURE (oracle) -- not product code.

Restatement of the terminal *trajectory* sampler of the reference,
`/root/reference/code/matlab/@CorTerminalModel/createEncounter.m`:

    PropagateTrajectory       :93-265   (one chain: aircraft x {forward, reverse})
    CreateStartDistribution   :268-294  (discretize_bayes.m:14-22 of distance, bearing, heading, altitude, speed)
    CheckTrajectoryConditions :296-329
    createEncounter           :41-90    (fwd/bck concatenation and time sort; the em-core `local_smooth`
                                         of :88-89 is outside the sampled path and is not restated)

The trajectory model files are missing from the public checkout; the oracle runs on any
model of the asserted layout (createEncounter.m:107-116), in practice the synthetic ones of
em_model_manned_bayes_b200/synthetic.py.

MATLAB built-ins are restated as: cosd/sind(x) = cos/sin(x*pi/180) with the exact values at multiples of
90 degrees; atan2d = atan2*180/pi; wrapTo360 (Mapping Toolbox): mod(x,360) with positive multiples of 360
mapped to 360; round(x,2) = round(100x)/100 half away from zero; norm([a;b]) = sqrt(a^2+b^2).

Uniforms ("stream spec v5", terminal part): Philox counter (sample_hi, sample_lo, attempt<<16 | purpose<<8 | chain, index)
with sample = global encounter index, chain = 2*aircraft + (0 forward, 1 reverse), index = step ii (1-based),
attempt = the inner `while is_resample` repetition; purpose TERM_SEL (5): lane d = the rand(2,1) row 2 of the d-th
dynamic variable (dbn_sample.m:133,144); purpose TERM_DD (6): lane d = the dediscretize draw of its event
(createEncounter.m:204,210,219).  PARITY UNPINNED (see oracle/sampler.py).
"""
from __future__ import annotations

import math

import numpy as np

from . import philox as px
from . import sampler as sp

FT_PER_NM = 6076.1154855643

# @CorTerminalModel/getDynamicLimits.m:14-62
DYN_LIMITS = {
    "GENERIC": dict(minVel_ft_s=50.0, maxVel_ft_s=506.0, maxTurnRate_deg_s=12.0, maxAltitude_ft=5000.0, maxVertRate_ft_s=6000 / 60),
    "RTCA228_A1": dict(minVel_ft_s=169.0, maxVel_ft_s=491.0, maxTurnRate_deg_s=1.5, maxAltitude_ft=5000.0, maxVertRate_ft_s=2500 / 60),
    "RTCA228_A2": dict(minVel_ft_s=68.0, maxVel_ft_s=338.0, maxTurnRate_deg_s=3.0, maxAltitude_ft=5000.0, maxVertRate_ft_s=1500 / 60),
    "RTCA228_A3": dict(minVel_ft_s=68.0, maxVel_ft_s=186.0, maxTurnRate_deg_s=7.0, maxAltitude_ft=5000.0, maxVertRate_ft_s=500 / 60),
    "TEST": dict(minVel_ft_s=68.0, maxVel_ft_s=186.0, maxTurnRate_deg_s=7.0, maxAltitude_ft=1200.0, maxVertRate_ft_s=500 / 60),
}


def cosd(x):
    r = math.fmod(x, 360.0)
    if r % 90.0 == 0.0:
        return [1.0, 0.0, -1.0, 0.0][int(round(r / 90.0)) % 4]
    return math.cos(r * (math.pi / 180.0))


def sind(x):
    r = math.fmod(x, 360.0)
    if r % 90.0 == 0.0:
        return [0.0, 1.0, 0.0, -1.0][int(round(r / 90.0)) % 4]
    return math.sin(r * (math.pi / 180.0))


def atan2d(y, x):
    return math.atan2(y, x) * (180.0 / math.pi)


def wrap_to_360(x):
    positive = x > 0
    x = x - 360.0 * math.floor(x / 360.0)      # mod(x, 360)
    if x == 0.0 and positive:
        x = 360.0
    return x


def round2(x):
    y = x * 100.0
    return (math.floor(abs(y) + 0.5) * (1.0 if y >= 0 else -1.0)) / 100.0


def sign(x):
    return int(x > 0) - int(x < 0)


class TerminalKeyed:
    """Uniform provider of the trajectory chains (context-keyed, see module docstring)."""

    def __init__(self, seed):
        self.seed = int(seed)
        self.sample = self.chain = self.step = self.attempt = 0
        self.dyn = []

    def bind(self, n_initial, temporal_map, resample_rates):
        self.dyn = [int(v) for v in np.asarray(temporal_map)[:, 1]]
        return self

    def begin(self, sample, attempt=0):
        pass

    def _u(self, purpose, lane):
        k = px.word(self.seed, self.sample, self.attempt, purpose, self.step, lane, sub=self.chain)
        return float(px.u01(int(k)))

    def select_init(self, var):
        raise AssertionError("every initial variable is preset by CreateStartDistribution")

    def trans_column(self, var_t1, t_max):      # dbn_sample.m:133 rand(2,1); row 1 unused
        assert t_max == 2
        return np.array([np.nan, self._u(px.P_TERM_SEL, self.dyn.index(int(var_t1)))])

    def dedisc(self, d):
        return self._u(px.P_TERM_DD, d)


def propagate_trajectory(parms, alpha_i, alpha_t, is_ownship, dt_s, x0_nm, y0_nm, z0_ft, v0_ft_s, heading0_deg, intent,
                         tmax_s, dyn, U, max_states=100000):
    """createEncounter.m:93-265.  Returns dict of lists t_s, x_nm, y_nm, z_ft, heading_deg, v_ft_s."""
    lab = parms.labels_initial
    assert lab[3] == '"heading"' and lab[4] == '"altitude"' and lab[5] == '"speed"'          # :107-109
    i_dist, i_bear, i_head, i_alt, i_spd = (lab.index('"%s"' % s) + 1 for s in ("distance", "bearing", "heading", "altitude", "speed"))
    dd = parms.boundaries
    alt_edges, spd_edges = np.asarray(dd[i_alt - 1]), np.asarray(dd[i_spd - 1])
    k = np.nonzero(alt_edges <= dyn["maxAltitude_ft"])[0]
    valid_alt = set(range(1, int(k[-1]) + 2)) if k.size else set()                          # :120
    s_ = np.nonzero(~(spd_edges >= dyn["minVel_ft_s"]))[0]
    e_ = np.nonzero(spd_edges <= dyn["maxVel_ft_s"])[0]
    valid_v = set(range(int(s_[-1]) + 1, int(e_[-1]) + 2)) if s_.size and e_.size else set()  # :123-125
    bounds_dist = parms.bounds_initial[i_dist - 1]

    traj = dict(t_s=[], x_nm=[], y_nm=[], z_ft=[], heading_deg=[], v_ft_s=[])
    t_s = 0.0
    xy = [x0_nm, y0_nm]
    c, s = cosd(heading0_deg), sind(heading0_deg)
    v = [c * v0_ft_s - s * 0.0, s * v0_ft_s + c * 0.0]                                      # :148
    z_ft, heading_deg = z0_ft, heading0_deg
    ii = 1
    go = True
    while go:
        U.step = ii
        traj["t_s"].append(t_s); traj["x_nm"].append(xy[0]); traj["y_nm"].append(xy[1])      # :163-168
        traj["z_ft"].append(z_ft); traj["heading_deg"].append(heading_deg)
        traj["v_ft_s"].append(math.sqrt(v[0] * v[0] + v[1] * v[1]))
        xy = [xy[0] + (v[0] * dt_s) / FT_PER_NM, xy[1] + (v[1] * dt_s) / FT_PER_NM]          # :171-173
        curr_hdg = wrap_to_360(atan2d(v[1], v[0]))                                           # :176-177
        traj["heading_deg"][ii - 1] = curr_hdg
        if ii > 1:                                                                           # :180-184
            prev = traj["z_ft"][ii - 2]
            diff = z_ft - prev
            traj["z_ft"][ii - 1] = prev + sign(diff) * min(dyn["maxVertRate_ft_s"], abs(diff))
        # CreateStartDistribution :268-294
        dist = math.sqrt(xy[0] * xy[0] + xy[1] * xy[1])
        start = [intent,
                 sp.discretize_bayes(dist, parms.cutpoints_initial[i_dist - 1]),
                 sp.discretize_bayes(wrap_to_360(atan2d(xy[1], xy[0])), parms.cutpoints_initial[i_bear - 1]),
                 sp.discretize_bayes(heading_deg, parms.cutpoints_initial[i_head - 1]),
                 sp.discretize_bayes(z_ft, parms.cutpoints_initial[i_alt - 1]),
                 sp.discretize_bayes(math.sqrt(v[0] * v[0] + v[1] * v[1]), parms.cutpoints_initial[i_spd - 1])]
        heading_discrete = start[3]
        attempt = 0
        resample = True
        while resample:                                                                      # :192-237
            U.attempt = attempt
            _, events, _ = sp.dbn_sample(parms, alpha_i, alpha_t, 2, start, U)
            resample = False
            for (_, var, b) in events:
                var, b = int(var), int(b)
                if var == 4:
                    if b != heading_discrete:
                        e = dd[i_head - 1]
                        heading_deg = float(e[b - 1]) + (float(e[b]) - float(e[b - 1])) * U.dedisc(0)
                    resample = False
                elif var == 5:
                    if b in valid_alt:
                        z_ft = float(alt_edges[b - 1]) + (float(alt_edges[b]) - float(alt_edges[b - 1])) * U.dedisc(1)
                        resample = False
                    else:
                        resample = True
                elif var == 6:
                    if b in valid_v:
                        v1 = float(spd_edges[b - 1]) + (float(spd_edges[b]) - float(spd_edges[b - 1])) * U.dedisc(2)
                        v1 = max(v1, dyn["minVel_ft_s"])
                        v1 = min(v1, dyn["maxVel_ft_s"])
                        c, s = cosd(heading_deg), sind(heading_deg)
                        v = [c * v1 - s * 0.0, s * v1 + c * 0.0]
                        resample = False
                    else:
                        resample = True
                if resample:
                    break
            attempt += 1
            if attempt > 65535:
                raise sp.OracleError("trajectory resample loop did not terminate")
        turn1 = round2(heading_deg - curr_hdg)                                               # :240-245
        delta = min(abs(turn1), dyn["maxTurnRate_deg_s"]) * sign(turn1)
        c, s = cosd(delta), sind(delta)
        v = [c * v[0] - s * v[1], s * v[0] + c * v[1]]                                       # :251
        t_s = t_s + dt_s
        ii += 1
        d_nm = math.sqrt(xy[0] * xy[0] + xy[1] * xy[1])                                      # :296-329
        violate = (abs(t_s) > tmax_s) or (d_nm > bounds_dist[1]) or (intent in (1, 2) and d_nm <= 0.25) or \
                  (is_ownship and xy[1] > 0.25)
        go = not violate
        if ii > max_states:
            raise sp.OracleError("trajectory did not terminate")
    return traj


def create_encounter_chains(models, geo, seed, sample, tmax_s=120, dyn=("GENERIC", "GENERIC")):
    """createEncounter.m:41-85 for one encounter.  `models[(aircraft, direction)]` -> Parms (aircraft 0/1,
    direction +1/-1); `geo`: dict with own_/int_ intent, distance, bearing, alt, heading, speed.
    Returns chains[(aircraft, direction)] -> trajectory dict, and the merged time-sorted trajectories."""
    chains = {}
    merged = []
    for ac, pre in ((0, "own_"), (1, "int_")):
        x0 = geo[pre + "distance"] * cosd(geo[pre + "bearing"])                              # :46-47
        y0 = geo[pre + "distance"] * sind(geo[pre + "bearing"])
        lim = DYN_LIMITS[dyn[ac].upper()]
        for di, dt in ((0, +1), (1, -1)):
            parms = models[(ac, dt)]
            U = TerminalKeyed(seed).bind(parms.n_initial, parms.temporal_map, parms.resample_rates)
            U.sample, U.chain = int(sample), 2 * ac + di
            a_i = sp.bn_dirichlet_prior(parms.N_initial, 0)                                   # :128-129
            a_t = sp.set_transition_priors(parms.G_transition, parms.r_transition, parms.temporal_map, 1)
            chains[(ac, dt)] = propagate_trajectory(parms, a_i, a_t, ac == 0, dt, x0, y0, geo[pre + "alt"], geo[pre + "speed"],
                                                    geo[pre + "heading"], int(geo[pre + "intent"]), tmax_s, lim, U)
        f, b = chains[(ac, +1)], chains[(ac, -1)]
        tr = {k: list(f[k]) + list(b[k][1:]) for k in f}                                      # :74-78
        order = np.argsort(np.asarray(tr["t_s"]), kind="stable")                             # :81-84
        merged.append({k: [tr[k][i] for i in order] for k in tr})
    return chains, merged


GEO_FIELDS = ("own_intent", "own_distance", "own_bearing", "own_alt", "own_heading", "own_speed",
              "int_intent", "int_distance", "int_bearing", "int_alt", "int_heading", "int_speed")


def create_encounters(model_set, geo, seed, first_sample=0, tmax_s=120, dyn=("GENERIC", "GENERIC")):
    """Batch driver used by the tests: `model_set[(group, intent-1)]` -> Parms with group in own_fwd / own_bck /
    int_fwd / int_bck (CorTerminalModel.m:12-30); `geo` (12, n) rows in GEO_FIELDS order.  Returns what
    emb_terminal_propagate writes: traj (5, 2, 2*tmax+1, n) float64 with NaN where an aircraft has no state
    (slot k <-> t_s = k - tmax) and len (4, n)."""
    geo = np.asarray(geo, dtype=np.float64)
    n = geo.shape[1]
    tmax = int(math.floor(tmax_s))
    traj = np.full((5, 2, 2 * tmax + 1, n), np.nan)
    length = np.zeros((4, n), dtype=np.int16)
    for s in range(n):
        g = {k: float(v) for k, v in zip(GEO_FIELDS, geo[:, s])}
        oi, ii = int(g["own_intent"]), int(g["int_intent"])
        if oi not in (1, 2) or ii not in (1, 2, 3):                                            # createEncounter.m:14-38
            raise sp.OracleError("Unknown int_intent")
        models = {(0, +1): model_set[("own_fwd", oi - 1)], (0, -1): model_set[("own_bck", oi - 1)],
                  (1, +1): model_set[("int_fwd", ii - 1)], (1, -1): model_set[("int_bck", ii - 1)]}
        chains, merged = create_encounter_chains(models, g, seed, first_sample + s, tmax_s, dyn)
        for ac in range(2):
            length[2 * ac, s] = len(chains[(ac, +1)]["t_s"])
            length[2 * ac + 1, s] = len(chains[(ac, -1)]["t_s"])
            m = merged[ac]
            slots = np.asarray(m["t_s"], dtype=np.int64) + tmax
            for f, name in enumerate(("x_nm", "y_nm", "z_ft", "heading_deg", "v_ft_s")):
                traj[f, ac, slots, s] = m[name]
    return traj, length


def get_generated_miss_distance(traj):
    """@CorTerminalModel/CorTerminalModel.m:117-133 for traj = [own, intruder] dicts of arrays (t_s, x_nm, y_nm, z_ft)."""
    t1, t2 = np.asarray(traj[0]["t_s"]), np.asarray(traj[1]["t_s"])
    _, ia, ib = np.intersect1d(t1, t2, return_indices=True)                                     # :119
    dx = np.asarray(traj[0]["x_nm"], dtype=np.float64)[ia] - np.asarray(traj[1]["x_nm"], dtype=np.float64)[ib]
    dy = np.asarray(traj[0]["y_nm"], dtype=np.float64)[ia] - np.asarray(traj[1]["y_nm"], dtype=np.float64)[ib]
    dxy_ft = np.sqrt(dx * dx + dy * dy) * FT_PER_NM                                             # :122
    dz_ft = np.asarray(traj[1]["z_ft"], dtype=np.float64)[ib] - np.asarray(traj[0]["z_ft"], dtype=np.float64)[ia]
    idx = int(np.argmin(dxy_ft))                                                                # :125 first minimum
    return float(dxy_ft[idx]), float(dz_ft[idx]), float(t1[ia[idx]]), int(ia[idx]) + 1, int(ib[idx]) + 1, int(ia.size)


def check_runway_proximity(tr, thres_dist_ft, thres_altlow_ft):
    """CorTerminalModel.m:187-210 (1.68781 is the reference's own constant)."""
    d_ft = np.hypot(np.asarray(tr["x_nm"], dtype=np.float64), np.asarray(tr["y_nm"], dtype=np.float64)) * 1.68781
    close = d_ft <= thres_dist_ft
    low = np.asarray(tr["z_ft"], dtype=np.float64)[close] <= thres_altlow_ft if close.any() else np.zeros(0, dtype=bool)
    return bool(close.any()), bool(np.any(low))
