"""Synthetic terminal *trajectory* DBNs.

The 20 trajectory model files of `model/correlated_terminal/` are missing from the public checkout
(`.MISSING_LARGE_BLOBS`), so config 5 of BASELINE.json can only be exercised with models
of the documented structure (`createEncounter.m:107-116` asserts the variable names and positions;
`doc/model_terminal_traj_fwd.png` / `_bck.png` give the edges):

    initial    : "intent" "distance" "bearing" "heading" "altitude" "speed"      (no edges: every variable
                 is preset by CreateStartDistribution, createEncounter.m:187,268-294)
    transition : heading'  <- distance, bearing, heading
                 altitude' <- distance, bearing, heading, altitude
                 speed'    <- distance, bearing, heading, speed
    no dynamic -> dynamic edge  =>  the frozen-parent branch of dbn_sample.m:95-166; the variable's own value
    at t is its LAST parent, which is what setTransitionPriors.m:20-27 relies on.

Counts are pseudo-random with a strong stay-in-bin diagonal (deterministic in `seed`); bin edges follow the
encounter-geometry model (`terminal_v3_radar_encounter_model.txt:29-38`) widened so that the dynamic limits
of `@CorTerminalModel/getDynamicLimits.m:14-62` both accept and reject bins.  No speed edge coincides with a
minVel/maxVel of that table: createEncounter.m:221-226 clamps a sampled speed to the limit, the clamped speed is rotated
and re-discretised through norm() in the next state (:290), and a limit sitting exactly on a bin edge would make that bin
depend on the last bit of cosd/sind."""
from __future__ import annotations

import numpy as np

from .em_write import em_write

DIST_EDGES = [0, 0.5, 1, 2, 3, 4, 5, 8]
ANGLE_EDGES = list(range(0, 361, 10))
ALT_EDGES = [0, 200, 500, 1000, 1500, 2000, 2500, 3000, 5000, 10000]
SPEED_EDGES = [0, 40, 100, 150, 200, 300, 400, 520, 600]


def terminal_trajectory_model_arrays(seed: int = 0, direction: int = +1, speed_edges=None):
    rs = np.random.RandomState(1000 + 2 * int(seed) + (1 if direction > 0 else 0))
    names = ["intent", "distance", "bearing", "heading", "altitude", "speed"]
    speed_edges = SPEED_EDGES if speed_edges is None else list(speed_edges)
    r_init = [3, len(DIST_EDGES) - 1, 36, 36, len(ALT_EDGES) - 1, len(speed_edges) - 1]
    suffix = "(t+1)" if direction > 0 else "(t-1)"
    labels_initial = ['"%s"' % n for n in names]
    labels_transition = ['"%s(t)"' % n for n in names] + ['"%s%s"' % (n, suffix) for n in ("heading", "altitude", "speed")]
    n, nt = 6, 9
    G_i = np.zeros((n, n), dtype=int)
    G_t = np.zeros((nt, nt), dtype=int)
    for child, parents in ((6, (1, 2, 3)), (7, (1, 2, 3, 4)), (8, (1, 2, 3, 5))):
        for p in parents:
            G_t[p, child] = 1
    r_t = r_init + [r_init[3], r_init[4], r_init[5]]
    N_i = [rs.randint(1, 1000, size=(r, 1)).astype(np.float64) for r in r_init]
    N_t = [None] * nt
    for child, self_t in ((6, 3), (7, 4), (8, 5)):
        parents = np.nonzero(G_t[:, child])[0]
        q = int(np.prod([r_t[p] for p in parents]))
        r = r_t[child]
        c = rs.randint(0, 30, size=(r, q)).astype(np.float64)
        c[rs.random_sample((r, q)) < 0.5] = 0.0                     # sparse off-diagonal mass, some all-zero columns
        block = q // r_t[self_t]                                    # the variable's own value is the slowest index
        for k in range(r):
            cols = slice(block * k, block * (k + 1))
            c[k, cols] += rs.randint(200, 2000, size=block)
            for nb in (k - 1, k + 1):                               # drift to the neighbouring bin
                if 0 <= nb < r:
                    c[nb, cols] += rs.randint(0, 120, size=block)
        N_t[child] = c
    boundaries = [[], DIST_EDGES, ANGLE_EDGES, ANGLE_EDGES, ALT_EDGES, speed_edges]
    return dict(labels_initial=labels_initial, G_initial=G_i, r_initial=r_init, N_initial=N_i,
                labels_transition=labels_transition, G_transition=G_t, r_transition=r_t, N_transition=N_t,
                boundaries=[np.asarray(b, dtype=np.float64) for b in boundaries], resample_rates=np.zeros(n))


def write_terminal_trajectory_model(path: str, seed: int = 0, direction: int = +1, speed_edges=None) -> str:
    """Write a synthetic forward (`direction=+1`) or reverse (`-1`) trajectory model in the reference's file format."""
    return em_write(path, **terminal_trajectory_model_arrays(seed, direction, speed_edges))


# CorTerminalModel.m:62 file-name stems of the ten trajectory models
TRAJECTORY_STEMS = ("ownship_landing_model", "ownship_takeoff_model", "ownship_landing_model_reverse",
                    "ownship_takeoff_model_reverse", "intruder_landing_model", "intruder_takeoff_model",
                    "intruder_transit_model", "intruder_landing_model_reverse", "intruder_takeoff_model_reverse",
                    "intruder_transit_model_reverse")


def write_terminal_model_set(directory: str, prefix: str = "terminal_v3_synthetic", seed: int = 0, speed_edges=None) -> dict:
    """Write the ten trajectory models a CorTerminalModel loads (`<prefix>_<stem>.txt`); returns {stem: path}."""
    import os
    os.makedirs(directory, exist_ok=True)
    paths = {}
    for k, stem in enumerate(TRAJECTORY_STEMS):
        path = os.path.join(directory, "%s_%s.txt" % (prefix, stem))
        if not os.path.exists(path):
            tmp = path + ".tmp%d" % os.getpid()
            write_terminal_trajectory_model(tmp, seed=10 * int(seed) + k, direction=-1 if stem.endswith("_reverse") else +1,
                                            speed_edges=speed_edges)
            os.replace(tmp, path)
        paths[stem] = path
    return paths
