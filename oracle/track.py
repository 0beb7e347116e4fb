"""TEST INFRASTRUCTURE (oracle) -- not product code.

Restatement of the per-track loop of /root/reference/code/matlab/sample2track.m:188-244: Euler integration of the
sampled rates (:199-218) and the CFIT / speed rejection (:234-244).  Inputs are the two tables the reference reads back
from em_sample's text files, after the unit conversions of :128-141.  PARITY UNPINNED (see oracle/sampler.py)."""
from __future__ import annotations

import numpy as np

from .terminal import cosd, sind

FT_PER_NM = 1852.0 / 0.3048


def integrate(z0_ft, speed0_kt, dv_kts, dh_ftmin, dpsi_degs, min_speed_kt, max_speed_kt):
    """One track.  dv/dh/dpsi: the T rows of the transition table of this id (temporal_map order is the caller's business).
    Returns time_s, x_ft, y_ft, z_ft (T+1 points), is_good."""
    ur_speed, ur_vertrate, ur_heading = FT_PER_NM / 3600.0, 1.0 / 60.0, 1.0           # :108-125
    z = [float(z0_ft)]
    speed = [float(speed0_kt) * ur_speed]                                             # :129
    heading, x, y, t = [0.0], [0.0], [0.0], [0]
    T = len(dv_kts)
    p = 0
    while t[-1] < T:                                                                  # :199
        uv = float(dh_ftmin[p]) * ur_vertrate                                         # :134, :203
        ua = float(dv_kts[p]) * ur_speed                                              # :135, :204
        ut = float(dpsi_degs[p]) * ur_heading                                         # :136, :205
        t.append(t[p] + 1)
        z.append(z[p] + uv)                                                           # :208-210
        speed.append(speed[p] + ua)
        heading.append(heading[p] + ut)
        x.append(x[p] + speed[p] * cosd(heading[p]))                                  # :212-213
        y.append(y[p] + speed[p] * sind(heading[p]))
        p += 1
    z_a, s_a = np.asarray(z), np.asarray(speed)
    is_cfit = bool(np.any(z_a < 0))                                                   # :235-237
    lo, hi = min_speed_kt * ur_speed, max_speed_kt * ur_speed                         # :140-141
    is_reject_speed = bool(np.any((s_a <= lo) | (s_a >= hi)))                         # :240
    return np.asarray(t), np.asarray(x), np.asarray(y), z_a, (not is_cfit) and (not is_reject_speed)
