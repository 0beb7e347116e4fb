"""Writer for the reference's ASCII model grammar (the inverse of em_read.m:47-107).

The reference has no writer; this one exists so that synthetic models (e.g. the terminal trajectory
DBNs that are missing from the public checkout) and the packed test fixtures can be
materialised as ordinary `model/*.txt` files and then go through the normal reader."""
from __future__ import annotations

import numpy as np


def _num(v) -> str:
    v = float(v)
    if v == int(v) and abs(v) < 1e15:
        return str(int(v))
    return repr(v)


def em_write(path, *, labels_initial, G_initial, r_initial, N_initial, labels_transition=None, G_transition=None,
             r_transition=None, N_transition=None, boundaries=None, resample_rates=None):
    """N_initial / N_transition: lists of r_i x q_i arrays (None for non-dynamic transition variables)."""
    out = []

    def field(name, lines):
        out.append("# " + name)
        out.extend(lines)

    def matrix(G):
        return [" ".join(str(int(v)) for v in row) + " " for row in np.asarray(G, dtype=int)]

    def counts(cells):
        flat = np.concatenate([np.asarray(c, dtype=np.float64).ravel(order="F") for c in cells if c is not None])
        return [" ".join(_num(v) for v in flat) + " "]

    field("labels_initial", [", ".join(labels_initial) + " "])
    field("G_initial", matrix(G_initial))
    field("r_initial", [" ".join(str(int(v)) for v in r_initial) + " "])
    field("N_initial", counts(N_initial))
    if labels_transition:
        field("labels_transition", [", ".join(labels_transition) + " "])
        field("G_transition", matrix(G_transition))
        field("r_transition", [" ".join(str(int(v)) for v in r_transition) + " "])
        field("N_transition", counts(N_transition))
    if boundaries is not None:
        field("boundaries", ["* " if len(b) == 0 else " ".join(_num(v) for v in b) + " " for b in boundaries])
    if resample_rates is not None:
        field("resample_rates", [" ".join(_num(v) for v in resample_rates) + " "])
    with open(path, "w", newline="\n") as f:
        f.write("\n".join(out) + "\n")
    return path
