"""Packed model archives (.npz) <-> reference-format model/*.txt files.

The reference's model FILES are not vendored in this repository.  Tests and bench.py on a GPU box
(where /root/reference does not exist) materialise the models they need from a packed fixture of their
count tables, `tests/golden/models.npz` (written by tests/golden/make_fixtures.py; the data is the
reference's, BSD-2-Clause, notice in tests/golden/MODELS_NOTICE.txt), and then read them through the
normal reader, exactly as a user would read `model/*.txt`.  The archive is test/bench data, not part of
the product: `EMB200_MODEL_ARCHIVE` overrides its location, and a user of the library passes real
`model/*.txt` paths."""
from __future__ import annotations

import os

import numpy as np

from .em_write import em_write

DEFAULT_ARCHIVE = os.environ.get("EMB200_MODEL_ARCHIVE") or os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "models.npz")


def _cells(flat, G, r, first):
    n = G.shape[0]
    out = [None] * n
    o = 0
    for i in range(first, n):
        q = int(np.prod(r[G[:, i].astype(bool)])) if G[:, i].any() else 1
        cnt = int(r[i]) * q
        out[i] = np.asarray(flat[o:o + cnt], dtype=np.float64).reshape((int(r[i]), q), order="F")
        o += cnt
    assert o == len(flat)
    return out


def model_names(archive: str = DEFAULT_ARCHIVE):
    with np.load(archive) as z:
        return sorted({k.split("/")[0] for k in z.files})


def materialize(dest_dir: str, names=None, archive: str = DEFAULT_ARCHIVE):
    """Write `<dest_dir>/<name>.txt` for each requested model; returns {name: path}."""
    os.makedirs(dest_dir, exist_ok=True)
    paths = {}
    with np.load(archive) as z:
        avail = sorted({k.split("/")[0] for k in z.files})
        for name in (names or avail):
            if name not in avail:
                raise KeyError("model %r not in %s" % (name, archive))
            path = os.path.join(dest_dir, name + ".txt")
            paths[name] = path
            if os.path.exists(path):
                continue
            g = lambda k: z[name + "/" + k]  # noqa: E731
            G_i, r_i = g("G_initial"), g("r_initial")
            kw = dict(labels_initial=str(g("labels_initial")).split("\n"), G_initial=G_i, r_initial=r_i,
                      N_initial=_cells(g("N_initial"), G_i, r_i, 0))
            if name + "/labels_transition" in z.files:
                G_t, r_t = g("G_transition"), g("r_transition")
                kw.update(labels_transition=str(g("labels_transition")).split("\n"), G_transition=G_t, r_transition=r_t,
                          N_transition=_cells(g("N_transition"), G_t, r_t, len(r_i)))
            bl = g("boundaries_len")
            flat = g("boundaries")
            b, o = [], 0
            for k in bl:
                b.append(flat[o:o + k])
                o += k
            kw.update(boundaries=b, resample_rates=g("resample_rates"))
            tmp = path + ".tmp%d" % os.getpid()
            em_write(tmp, **kw)
            os.replace(tmp, path)
    return paths
