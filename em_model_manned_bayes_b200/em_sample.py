"""em_sample.m:1-104 -- the legacy file-output driver (RUN_1_emsample.m:38-47), on top of the C ABI.

    em_sample(parameters_filename, initial_output_filename=..., transition_output_filename=...,
              num_initial_samples=100, num_transition_samples=60, start=None,
              isOverwriteZeroBoundaries=False, idxZeroBoundaries=(1, 2, 3), rng_seed=42)

Writes the same two text files: `initial.txt` ("id <labels>" header, one row `id v1 ... vn` per sample,
em_sample.m:59-66,88-92) and `transition.txt` ("initial_id t <labels of the (t+1) variables>" header, one row
`id t x_dyn...` per sample and second, :68-74,94-100), numbers formatted with C's %g like fprintf.
The samples come from emb_sample_tracks (dense compact output: column c is `samples(:, c+1)` of
events2samples.m), so nothing is sampled on the CPU.  `parms.prior = 'constant'` (em_sample.m:50) is an error in
the reference at HEAD (EncounterModel.m:194-203 accepts only numbers and 'dbe'); the intended
zero prior is used.  The random stream is the keyed Philox stream with seed `rng_seed`, not MATLAB's twister."""
from __future__ import annotations

import numpy as np

from .model import EncounterModel


def em_sample(parameters_filename: str, initial_output_filename: str, transition_output_filename: str,
              num_initial_samples: int = 100, num_transition_samples: int = 60, start=None,
              isOverwriteZeroBoundaries: bool = False, idxZeroBoundaries=(1, 2, 3), rng_seed: int = 42,
              device=None, chunk: int = 1 << 16) -> None:
    parms = EncounterModel(parameters_filename, idxZeroBoundaries=idxZeroBoundaries,
                           isOverwriteZeroBoundaries=isOverwriteZeroBoundaries)
    if start is not None and len(start):
        parms.start = list(start)                                        # em_sample.m:53-55
    n, T = int(num_initial_samples), int(num_transition_samples)
    dyn_t = [int(v) for v in parms.temporal_map[:, 0]]
    tv = list(parms.timevarying_vars)
    cols = [tv.index(v) for v in dyn_t]
    with open(initial_output_filename, "w", encoding="utf-8") as fi, \
            open(transition_output_filename, "w", encoding="utf-8") as ft:
        fi.write("id " + "".join("%s " % l for l in parms.labels_initial) + "\n")                      # :59-66
        ft.write("initial_id t " + "".join("%s " % parms.labels_transition[int(k) - 1]
                                           for k in parms.temporal_map[:, 1]) + "\n")                  # :68-74
        for first in range(0, n, chunk):
            m = min(chunk, n - first)
            res = parms.sample_tracks(m, T, seed=rng_seed, first_sample=first, device=device, want_bins=False)
            iv = res.init_values
            vals = res.values
            if device is not None:
                iv, vals = iv.cpu().numpy(), vals.cpu().numpy()
            iv = np.asarray(iv).T                                        # (m, n_initial) = samples(:, 1)
            vals = np.asarray(vals, dtype=np.float64)[:, cols, :]        # (m, n_dyn, T)
            for k in range(m):
                ii = first + k + 1
                fi.write("%d " % ii + " ".join("%g" % x for x in iv[k]) + "\n")                        # :88-92
                rows = vals[k].T
                ft.write("".join("%g %g " % (ii, j) + " ".join("%g" % x for x in rows[j]) + "\n" for j in range(T)))  # :94-100
