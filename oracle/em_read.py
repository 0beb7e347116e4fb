"""TEST INFRASTRUCTURE (oracle) -- not product code.

CPU restatement of the reference's model reader, `/root/reference/code/matlab/em_read.m`
and `bn_sort.m`.  Variables/bins stay 1-based wherever they are *values* (bin indices,
`temporal_map`, `order_*`, `zero_bins`) exactly as in MATLAB; Python containers holding them are
ordinary 0-based lists (entry i-1 describes variable i).

PARITY UNPINNED: the reference ships no tests / golden files for this path and
MATLAB is not available here; this restatement is pinned only by the hand-derived known answers of
SURVEY.md A.8 (tests/test_oracle_kat.py).
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


class EmReadError(Exception):
    pass


def bn_sort(G: np.ndarray) -> List[int]:
    """bn_sort.m:14-24 -> `toposort(digraph(G), 'Order', 'stable')`.

    MATLAB documents 'stable' as "ordered by node index where possible": the lexicographically
    smallest topological order, i.e. Kahn's algorithm always removing the smallest ready node.
    Returns 1-based variable ids.  (Tie-breaking only matters for the sequential MT19937
    provider; keyed-Philox results do not depend on it.)"""
    G = np.asarray(G, dtype=bool)
    n = G.shape[0]
    indeg = G.sum(axis=0).astype(int)  # G(parent, child)
    ready = [i for i in range(n) if indeg[i] == 0]
    heapq.heapify(ready)
    order = []
    while ready:
        i = heapq.heappop(ready)
        order.append(i + 1)
        for c in np.nonzero(G[i])[0]:
            indeg[c] -= 1
            if indeg[c] == 0:
                heapq.heappush(ready, int(c))
    if len(order) != n:
        raise EmReadError("Network could not be hierarchically sorted")  # bn_sort.m:23
    return order


@dataclass
class Parms:
    """The struct em_read returns (em_read.m:9-24) plus the dependent properties the
    EncounterModel class derives (EncounterModel.m:290-339)."""
    labels_initial: List[str] = field(default_factory=list)
    n_initial: int = 0
    G_initial: Optional[np.ndarray] = None
    order_initial: List[int] = field(default_factory=list)
    r_initial: Optional[np.ndarray] = None
    N_initial: List[np.ndarray] = field(default_factory=list)
    labels_transition: List[str] = field(default_factory=list)
    n_transition: int = 0
    G_transition: Optional[np.ndarray] = None
    order_transition: List[int] = field(default_factory=list)
    r_transition: Optional[np.ndarray] = None
    N_transition: List[Optional[np.ndarray]] = field(default_factory=list)
    boundaries: List[np.ndarray] = field(default_factory=list)
    resample_rates: Optional[np.ndarray] = None
    temporal_map: Optional[np.ndarray] = None
    zero_bins: List[List[int]] = field(default_factory=list)
    bounds_initial: Optional[np.ndarray] = None
    cutpoints_initial: List[np.ndarray] = field(default_factory=list)
    start: List[Optional[int]] = field(default_factory=list)

    @property
    def has_transition(self) -> bool:
        return self.n_transition > 0


def _numbers(line: str) -> np.ndarray:
    # textscan(line, '%f', 'Delimiter', ' ') : em_read.m:76
    s = line.replace(",", " ").split()
    out = []
    for tok in s:
        try:
            out.append(float(tok))
        except ValueError:
            break  # textscan stops at the first non-numeric token ('*' -> empty)
    return np.asarray(out, dtype=np.float64)


def _getdims(G, r, vars_1based):
    """em_read.m:200-206 : dims(ii,:) = [r(ii) prod(r(G(:,ii)))]"""
    n = G.shape[0]
    dims = np.zeros((n, 2), dtype=np.int64)
    for ii in vars_1based:
        parents = G[:, ii - 1]
        q = int(np.prod(r[parents])) if parents.any() else 1
        dims[ii - 1] = (int(r[ii - 1]), q)
    return dims


def _array2cells(x, dims):
    """em_read.m:191-198 : consecutive blocks, each reshaped column-major to r_i x q_i."""
    cells = []
    index = 0
    for (ri, qi) in dims:
        cnt = int(ri) * int(qi)
        if cnt == 0:
            cells.append(None)
            continue
        if index + cnt > x.size:
            raise EmReadError("count table shorter than sum(r_i * q_i)")
        cells.append(np.array(x[index:index + cnt], dtype=np.float64).reshape((int(ri), int(qi)), order="F"))
        index += cnt
    return cells, index


def extract_zero_bins(boundaries):
    """em_read.m:143-156"""
    out = []
    for b in boundaries:
        z: List[int] = []
        if b.size > 2:
            for j in range(2, b.size + 1):  # j = 2:numel(b)
                if b[j - 2] < 0 and b[j - 1] > 0:
                    z = [j - 1]
        out.append(z)
    return out


def extract_temporal_map(labels_transition):
    """em_read.m:158-177 : rows [idx of X(t), idx of label containing X(t+1) or X(t-1)]"""
    rows = []
    for ii, lab in enumerate(labels_transition, start=1):
        t = lab.find("(t)")
        if t >= 0:
            stem = lab[: t + 1]  # labels{ii}(1:t) : up to and including '('
            fut = [k for k, l2 in enumerate(labels_transition, start=1) if (stem + "t+1)") in l2]
            past = [k for k, l2 in enumerate(labels_transition, start=1) if (stem + "t-1)") in l2]
            for k in fut:
                rows.append((ii, k))
            for k in past:
                rows.append((ii, k))
    return np.asarray(rows, dtype=np.int64).reshape(-1, 2)


def em_read(parameters_filename: str, isOverwriteZeroBoundaries: bool = False,
            idxZeroBoundaries=(1, 2, 3)) -> Parms:
    """em_read.m:1-141"""
    with open(parameters_filename, "r", newline="") as f:
        raw = f.read()
    # textscan(..., 'EndOfLine','\r\n','Whitespace','\r\n') : split on CR/LF, drop empty lines
    lines = [ln for ln in raw.replace("\r", "\n").split("\n") if ln.strip() != ""]
    idx_field = [i for i, ln in enumerate(lines) if "#" in ln]
    p = Parms()
    seen = []
    for i in idx_field:
        name = lines[i].strip()
        row = i + 1
        seen.append(name)
        if name == "# labels_initial":
            p.labels_initial = [s.strip() for s in lines[row].split(",")]
            p.n_initial = len(p.labels_initial)
        elif name == "# G_initial":
            p.G_initial = np.stack([_numbers(lines[row + k]) for k in range(p.n_initial)]).astype(bool)
            p.order_initial = bn_sort(p.G_initial)
        elif name == "# r_initial":
            p.r_initial = _numbers(lines[row]).astype(np.int64)
        elif name == "# N_initial":
            dims = _getdims(p.G_initial, p.r_initial, range(1, p.n_initial + 1))
            p.N_initial, _ = _array2cells(_numbers(lines[row]), dims)
        elif name == "# labels_transition":
            p.labels_transition = [s.strip() for s in lines[row].split(",")]
            p.n_transition = len(p.labels_transition)
        elif name == "# G_transition":
            p.G_transition = np.stack([_numbers(lines[row + k]) for k in range(p.n_transition)]).astype(bool)
            p.order_transition = bn_sort(p.G_transition)
        elif name == "# r_transition":
            p.r_transition = _numbers(lines[row]).astype(np.int64)
        elif name == "# N_transition":
            dims = _getdims(p.G_transition, p.r_transition, range(p.n_initial + 1, p.n_transition + 1))
            p.N_transition, _ = _array2cells(_numbers(lines[row]), dims)
        elif name == "# boundaries":
            p.boundaries = [_numbers(lines[row + jj]) for jj in range(p.n_initial)]
        elif name == "# resample_rates":
            p.resample_rates = _numbers(lines[row])
        else:
            raise EmReadError("Unknown field: %s" % name)  # em_read.m:105

    if "# labels_transition" in seen:
        p.temporal_map = extract_temporal_map(p.labels_transition)
    if "# boundaries" in seen:
        p.zero_bins = extract_zero_bins(p.boundaries)
        if isOverwriteZeroBoundaries:  # em_read.m:119-121 (after zero-bin extraction)
            for k in idxZeroBoundaries:
                p.boundaries[k - 1] = np.zeros(0)
        p.bounds_initial = np.zeros((p.n_initial, 2))
        p.cutpoints_initial = []
        for ii in range(p.n_initial):
            if p.boundaries[ii].size == 0:
                n = p.N_initial[ii].shape[0]
                p.cutpoints_initial.append(np.arange(2, n + 1, dtype=np.float64))
            else:
                p.bounds_initial[ii] = (p.boundaries[ii].min(), p.boundaries[ii].max())
                p.cutpoints_initial.append(p.boundaries[ii][1:-1].copy())
    p.start = [None] * p.n_initial  # EncounterModel.m:259-261 preallocStart
    return p
