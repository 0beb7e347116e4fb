"""TEST INFRASTRUCTURE (oracle) -- not product code.

Philox4x32-10 counter-based generator (Salmon et al., "Parallel random numbers: as easy
as 1, 2, 3", SC'11; the Random123 `philox4x32_R(10, ctr, key)` function) and the keyed
stream layout ("stream spec v5") shared by the oracle and the CUDA sampler.

The reference (`/root/reference/code/matlab/select_random.m:14`, `dbn_sample.m:133`,
`resample_events.m:24`, `dediscretize.m:39`) draws from MATLAB's global `rand`.  The B200
sampler replaces *when* a uniform is consumed by *what it is for*: every uniform is a pure
function of (seed, sample, attempt, purpose, index, lane).  The oracle is fed exactly these
uniforms (uniform-injection), so bin indices must be bit-identical.

Stream spec v5
--------------
key     = (seed & 0xffffffff, seed >> 32)
counter = (sample >> 32, sample & 0xffffffff, (attempt << 16) | (purpose << 8) | sub, index)
A call returns four 32-bit words; `lane` picks one.  (With the sample in counter word 1 and the index in word 3,
and words 0 and 2 the same for every track of a launch, the first three Philox rounds factor into a part that only
depends on the track and a part that only depends on the index -- the CUDA kernel computes each once and runs 7 of
the 10 rounds per call, csrc/emb_device.cuh: philox_track / philox_call / philox_finish.  The words are the plain
Philox4x32-10 words of that counter.)

purpose INIT (1):    ONE word per initial variable, and four consecutive samples share a call:
                         k_i(s) = philox(counter with sample := s >> 2, index = i)[lane = s & 3],   i = 0 .. n_initial-1
                     (a thread that samples four consecutive samples consumes whole calls: n_initial calls per four
                     samples).  The word selects the variable, u_sel = (k_i + 0.5) 2**-32 (bn_sample.m:55), and the
                     de-discretisation of the initial value (dbn_hierarchical_sample.m:29) takes
                         u_dd = ((((k_i * B + k_j) mod 2**32) >> 9) + 0.5) 2**-23,   j = (i + 1) mod max(n_initial, 2)
                     (B as below; k_j is an independent full-entropy word, so u_dd is exactly uniform and independent of
                     the select of variable i; for a single-variable network j = 1 is one extra word).
purpose STEP (2):    ONE word per (second, variable): k(e, g), e = 1..T the second, g the ordinal of the variable
                     among the *gated* variables -- in increasing id, the initial variables that have a resample
                     rate > 0 or are dynamic (temporal_map column 1); nw = their number.  One Philox call holds four
                     consecutive seconds of ONE variable:  index = (e >> 2) * nw + g,  lane = e & 3
                     (so a thread that owns one variable of one track consumes whole calls, and a thread that owns
                     the whole track consumes nw calls per four seconds; the lane of e = 0 is unused).
                     attempt = 0 ALWAYS: the driver's rejection test (UncorEncounterModel.m:275) reads only the
                     initial draw, so the seconds of the accepted attempt are the same words whichever attempt
                     was accepted (the seconds of a rejected attempt are discarded by the reference anyway).
purpose LAYER (4):   index = 0, lane = 0 -> altitude-layer draw (UncorEncounterModel.m:260)
purpose TERM_* (5+): terminal trajectory chains, see oracle/terminal.py

word -> uniform: u = (k + 0.5) * 2**-32  (strictly inside (0,1), exact in fp64).

What happens to variable v (gated ordinal g) in second e comes from its word k = k(e, g) and, for the value, the word
of the same variable in the cyclically next second of the same call, k' = k(e', g), e' = (e & ~3) | ((e + 1) & 3), with
A = 0x9E3779B1, B = 0x85EBCA6B, both odd:
  * transition select of a dynamic variable, loop index t = e + 1 (dbn_sample.m:77 / :133,144):
        u_sel  = (k + 0.5) 2**-32
  * resample gate `rand < rate` (resample_events.m:24):
        u_gate = (((k * A) mod 2**32) + 0.5) 2**-32,   i.e. it fires iff (k * A) mod 2**32 < G,
        G = #{h : (h + 0.5) 2**-32 < rate}  (G = 0 for rate 0)
  * every de-discretisation of v that takes effect in second e -- the re-emitted bin of a fired gate
    and/or the new bin of a transition event (dbn_hierarchical_sample.m:35):
        u_dd   = ((((k * B + k') mod 2**32) >> 9) + 0.5) 2**-23
The reference draws these three uniforms independently.  Here u_dd is exactly uniform and independent of v's own
select and gate in that second (k' is an independent full-entropy word), and pairwise independent of every other
decision; what remains coupled is (i) select and gate of the same variable in the same second, through the odd
multiplier A that spreads any interval of k evenly over the gate word (a Kronecker lattice: given a transition of
probability p the gate frequency is off by O(1/(p 2**32))), and (ii) three-way: u_dd of v in second e given *both* v's
decisions in e and in e'.  (spec v3 took u_dd from k alone, so after a rare transition the value could only take
p 2**32 distinct values; spec v4 used the next *variable's* word as partner, which ties the four variables of a track to
one thread; spec v1/v2 spent separate Philox calls: 7 words per second for the 7-variable models instead of 4.)
The 23-bit resolution of u_dd makes the word -> float conversion exact in fp32 arithmetic on the GPU.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

P_INIT = 1
P_STEP = 2
P_LAYER = 4
P_TERM_SEL = 5
P_TERM_DD = 6

TWO_M32 = 2.0 ** -32


def philox4x32_10(ctr, key):
    """ctr: uint32 array (..., 4); key: (k0, k1) python ints.  Returns uint32 (..., 4)."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    c0 = ctr[..., 0].astype(np.uint64)
    c1 = ctr[..., 1].astype(np.uint64)
    c2 = ctr[..., 2].astype(np.uint64)
    c3 = ctr[..., 3].astype(np.uint64)
    k0 = int(key[0]) & 0xFFFFFFFF
    k1 = int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0  # 32x32 -> 64, no overflow in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    out = np.stack([c0, c1, c2, c3], axis=-1)
    return out.astype(np.uint32)


def make_counter(sample, index, attempt, purpose, sub=0):
    """Vectorised counter builder (broadcasts its arguments)."""
    sample = np.asarray(sample, dtype=np.uint64)
    index = np.asarray(index, dtype=np.uint64)
    attempt = np.asarray(attempt, dtype=np.uint64)
    sub = np.asarray(sub, dtype=np.uint64)
    w3 = (attempt << np.uint64(16)) | np.uint64(purpose << 8) | sub
    sample, index, w3 = np.broadcast_arrays(sample, index, w3)
    ctr = np.empty(sample.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = (sample >> np.uint64(32)).astype(np.uint32)
    ctr[..., 1] = (sample & MASK32).astype(np.uint32)
    ctr[..., 2] = (w3 & MASK32).astype(np.uint32)
    ctr[..., 3] = (index & MASK32).astype(np.uint32)
    return ctr


def seed_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return (seed & 0xFFFFFFFF, seed >> 32)


def word(seed, sample, attempt, purpose, index, lane, sub=0):
    """One 32-bit word of the keyed stream (scalar or broadcast)."""
    out = philox4x32_10(make_counter(sample, index, attempt, purpose, sub), seed_key(seed))
    lane = np.asarray(lane)
    if lane.ndim == 0:
        return out[..., int(lane)]
    return np.take_along_axis(out, lane[..., None].astype(np.int64), axis=-1)[..., 0]


def word_at(seed, sample, attempt, purpose, position):
    """Word at linear position p of a (sample, attempt, purpose) stream: index=p//4, lane=p%4."""
    position = np.asarray(position, dtype=np.int64)
    return word(seed, sample, attempt, purpose, position // 4, position % 4)


def u01(k):
    """32-bit word -> uniform strictly inside (0,1), exact in fp64."""
    return (np.asarray(k, dtype=np.float64) + 0.5) * TWO_M32


GATE_MULT = 0x9E3779B1
DD_MULT = 0x85EBCA6B


def gate_word(k):
    """step word -> the 32-bit word the resample gate compares (stream spec v5)."""
    return (int(k) * GATE_MULT) & 0xFFFFFFFF


def init_word(seed, sample, attempt, i):
    """select word k_i of initial variable i (0-based) for `sample` (stream spec v5: four consecutive samples share a call)."""
    return int(word(seed, int(sample) >> 2, attempt, P_INIT, int(i), int(sample) & 3))


def init_partner(i, n_initial):
    """index of the word mixed into the de-discretisation word of initial variable i"""
    return (int(i) + 1) % max(int(n_initial), 2)


def init_dd_uniform(k, k_partner):
    """de-discretisation uniform of an initial variable: the same 23-bit construction as dd_uniform"""
    return dd_uniform(k, k_partner)


def step_position(e, g, nw):
    """(index, lane) of the step word of second e, gated ordinal g (stream spec v5)."""
    return (int(e) >> 2) * int(nw) + int(g), int(e) & 3


def partner_second(e):
    """second whose word (same variable) is mixed into the de-discretisation word of second e (stream spec v5)."""
    return (int(e) & ~3) | ((int(e) + 1) & 3)


def dd_uniform(k, k_next=0):
    """step word k and partner word k_next (same variable, second partner_second(e))
    -> de-discretisation uniform (stream spec v5), exact in fp64."""
    h = (int(k) * DD_MULT + int(k_next)) & 0xFFFFFFFF
    return (float(h >> 9) + 0.5) * 2.0 ** -23


def gate_threshold(rate: float) -> int:
    """G = number of words k in [0, 2^32) with (k + 0.5) * 2^-32 < rate, via the literal fp64 test
    the reference performs (`rand(size(rates)) < rates`, resample_events.m:24)."""
    rate = float(rate)
    if not rate > 0.0:
        return 0
    if rate >= 1.0:
        return 1 << 32
    g = int(np.floor(rate * 4294967296.0))
    g = max(0, min(g, (1 << 32) - 1))

    def fires(k):  # literal comparison
        return (float(k) + 0.5) * TWO_M32 < rate

    while g > 0 and not fires(g - 1):
        g -= 1
    while g < (1 << 32) and fires(g):
        g += 1
    return g
