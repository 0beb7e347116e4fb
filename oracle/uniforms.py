"""TEST INFRASTRUCTURE (oracle) -- not product code.

Uniform providers for the restated reference sampler.  The restatement (oracle/sampler.py) asks the
provider for a uniform at every place the reference calls `rand`, passing the *context* of the
draw.  Two providers:

* KeyedPhilox  -- the product stream (oracle/philox.py, "stream spec v5"): the uniform is a pure
  function of the context.  This is the uniform-injection hook of BASELINE.json's north star: the
  reference algorithm, fed these uniforms, must give bit-identical bins to the CUDA sampler.
* MTStream     -- MATLAB's `rng(seed,'twister'); rand` emulation (MT19937 `genrand_res53`, which
  is what numpy's RandomState.random_sample produces) consumed strictly in the reference's call
  order, ignoring context.  Lets a MATLAB user compare
  `UncorEncounterModel.sample(n, T, 'seed', s)` against the oracle.  UNVERIFIED against MATLAB here.

Every provider also records a tape (list of (context, u)) when `record=True`, which is what the
MATLAB-side shadow `rand` replays (matlab/inject/).
"""
from __future__ import annotations

import numpy as np

from . import philox as px


class _Base:
    record = False

    def __init__(self):
        self.tape = []

    def _rec(self, ctx, u):
        if self.record:
            self.tape.append((ctx, float(u)))
        return u


class KeyedPhilox(_Base):
    """Context-keyed uniforms.  `bind(parms)` must be called once per model so that the dynamic and
    gated ordinals (stream spec v5) are known."""

    def __init__(self, seed: int, record: bool = False):
        super().__init__()
        self.seed = int(seed)
        self.record = record
        self.sample = 0
        self.attempt = 0

    # -- model binding -------------------------------------------------------------------------
    def bind(self, n_initial, temporal_map, resample_rates):
        self.n_initial = int(n_initial)
        tm = np.zeros((0, 2), dtype=np.int64) if temporal_map is None else np.asarray(temporal_map)
        self.dyn_vars_t = [int(v) for v in tm[:, 0]]       # variable ids at time t (1-based)
        self.dyn_vars_t1 = [int(v) for v in tm[:, 1]]      # their (t+1)/(t-1) counterparts
        rates = np.zeros(self.n_initial) if resample_rates is None else np.asarray(resample_rates, dtype=np.float64)
        self.rates = rates
        self.gated = [i + 1 for i in range(rates.size) if rates[i] > 0 or (i + 1) in self.dyn_vars_t]
        self.G = {v: px.gate_threshold(rates[v - 1]) for v in self.gated}
        self.nd = len(self.dyn_vars_t)
        self.nw = len(self.gated)
        return self

    def begin(self, sample: int, attempt: int = 0):
        self.sample = int(sample)
        self.attempt = int(attempt)

    def _w(self, purpose, position):
        assert purpose != px.P_STEP
        return int(px.word_at(self.seed, self.sample, self.attempt, purpose, position))

    # -- draws ---------------------------------------------------------------------------------
    def select_init(self, var):                      # bn_sample.m:55 -> select_random.m:14
        return self._rec(("init_sel", var), px.u01(px.init_word(self.seed, self.sample, self.attempt, var - 1)))

    def dedisc_init(self, var):                      # dbn_hierarchical_sample.m:29 -> dediscretize.m:39
        k = px.init_word(self.seed, self.sample, self.attempt, var - 1)
        kp = px.init_word(self.seed, self.sample, self.attempt, px.init_partner(var - 1, self.n_initial))
        return self._rec(("init_dd", var), px.init_dd_uniform(k, kp))

    def _step(self, e, g):
        """step word of second e, gated ordinal g (spec v5: no attempt in the step stream)"""
        index, lane = px.step_position(e, g, self.nw)
        return int(px.word(self.seed, self.sample, 0, px.P_STEP, index, lane))

    def _sel_word(self, t, var_t1):
        g = self.gated.index(self.dyn_vars_t[self.dyn_vars_t1.index(int(var_t1))])
        return self._step(t - 1, g)

    def select_trans(self, t, var_t1):               # dbn_sample.m:77 (slow branch)
        return self._rec(("trans_sel", t, var_t1), px.u01(self._sel_word(t, var_t1)))

    def trans_column(self, var_t1, t_max):           # dbn_sample.m:133  rand(t_max,1); row 1 unused
        col = np.full(t_max, np.nan)
        self._rec(("trans_sel", 1, var_t1), 0.5)     # row 1 of rand(t_max,1) is drawn but never used (tape padding)
        for t in range(2, t_max + 1):
            col[t - 1] = px.u01(self._sel_word(t, var_t1))
            self._rec(("trans_sel", t, var_t1), col[t - 1])
        return col

    def gates(self, second):                         # resample_events.m:24  rand(size(rates))
        u = np.full(self.n_initial, 0.5)
        for g, v in enumerate(self.gated):
            u[v - 1] = px.u01(px.gate_word(self._step(second, g)))
        if self.record:
            for v in range(1, self.n_initial + 1):
                self.tape.append((("gate", second, v), float(u[v - 1])))
        return u

    def dedisc_event(self, kind, second, var):       # dbn_hierarchical_sample.m:35
        # both kinds (re-emitted bin of a fired gate, new bin of a transition) read the variable's word of that second
        g = self.gated.index(int(var))
        k = self._step(second, g)
        kn = self._step(px.partner_second(second), g)
        if kind == "gate":
            assert px.gate_word(k) < self.G[int(var)]
        return self._rec(("event_dd", kind, second, var), px.dd_uniform(k, kn))

    def layer(self):                                 # UncorEncounterModel.m:260
        return self._rec(("layer",), px.u01(self._w(px.P_LAYER, 0)))


class MTStream(_Base):
    """Sequential MT19937 genrand_res53 stream in the reference's consumption order."""

    def __init__(self, seed: int, record: bool = False):
        super().__init__()
        self.rs = np.random.RandomState(int(seed))
        self.record = record

    def bind(self, n_initial, temporal_map, resample_rates):
        self.n_initial = int(n_initial)
        return self

    def begin(self, sample, attempt=0):
        pass

    def _n(self, ctx):
        return self._rec(ctx, self.rs.random_sample())

    def select_init(self, var):
        return self._n(("init_sel", var))

    def dedisc_init(self, var):
        return self._n(("init_dd", var))

    def select_trans(self, t, var_t1):
        return self._n(("trans_sel", t, var_t1))

    def trans_column(self, var_t1, t_max):
        col = self.rs.random_sample(t_max)
        if self.record:
            for t in range(1, t_max + 1):
                self.tape.append((("trans_sel", t, var_t1), float(col[t - 1])))
        return col

    def gates(self, second):
        u = self.rs.random_sample(self.n_initial)
        if self.record:
            for v in range(1, self.n_initial + 1):
                self.tape.append((("gate", second, v), float(u[v - 1])))
        return u

    def dedisc_event(self, kind, second, var):
        return self._n(("event_dd", kind, second, var))

    def layer(self):
        return self._n(("layer",))
