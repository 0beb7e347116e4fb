"""The fixtures matlab/verify_parity.m replays into the real reference (tests/golden/matlab_parity/) must be what the oracle
produces today: the tape in the reference's rand order and the outputs computed from it.  A stale fixture would make the
MATLAB-side verdict meaningless."""
import os
import shutil
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_matlab_parity as MP  # noqa: E402

FX = os.path.join(HERE, "golden", "matlab_parity")


@pytest.mark.parametrize("case", sorted(MP.CASES))
def test_fixture_matches_the_oracle(model_paths, tmp_path, case):
    fname = MP.CASES[case][0]
    stem = os.path.splitext(os.path.basename(fname))[0]
    dst = tmp_path / fname
    dst.parent.mkdir(parents=True, exist_ok=True)
    shutil.copy(model_paths[stem], dst)
    d = MP.build(case, model_dir=str(tmp_path))
    for k, a in d.items():
        f = os.path.join(FX, "%s_%s.txt" % (case, k))
        got = np.loadtxt(f, ndmin=2) if os.path.getsize(f) else np.zeros((0, 4))
        want = np.atleast_2d(a) if k != "tape" else np.asarray(a)[:, None]
        assert got.shape == want.shape or (got.size == 0 and want.size == 0), (case, k)
        assert np.array_equal(got, want), (case, k)


def test_tape_has_one_entry_per_reference_rand_call(model_paths):
    """per attempt of one uncor track the reference consumes n_free initial draws, t_max per dynamic variable
    (fast branch: rand(t_max,1)), T * n_initial gate draws, and one draw per de-discretised value."""
    from oracle.drivers import uncor_sample
    from oracle.em_read import em_read
    from oracle.uniforms import KeyedPhilox
    p = em_read(model_paths["uncor_1200code_v2p1"])
    U = KeyedPhilox(1, record=True)
    out = uncor_sample(p, 1, 50, U)
    kinds = [c[0] for c, _ in U.tape]
    assert out[0].attempts == 1
    assert kinds.count("init_sel") == p.n_initial and kinds.count("trans_sel") == 3 * 50 and kinds.count("gate") == 50 * p.n_initial
    n_dd = sum(1 for i in range(p.n_initial) if len(p.boundaries[i]) and not (p.zero_bins[i] and out[0].initial_bins[i] == p.zero_bins[i][0]))
    assert kinds.count("init_dd") == n_dd
    # order: all initial selects, then the three rand(t_max,1) columns, then gates second by second, then values
    first = {k: kinds.index(k) for k in ("init_sel", "trans_sel", "gate", "init_dd")}
    assert first["init_sel"] < first["trans_sel"] < first["gate"] < first["init_dd"]
