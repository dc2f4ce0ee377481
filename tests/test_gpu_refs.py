"""GPU parity of the reference extraction (IRLS robust mean + closest observation) vs the oracle."""
import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("channels,track_len", [(128, 4), (128, 9), (16, 5), (64, 3)])
def test_references_match_oracle(channels, track_len):
    prob, gt = synthetic.make_ba_scene(n_cams=10, n_points=80, track_len=track_len, channels=channels, seed=11)
    ic = _capi.default_interp()
    r_cpu, s_cpu = O.refs_compute(prob, ic, iters=100)
    r_gpu, s_gpu = _engine.refs_compute(prob, ic, iters=100)
    assert np.array_equal(s_gpu, s_cpu)          # source observation index: bit-exact
    assert np.abs(r_gpu - r_cpu).max() < 1e-12   # the descriptor of that observation


def test_references_with_outlier_observation():
    prob, gt = synthetic.make_ba_scene(n_cams=8, n_points=40, track_len=6, channels=128, seed=12)
    # corrupt one observation per point: the robust mean must not pick it
    rng = np.random.default_rng(0)
    for p in range(40):
        o = p * 6 + int(rng.integers(0, 6))
        prob.patches[o] = rng.normal(0, 0.1, prob.patches[o].shape).astype(np.float16)
    ic = _capi.default_interp()
    r_cpu, s_cpu = O.refs_compute(prob, ic)
    r_gpu, s_gpu = _engine.refs_compute(prob, ic)
    assert np.array_equal(s_gpu, s_cpu)
    assert np.abs(r_gpu - r_cpu).max() < 1e-12
