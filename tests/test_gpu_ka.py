"""GPU parity of featuremetric keypoint adjustment against the oracle (bounded LM + line search)."""
import numpy as np
import pytest

import oracle_lib as O
from ka_util import make_ka_problem
from pixsfm._pixsfm import _capi, _engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("channels,kp_sigma,bound", [(128, 1.0, 4.0), (16, 0.7, 4.0), (128, 2.5, 1.0)])
def test_keypoint_adjustment_matches_oracle(channels, kp_sigma, bound):
    prob, sc, lab = make_ka_problem(n_images=6, n_tracks=40, track_len=4, channels=channels, seed=2, kp_sigma=kp_sigma,
                                    bound=bound, max_per_problem=20)
    ic = _capi.default_interp(); so = _capi.default_ka_options()
    p_cpu, p_gpu = prob.copy(), prob.copy()
    c0, c1 = O.ka_solve(p_cpu, ic, so)
    s = _engine.ka_run(p_gpu, ic, so)
    assert abs(s["initial_cost"] - c0) <= 1e-9 * c0
    assert abs(s["final_cost"] - c1) <= 1e-6 * c1
    assert c1 < c0
    assert np.abs(p_gpu.keypoints - p_cpu.keypoints).max() < 1e-5
    # roots did not move, the others did and got closer to the truth on average
    roots = lab["roots"].astype(bool)
    assert np.array_equal(p_gpu.keypoints[roots], prob.keypoints[roots])
    moved = ~roots
    assert np.abs(p_gpu.keypoints[moved] - prob.keypoints[moved]).max() > 1e-3
    assert s["kernel_launches"] > 0


def test_ka_unconstrained_and_single_problem():
    prob, sc, lab = make_ka_problem(n_images=5, n_tracks=12, track_len=4, channels=128, seed=4, kp_sigma=0.5,
                                    bound=-1.0, max_per_problem=1000)
    prob.patches_are_sparse = False  # dense maps and bound<=0: no box constraints (keypoint_optimizer.h:128)
    assert prob.n_problems == 1
    ic = _capi.default_interp(); so = _capi.default_ka_options()
    p_cpu, p_gpu = prob.copy(), prob.copy()
    c0, c1 = O.ka_solve(p_cpu, ic, so)
    s = _engine.ka_run(p_gpu, ic, so)
    assert abs(s["final_cost"] - c1) <= 1e-6 * c1
    assert np.abs(p_gpu.keypoints - p_cpu.keypoints).max() < 1e-5


@pytest.mark.parametrize("channels,n_queries,bound", [(128, 3, 4.0), (16, 1, 4.0), (128, 2, -1.0)])
def test_query_keypoint_adjustment_matches_oracle(channels, n_queries, bound):
    """Query KA (localization/src/single_query_keypoint_optimizer.h:84-203): keypoints against FIXED reference
    descriptors, all keypoints of a query in ONE trust region; block-diagonal normal equations on the GPU."""
    from ka_util import make_query_ka_problem
    prob, sc = make_query_ka_problem(n_queries=n_queries, bound=bound, n_images=6, n_tracks=40, track_len=4,
                                     channels=channels, seed=6, kp_sigma=1.0)
    assert len(prob.keypoints) == 160       # one query holds up to 160 keypoints: beyond the dense in-smem path (80)
    ic = _capi.default_interp(); so = _capi.default_ka_options()
    p_cpu, p_gpu = prob.copy(), prob.copy()
    c0, c1 = O.ka_solve(p_cpu, ic, so)
    s = _engine.ka_run(p_gpu, ic, so)
    assert abs(s["initial_cost"] - c0) <= 1e-9 * c0
    assert abs(s["final_cost"] - c1) <= 1e-6 * c1
    assert c1 < 0.9 * c0
    assert np.abs(p_gpu.keypoints - p_cpu.keypoints).max() < 1e-5
    assert np.abs(p_gpu.keypoints - prob.keypoints).max() > 1e-2
