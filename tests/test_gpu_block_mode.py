"""Block mode (pxr_block.cuh / pxr_ba_block.cu): the multi-GPU LM iteration — image-block assembly into one packed
buffer, deterministic replicated solve, ONE host synchronisation — run on ONE GPU (PXR_BLOCK_ASSEMBLY=1, no
communication) against the single-GPU driver and against the oracle.  The N>1 behaviour itself is checked by
scripts/check_multi_gpu.py under torchrun."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic

pytestmark = pytest.mark.gpu


class block_mode:
    def __init__(self, sparse=False):
        self.sparse = sparse

    def __enter__(self):
        os.environ["PXR_BLOCK_ASSEMBLY"] = "1"
        if self.sparse:
            os.environ["PXR_PCG_SPARSE"] = "1"

    def __exit__(self, *a):
        os.environ.pop("PXR_BLOCK_ASSEMBLY", None)
        os.environ.pop("PXR_PCG_SPARSE", None)


def _scene(**kw):
    args = dict(n_cams=8, n_points=120, track_len=4, channels=16, seed=11)
    args.update(kw)
    prob, _ = synthetic.make_ba_scene(**args)
    ic = _capi.default_interp()
    prob.refs, _ = O.refs_compute(prob, ic)
    return prob, ic


def _same_trajectory(a, b, tol):
    assert len(a["iterations"]) == len(b["iterations"])
    for x, y in zip(a["iterations"], b["iterations"]):
        assert x["step_is_successful"] == y["step_is_successful"] and x["step_is_valid"] == y["step_is_valid"]
        assert abs(x["cost"] - y["cost"]) <= tol * abs(y["cost"])
        if np.isfinite(x["gradient_max_norm"]) and np.isfinite(y["gradient_max_norm"]):
            assert abs(x["gradient_max_norm"] - y["gradient_max_norm"]) <= 1e-6 * abs(y["gradient_max_norm"]) + 1e-15


@pytest.mark.parametrize("inner,shared", [(0, False), (1, False), (0, True), (1, True)])
def test_block_mode_equals_the_single_gpu_driver_and_the_oracle(inner, shared):
    prob, ic = _scene(shared_camera=shared)
    so = _capi.default_ba_options(max_num_iterations=10, use_inner_iterations=inner)
    p_ref, p_blk, p_cpu = prob.copy(), prob.copy(), prob.copy()
    s_ref = _engine.ba_run(p_ref, ic, so)
    with block_mode():
        s_blk = _engine.ba_run(p_blk, ic, so)
    s_cpu = O.ba_solve(p_cpu, ic, so)
    _same_trajectory(s_blk, s_ref, 1e-10)
    _same_trajectory(s_blk, s_cpu, 1e-6)
    assert abs(s_blk["final_cost"] - s_ref["final_cost"]) <= 1e-10 * s_ref["final_cost"]
    for name in ("qvec", "tvec", "xyz", "cam_params"):
        assert np.abs(getattr(p_blk, name) - getattr(p_ref, name)).max() < 1e-9
        assert np.abs(getattr(p_blk, name) - getattr(p_cpu, name)).max() < 1e-6
    assert s_blk["num_inner_iteration_steps"] == s_ref["num_inner_iteration_steps"]


@pytest.mark.parametrize("sparse", [False, True])
def test_block_mode_iterative_schur(sparse):
    prob, ic = _scene(n_cams=12, n_points=200, seed=5)
    so = _capi.default_ba_options(max_num_iterations=8, use_inner_iterations=0, linear_solver=3)
    p_ref, p_blk = prob.copy(), prob.copy()
    s_ref = _engine.ba_run(p_ref, ic, so)                       # dense PCG, single-GPU driver
    with block_mode(sparse):
        s_blk = _engine.ba_run(p_blk, ic, so)
    _same_trajectory(s_blk, s_ref, 1e-8)
    assert np.abs(p_blk.xyz - p_ref.xyz).max() < 1e-7 and np.abs(p_blk.qvec - p_ref.qvec).max() < 1e-7
    # the replicated solve is a deterministic function of the blocks: two runs are bit-identical
    p_again = prob.copy()
    with block_mode(sparse):
        s_again = _engine.ba_run(p_again, ic, so)
    assert [i["linear_solver_iterations"] for i in s_again["iterations"]] == [i["linear_solver_iterations"] for i in s_blk["iterations"]]


def test_block_mode_step_matches_debug_linearize_of_the_dense_path():
    prob, ic = _scene(seed=3)
    so = _capi.default_ba_options(use_inner_iterations=0)
    lin = O.ba_linearize(prob, ic, so, radius=1e4)
    ref = _engine.BAHandle(prob.copy(), ic, so).debug_linearize(lin["nc"], lin["nl"], radius=1e4)
    with block_mode():
        blk = _engine.BAHandle(prob.copy(), ic, so).debug_linearize(lin["nc"], lin["nl"], radius=1e4, dense=False)
    assert abs(blk["cost"] - ref["cost"]) <= 1e-13 * ref["cost"]
    assert np.abs(blk["gc"] - ref["gc"]).max() <= 1e-10 * np.abs(ref["gc"]).max()
    assert np.abs(blk["delta"] - ref["delta"]).max() <= 1e-9 * np.abs(ref["delta"]).max()
    assert np.abs(blk["delta"] - lin["delta"]).max() <= 1e-5 * np.abs(lin["delta"]).max()
    assert abs(blk["model_cost_change"] - ref["model_cost_change"]) <= 1e-10 * abs(ref["model_cost_change"])


def test_block_mode_constant_cameras_and_gradient_tolerance():
    """no camera columns at all (triangulation-style BA): no reduced system, no collective; and a gradient tolerance that
    stops the solve"""
    prob, ic = _scene(refine_extrinsics=False, refine_focal=False, refine_extra=False)
    so = _capi.default_ba_options(max_num_iterations=6, use_inner_iterations=0)
    p_ref, p_blk = prob.copy(), prob.copy()
    s_ref = _engine.ba_run(p_ref, ic, so)
    with block_mode():
        s_blk = _engine.ba_run(p_blk, ic, so)
    _same_trajectory(s_blk, s_ref, 1e-10)
    assert np.abs(p_blk.xyz - p_ref.xyz).max() < 1e-10
    # gradient tolerance (ceres: max-norm of x - Plus(x, -g)): the oracle stops at the same iteration
    so2 = _capi.default_ba_options(max_num_iterations=30, use_inner_iterations=0, gradient_tolerance=2e-2)
    prob2, _ = _scene()
    p_ref, p_blk, p_cpu = prob2.copy(), prob2.copy(), prob2.copy()
    s_ref = _engine.ba_run(p_ref, ic, so2)
    with block_mode():
        s_blk = _engine.ba_run(p_blk, ic, so2)
    s_cpu = O.ba_solve(p_cpu, ic, so2)
    assert s_cpu["termination_type"] == s_ref["termination_type"] == s_blk["termination_type"] == 0
    assert s_cpu["num_iterations"] == s_ref["num_iterations"] == s_blk["num_iterations"] < 31
    assert abs(s_ref["iterations"][-1]["gradient_max_norm"] - s_cpu["iterations"][-1]["gradient_max_norm"]) < 1e-6
