"""Window residency of the patch slab (csrc/pxr_resident.cuh): pxr_ba_run brings over a W x W tap window per observation
(packed on the host, pinned or pageable source), fetches whole patches for observations that leave it, and must give
the results of the same solve on fully uploaded patches: the same taps go through the same arithmetic.  In deterministic
mode (fixed-order sums) that means bit for bit; with the default atomics what is left is their run-to-run noise."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic

pytestmark = pytest.mark.gpu


def _scene(**kw):
    args = dict(n_cams=8, n_points=150, track_len=5, channels=128, seed=21)
    args.update(kw)
    prob, _ = synthetic.make_ba_scene(**args)
    ic = _capi.default_interp()
    prob.refs, _ = O.refs_compute(prob, ic)
    return prob, ic


def _pinned_copy(prob):
    pin = _engine.PinnedArray(prob.patches.shape, prob.patches.dtype)
    pin.array[...] = prob.patches
    q = prob.copy()
    q.patches = pin.array
    q._patches_ptr = pin.array.ctypes.data
    return q, pin


def _run(prob, ic, so, window):
    q, pin = _pinned_copy(prob)
    old = os.environ.get("PXR_RESIDENT_WINDOW")
    os.environ["PXR_RESIDENT_WINDOW"] = str(window)
    try:
        s = _engine.ba_run(q, ic, so)
    finally:
        if old is None:
            os.environ.pop("PXR_RESIDENT_WINDOW", None)
        else:
            os.environ["PXR_RESIDENT_WINDOW"] = old
    out = (s, q.qvec.copy(), q.tvec.copy(), q.xyz.copy(), q.cam_params.copy())
    pin.close()
    return out


def _identical(a, b, exact):
    """exact: the solves ran in deterministic mode (fixed-order sums) -> every bit must agree; otherwise the fp64 atomics
    of the assembly leave run-to-run noise that the LM iterations amplify (a few 1e-11 on the final cost at most)"""
    sa, sb = a[0], b[0]
    assert len(sa["iterations"]) == len(sb["iterations"])
    assert sa["iterations"][0]["cost"] == sb["iterations"][0]["cost"]          # the first evaluation has no atomics in it: bit-equal
    tol = 0.0 if exact else 1e-9
    for x, y in zip(sa["iterations"], sb["iterations"]):
        assert x["step_is_successful"] == y["step_is_successful"] and x["step_is_valid"] == y["step_is_valid"]
        for key in ("cost", "cost_change", "step_norm", "relative_decrease", "trust_region_radius"):
            assert abs(x[key] - y[key]) <= tol * abs(y[key]) + (0.0 if exact else 1e-13), key
    assert abs(sa["final_cost"] - sb["final_cost"]) <= tol * sb["final_cost"]
    assert sa["num_inner_iteration_steps"] == sb["num_inner_iteration_steps"]
    for u, v in zip(a[1:], b[1:]):
        assert np.array_equal(u, v) if exact else np.abs(u - v).max() <= 1e-8


@pytest.mark.parametrize("deterministic", [1, 0])
@pytest.mark.parametrize("inner", [0, 1])
@pytest.mark.parametrize("window", [8, 4])
def test_windowed_solve_is_bit_identical_to_the_fully_resident_one(inner, window, deterministic):
    # a rough start (5x the usual perturbation) so that points do leave small windows
    prob, ic = _scene(rot_sigma_deg=0.06, pt_sigma=0.012)
    so = _capi.default_ba_options(max_num_iterations=12, use_inner_iterations=inner, deterministic=deterministic)
    full = _run(prob, ic, so, 0)
    win = _run(prob, ic, so, window)
    assert full[0]["resident_window"] == 0 and win[0]["resident_window"] == window
    _identical(win, full, exact=bool(deterministic))
    pbytes = prob.patches.nbytes
    assert full[0]["h2d_bytes"] >= pbytes
    # every observation got its window, violators their whole patch: far fewer bytes than the slab unless W = 4 met a
    # rough start (then many patches are refetched) — and never more than window + slab
    expect = prob.n_obs * window * window * prob.channels * 2 + win[0]["resident_refetched"] * 16 * 16 * prob.channels * 2
    assert abs(win[0]["h2d_bytes"] - expect) < 2e6
    if window == 4:
        assert win[0]["resident_refetched"] > 0 and win[0]["resident_passes_repeated"] > 0
    if window == 8:
        assert win[0]["h2d_bytes"] < 0.6 * pbytes


def test_windowed_solve_matches_the_oracle():
    prob, ic = _scene(seed=4)
    so = _capi.default_ba_options(max_num_iterations=8, use_inner_iterations=1)
    win = _run(prob, ic, so, 8)
    p_cpu = prob.copy()
    s_cpu = O.ba_solve(p_cpu, ic, so)
    assert abs(win[0]["final_cost"] - s_cpu["final_cost"]) <= 1e-6 * s_cpu["final_cost"]
    assert np.abs(win[3] - p_cpu.xyz).max() < 1e-6 and np.abs(win[1] - p_cpu.qvec).max() < 1e-6


def test_shared_patches_and_pageable_sources_fall_back_to_what_is_safe():
    prob, ic = _scene(n_points=60, seed=9)
    so = _capi.default_ba_options(max_num_iterations=6, use_inner_iterations=0)
    # pageable numpy source: same window upload (the host threads pack the windows out of whatever memory it is)
    p1, p0 = prob.copy(), prob.copy()
    s1 = _engine.ba_run(p1, ic, so)
    assert s1["resident_window"] == 8 and s1["h2d_bytes"] < 0.5 * prob.patches.nbytes
    os.environ["PXR_RESIDENT_WINDOW"] = "0"
    try:
        s0 = _engine.ba_run(p0, ic, so)
    finally:
        os.environ.pop("PXR_RESIDENT_WINDOW", None)
    assert s0["resident_window"] == 0 and abs(s1["final_cost"] - s0["final_cost"]) <= 1e-9 * s0["final_cost"]
    assert np.abs(p1.xyz - p0.xyz).max() < 1e-8
    # two observations reading ONE patch (a dense map does that): such patches come over whole, the rest as windows
    q, pin = _pinned_copy(prob)
    n = q.n_obs
    obs_patch = np.arange(n, dtype=np.int64)
    obs_patch[1] = 0                       # observation 1 now reads patch 0 as well
    q.obs_patch = obs_patch
    q.corner[1] = q.corner[0]; q.scale[1] = q.scale[0]
    q2 = q.copy(); q2.patches = pin.array; q2._patches_ptr = pin.array.ctypes.data
    os.environ["PXR_RESIDENT_WINDOW"] = "0"
    try:
        s_full = _engine.ba_run(q, ic, so)
    finally:
        os.environ.pop("PXR_RESIDENT_WINDOW", None)
    s_win = _engine.ba_run(q2, ic, so)
    assert s_win["resident_window"] == 8 and s_full["resident_window"] == 0
    assert abs(s_win["final_cost"] - s_full["final_cost"]) <= 1e-9 * s_full["final_cost"] and np.abs(q.xyz - q2.xyz).max() < 1e-8
    pin.close()
