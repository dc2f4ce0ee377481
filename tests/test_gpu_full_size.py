"""BASELINE configs[2] at FULL size — 200 cameras / 50 000 points / 500 000 observations, 128-ch fp16 16x16 patches
(32.8 GB, generated on the device exactly as bench.py does) — checked through properties that do not need an oracle
run of the whole problem:
  * a random sample of whole points (all their observations) pulled back to the host: the per-observation outputs of
    the full-size launch (||r||^2, G^T r, G^T G, projections) and the references equal the oracle's on that sample;
  * the cost is additive over any partition of the points, and the per-observation outputs of a partition are the
    slices of the full launch, bit for bit;
  * evaluation is idempotent (bit-identical when repeated);
  * LM steps on the full problem never increase the cost, keep the gauge (constant pose / constant tvec component) and
    unit quaternions."""
import ctypes as C
import sys

import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _engine

pytestmark = pytest.mark.gpu


def _sub_problem(args, g, prob, d_patches, refs, p0, p1):
    """points [p0, p1) of the full problem on the SAME device slab (observations index their patches)"""
    L = min(args.track, args.cams)
    o0, o1 = p0 * L, p1 * L
    n_obs = len(g["obs_pt"])
    return _capi.BAProblem(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask,
                           qvec=prob.qvec, tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const,
                           tvec_const_mask=prob.tvec_const_mask, xyz=g["xyz"][p0:p1], point_const=np.zeros(p1 - p0, np.uint8),
                           obs_img=g["obs_img"][o0:o1], obs_pt=g["obs_pt"][o0:o1] - p0, patches=d_patches,
                           corner=g["corners"], scale=g["scale"], refs=refs[p0:p1], obs_patch=np.arange(o0, o1, dtype=np.int64),
                           patches_on_device=True, patch_shape=(n_obs, args.ps, args.ps, args.channels), patch_dtype=0)


@pytest.fixture(scope="module")
def s3():
    import bench
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        args = bench.parse()
    finally:
        sys.argv = argv
    assert (args.cams, args.points, args.channels, args.ps) == (200, 50000, 128, 16)
    ctx = _capi.default_context()
    g = bench.geometry(args, 0)
    n_obs = len(g["obs_pt"])
    assert n_obs == 500000
    d_patches = _engine.synth_patches_device(n_obs, args.ps, args.channels, g["uv0"], g["obs_pt"], seed=args.seed * 7919,
                                             noise=0.01, ctx=ctx)
    prob = bench.make_problem(args, g, d_patches, True)
    ic = _capi.default_interp()
    refs, src = _engine.refs_compute(prob, ic, ctx=ctx)
    prob.refs = refs
    so = _capi.default_ba_options(max_num_iterations=4)
    full = _engine.BAHandle(prob, ic, so, ctx=ctx)
    ev = full.evaluate()
    yield dict(args=args, g=g, prob=prob, d_patches=d_patches, ic=ic, so=so, refs=refs, src=src, handle=full, ev=ev, ctx=ctx)
    del full
    _engine.device_free(d_patches, ctx)


def test_sampled_points_of_the_full_launch_match_the_oracle(s3):
    args, g, prob, ev = s3["args"], s3["g"], s3["prob"], s3["ev"]
    L = min(args.track, args.cams)
    rng = np.random.default_rng(5)
    pts = np.sort(rng.choice(args.points, 160, replace=False))
    obs = (pts[:, None] * L + np.arange(L)[None, :]).reshape(-1)
    pbytes = args.ps * args.ps * args.channels * 2
    host = np.empty((len(obs), args.ps, args.ps, args.channels), np.float16)
    for k, o in enumerate(obs):          # pull exactly the sampled patches out of the 32.8 GB slab
        _engine.memcpy_d2h(host[k], s3["d_patches"] + int(o) * pbytes, pbytes, s3["ctx"])
    small = _capi.BAProblem(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask,
                            qvec=prob.qvec, tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const,
                            tvec_const_mask=prob.tvec_const_mask, xyz=g["xyz"][pts], point_const=np.zeros(len(pts), np.uint8),
                            obs_img=g["obs_img"][obs], obs_pt=np.repeat(np.arange(len(pts), dtype=np.int64), L), patches=host,
                            corner=g["corners"][obs], scale=g["scale"][obs])
    # references: the IRLS robust mean + closest observation of the full-size extraction
    refs_cpu, src_cpu = O.refs_compute(small, s3["ic"])
    # source observation (an index into the observation list), taken relative to the point's first observation
    assert np.array_equal(src_cpu - np.arange(len(pts)) * L, s3["src"][pts] - pts * L)
    assert np.abs(refs_cpu - s3["refs"][pts]).max() < 1e-12
    small.refs = np.ascontiguousarray(s3["refs"][pts])
    cpu = O.ba_evaluate(small, s3["ic"], s3["so"])
    for key, tol in (("xy", 1e-9), ("sq_norm", 1e-12), ("gtr", 1e-11), ("gtg", 1e-10)):
        a, b = ev[key][obs], cpu[key]
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), key


def test_cost_is_additive_over_partitions_and_outputs_are_slices(s3):
    args, g, prob, ev = s3["args"], s3["g"], s3["prob"], s3["ev"]
    L = min(args.track, args.cams)
    cuts = [0, 12345, 30001, args.points]
    total = 0.0
    for p0, p1 in zip(cuts[:-1], cuts[1:]):
        sub = _sub_problem(args, g, prob, s3["d_patches"], s3["refs"], p0, p1)
        e = _engine.BAHandle(sub, s3["ic"], s3["so"], ctx=s3["ctx"]).evaluate()
        total += e["cost"]
        for key in ("sq_norm", "gtr", "gtg", "xy"):
            assert np.array_equal(e[key], ev[key][p0 * L:p1 * L]), key
    assert abs(total - ev["cost"]) <= 1e-12 * ev["cost"]
    # and the cost is what its definition says: sum of 0.5 * rho(||r||^2), Cauchy(0.25)
    s = ev["sq_norm"]
    a2 = 0.25 ** 2
    assert abs(0.5 * np.sum(a2 * np.log1p(s / a2)) - ev["cost"]) <= 1e-12 * ev["cost"]


def test_evaluation_is_idempotent(s3):
    again = s3["handle"].evaluate()
    for key in ("sq_norm", "gtr", "gtg", "xy"):
        assert np.array_equal(again[key], s3["ev"][key]), key
    assert abs(again["cost"] - s3["ev"]["cost"]) <= 1e-13 * s3["ev"]["cost"]      # cost reduction order may differ


def test_lm_steps_on_the_full_problem_decrease_the_cost_and_keep_the_gauge(s3):
    prob, h = s3["prob"], s3["handle"]
    before = dict(q=prob.qvec.copy(), t=prob.tvec.copy(), cam=prob.cam_params.copy(), xyz=prob.xyz.copy())
    s = h.iterate(4)
    its = s["iterations"]
    assert len(its) >= 4 and s["kernel_launches"] > 0
    cost = its[0]["cost"]
    assert abs(cost - s3["ev"]["cost"]) <= 1e-12 * cost
    for it in its[1:]:
        if it["step_is_successful"]:
            assert it["cost"] < cost
            cost = it["cost"]
        else:
            assert it["cost"] <= cost * (1 + 1e-15)
    assert cost < its[0]["cost"]
    try:
        h.read_params()          # writes into the problem's arrays
        q, t = prob.qvec, prob.tvec
        assert np.all(np.isfinite(q)) and np.all(np.isfinite(t)) and np.all(np.isfinite(prob.xyz))
        assert np.abs(np.linalg.norm(q, axis=1) - 1.0).max() < 1e-12
        assert np.array_equal(q[0], before["q"][0]) and np.array_equal(t[0], before["t"][0])      # constant pose
        assert t[1][0] == before["t"][1][0] and not np.array_equal(t[1], before["t"][1])         # constant tvec component
        assert np.array_equal(prob.cam_params[:, 1:3], before["cam"][:, 1:3])                    # principal point not refined
        assert not np.array_equal(prob.xyz, before["xyz"])
    finally:
        prob.qvec[:] = before["q"]; prob.tvec[:] = before["t"]; prob.cam_params[:] = before["cam"]; prob.xyz[:] = before["xyz"]
