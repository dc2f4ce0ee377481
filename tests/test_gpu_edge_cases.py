"""Edge cases of the BA path on the GPU against the oracle: degenerate parameterisations, ragged and extreme inputs."""
import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic

pytestmark = pytest.mark.gpu


def _scene(**kw):
    args = dict(n_cams=6, n_points=60, track_len=4, channels=16, seed=3)
    args.update(kw)
    prob, gt = synthetic.make_ba_scene(**args)
    ic = _capi.default_interp()
    prob.refs = O.refs_compute(prob, ic)[0]
    return prob, ic


def _same(a, b, tol=1e-6):
    assert np.abs(a.qvec - b.qvec).max() < tol and np.abs(a.tvec - b.tvec).max() < tol
    assert np.abs(a.xyz - b.xyz).max() < tol
    assert np.abs(a.cam_params[:, 1:] - b.cam_params[:, 1:]).max() < tol
    assert np.abs(a.cam_params[:, 0] / b.cam_params[:, 0] - 1).max() < tol


@pytest.mark.parametrize("inner", [0, 1])
def test_pure_triangulation_no_camera_unknowns(inner):
    """BASELINE configs[3] shape (ETH3D triangulation: refine_extrinsics / focal / extra all False,
    configs/pixsfm_eth3d.yaml): the reduced camera system is empty, every point is its own 3x3 problem."""
    prob, ic = _scene()
    prob.pose_const[:] = 1
    prob.cam_const_mask[:] = 0xFFFFFFFF
    so = _capi.default_ba_options(use_inner_iterations=inner, max_num_iterations=10)
    a, b = prob.copy(), prob.copy()
    s_ref = O.ba_solve(a, ic, so); s_gpu = _engine.ba_run(b, ic, so)
    assert s_gpu["num_iterations"] == s_ref["num_iterations"]
    assert abs(s_gpu["final_cost"] - s_ref["final_cost"]) <= 1e-6 * s_ref["final_cost"]
    assert s_ref["final_cost"] < s_ref["initial_cost"]
    _same(a, b)
    assert np.array_equal(b.qvec, prob.qvec) and np.array_equal(b.cam_params, prob.cam_params)


def test_single_free_image_against_constant_points():
    """Query-BA shape (localization/src/single_query_bundle_optimizer.h:96-127): one image's pose (and focal) free,
    every 3D point constant -> no point blocks, a 7-unknown camera system."""
    prob, ic = _scene(n_cams=4, n_points=80)
    prob.point_const[:] = 1
    prob.pose_const[:] = 1; prob.pose_const[2] = 0
    prob.tvec_const_mask[:] = 0
    so = _capi.default_ba_options(use_inner_iterations=0, max_num_iterations=12)
    a, b = prob.copy(), prob.copy()
    s_ref = O.ba_solve(a, ic, so); s_gpu = _engine.ba_run(b, ic, so)
    assert s_gpu["num_iterations"] == s_ref["num_iterations"]
    assert abs(s_gpu["final_cost"] - s_ref["final_cost"]) <= 1e-6 * s_ref["final_cost"]
    _same(a, b)
    assert np.array_equal(b.xyz, prob.xyz)
    assert np.abs(b.qvec[2] - prob.qvec[2]).max() > 0


def test_single_observation_tracks_and_unobserved_cameras():
    """Under-determined points (track length 1) and cameras without any observation: LM damping keeps the point
    blocks invertible, unobserved blocks stay where they are — same behaviour as the oracle."""
    prob, ic = _scene(n_cams=7, n_points=40, track_len=3)
    keep = np.ones(prob.n_obs, bool)
    first = np.r_[0, np.cumsum(np.bincount(prob.obs_pt, minlength=40))[:-1]]
    for p in range(0, 40, 4):
        keep[first[p] + 1:first[p] + 3] = False           # these points keep a single observation
    keep[prob.obs_img == 6] = False                        # camera 6 is never observed
    sub = _capi.BAProblem(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask,
                          qvec=prob.qvec, tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const,
                          tvec_const_mask=prob.tvec_const_mask, xyz=prob.xyz, point_const=prob.point_const,
                          obs_img=prob.obs_img[keep], obs_pt=prob.obs_pt[keep],
                          patches=np.ascontiguousarray(prob.patches[keep]), corner=prob.corner[keep], scale=prob.scale[keep],
                          refs=prob.refs)
    so = _capi.default_ba_options(use_inner_iterations=1, max_num_iterations=8)
    a, b = sub.copy(), sub.copy()
    s_ref = O.ba_solve(a, ic, so); s_gpu = _engine.ba_run(b, ic, so)
    assert s_gpu["num_iterations"] == s_ref["num_iterations"]
    assert abs(s_gpu["final_cost"] - s_ref["final_cost"]) <= 1e-6 * s_ref["final_cost"]
    # single-observation points are free along their ray: the solution is only determined up to the LM damping,
    # so compare at the reference's own BA tolerance (bundle_optimizer_test.cc:52)
    _same(a, b, 1e-4)


def test_observations_far_outside_their_patch_are_clamped_not_fatal():
    prob, ic = _scene(n_points=30)
    prob.xyz[::5] += 0.2          # projections land tens of pixels outside the 16x16 patch: per-tap clamping
    so = _capi.default_ba_options(use_inner_iterations=0, max_num_iterations=5)
    ref = O.ba_evaluate(prob, ic, so, residuals=True)
    got = _engine.BAHandle(prob, ic, so).evaluate(residuals=True)
    assert np.isfinite(got["cost"]) and abs(got["cost"] - ref["cost"]) <= 1e-11 * ref["cost"]
    assert np.abs(got["residuals"] - ref["residuals"]).max() <= 1e-12 * np.abs(ref["residuals"]).max()
    a, b = prob.copy(), prob.copy()
    s_ref = O.ba_solve(a, ic, so); s_gpu = _engine.ba_run(b, ic, so)
    assert abs(s_gpu["final_cost"] - s_ref["final_cost"]) <= 1e-6 * s_ref["final_cost"]


def test_track_longer_than_the_reference_extraction_limit_is_refused():
    prob, gt = synthetic.make_ba_scene(n_cams=4, n_points=2, track_len=4, channels=16, seed=1)
    rep = 70                                         # 4 * 70 = 280 observations of point 0 (> 256)
    idx = np.r_[np.tile(np.where(prob.obs_pt == 0)[0], rep), np.where(prob.obs_pt == 1)[0]]
    big = _capi.BAProblem(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask,
                          qvec=prob.qvec, tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const,
                          tvec_const_mask=prob.tvec_const_mask, xyz=prob.xyz, point_const=prob.point_const,
                          obs_img=prob.obs_img[idx], obs_pt=prob.obs_pt[idx], patches=np.ascontiguousarray(prob.patches[idx]),
                          corner=prob.corner[idx], scale=prob.scale[idx])
    with pytest.raises((ValueError, _capi.PxrError)):
        _engine.refs_compute(big, _capi.default_interp())


def test_two_level_refine_multilevel_runs_coarse_to_fine():
    import copy
    from pixsfm import bundle_adjustment as ba_pkg, features
    from recon_util import make_reconstruction
    rec, fm1, _, _ = make_reconstruction(n_cams=5, n_points=40, track_len=3, channels=16, seed=8)
    fm = features.FeatureManager([16, 16], np.float16)
    fm.fsets[0] = fm1.fset(0); fm.fsets[1] = copy.deepcopy(fm1.fset(0))
    out = ba_pkg.BundleAdjuster.create({"optimizer": {"solver": {"max_num_iterations": 5}}}).refine_multilevel(rec, fm)
    assert len(out["summary"]) == 2 and len(out["references"]) == 2
    # levels are processed in reverse index order (util/misc.py:19-23); the second pass starts where the first ended
    assert out["summary"][1].initial_cost <= out["summary"][0].initial_cost


def test_persistent_cholesky_bail_out_falls_back_to_per_panel_launches(monkeypatch):
    """If the persistent tile-DAG kernel ever gives up on a wait (its CTAs not co-resident: GPU shared / partitioned),
    the step is redone on the launch-per-panel path and the solve still matches the oracle."""
    prob, ic = _scene(n_cams=40, n_points=400, track_len=5)
    so = _capi.default_ba_options(use_inner_iterations=0, max_num_iterations=6)
    a, b = prob.copy(), prob.copy()
    s_ref = O.ba_solve(a, ic, so)
    monkeypatch.setenv("PXR_CHOL_TEST_ABORT", "1")
    s_gpu = _engine.ba_run(b, ic, so)
    assert s_gpu["num_iterations"] == s_ref["num_iterations"]
    assert abs(s_gpu["final_cost"] - s_ref["final_cost"]) <= 1e-6 * s_ref["final_cost"]
    _same(a, b)
