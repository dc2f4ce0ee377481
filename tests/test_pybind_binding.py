"""The pybind11 binding of the C-ABI (pixel-perfect-sfm_b200/bindings/pxr_pybind.cc — the C++ form of what
INTEGRATION.md describes): it loads, reaches the same library as the ctypes mirror, gives the same answers for the
host-side algorithms, maps status codes to the reference's exception types and builds a problem description from a
dict of arrays.  Device entry points can only be checked here up to the "no device" error."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixel-perfect-sfm_b200", "bindings"))
pb = pytest.importorskip("_pxr_pybind")

from pixsfm._pixsfm import _capi, _engine      # noqa: E402
from pixsfm.util import synthetic               # noqa: E402
from conftest import HAS_GPU                    # noqa: E402


def test_same_library_and_host_algorithms_as_the_ctypes_mirror():
    assert pb.version() == _capi.load_lib().pxr_version()
    sc = synthetic.make_ka_scene(n_images=5, n_tracks=30, track_len=4, channels=16, seed=8)
    tl, scores, roots = _engine.graph_labels(sc["node_image"], sc["edge_src"], sc["edge_dst"], sc["edge_sim"])
    tl2 = pb.compute_track_labels(sc["node_image"], sc["edge_src"], sc["edge_dst"], sc["edge_sim"])
    assert tl2.dtype == np.int64 and np.array_equal(tl2, tl)
    sc2 = pb.compute_score_labels(len(tl), sc["edge_src"], sc["edge_dst"], sc["edge_sim"], tl2)
    assert np.array_equal(sc2, scores)
    assert np.array_equal(pb.compute_root_labels(tl2, sc2), roots)
    labels, n = pb.ka_problem_labels(tl2, 12)
    want, n_want = _engine.ka_problem_labels(tl, 12)
    assert n == n_want and np.array_equal(labels, want)
    w = np.array([5, 1, 9, 3, 3, 7], np.int64)
    assert np.array_equal(pb.shard_ka_problems(w, 3), _engine.ka_shard_plan(w, 3))
    obs_pt = np.repeat(np.arange(10, dtype=np.int64), 3)
    pbeg, obeg = pb.shard_points(10, obs_pt, 2)
    assert list(pbeg) == [0, 5, 10] and list(obeg) == [0, 15, 30]
    # lists and other integer dtypes are converted, as pybind11's numpy casters do for the reference's Eigen arguments
    assert np.array_equal(pb.compute_root_labels(list(tl), list(scores)), roots)


def test_status_codes_become_the_references_exception_types():
    with pytest.raises(ValueError):                       # unsorted observations: PXR_ERR_INVALID_ARGUMENT
        pb.shard_points(3, np.array([2, 1, 0], np.int64), 2)
    with pytest.raises(ValueError):
        pb.compute_track_labels(np.zeros(2, np.int32), np.zeros(1, np.int64), np.zeros(2, np.int64), np.zeros(1))
    if not HAS_GPU:
        with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
            pb.Context(0)


def test_problem_description_from_a_dict_of_arrays():
    prob, _ = synthetic.make_ba_scene(n_cams=4, n_points=20, track_len=3, channels=16, seed=2)
    d = dict(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask, qvec=prob.qvec,
             tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const, tvec_const_mask=prob.tvec_const_mask,
             xyz=prob.xyz, point_const=prob.point_const, obs_img=prob.obs_img, obs_pt=prob.obs_pt, patches=prob.patches,
             corner=prob.corner, scale=prob.scale)
    got = pb.describe_ba_problem(d)
    assert got == dict(n_cameras=len(prob.cam_model), n_images=4, n_points=20, n_obs=prob.n_obs, n_patches=prob.n_obs,
                       patch_dtype=0, ph=prob.ph, pw=prob.pw, channels=16, has_refs=False, has_obs_patch=False)
    d["refs"] = np.zeros((20, 16))
    assert pb.describe_ba_problem(d)["has_refs"] is True
    bad = dict(d); del bad["obs_pt"]
    with pytest.raises(ValueError, match="missing field 'obs_pt'"):
        pb.describe_ba_problem(bad)
    bad = dict(d, qvec=prob.qvec.astype(np.float32))      # in/out arrays are written by the library: no silent copies
    with pytest.raises(ValueError, match="qvec"):
        pb.describe_ba_problem(bad)
    bad = dict(d, patches=prob.patches.astype(np.int16))
    with pytest.raises(ValueError, match="float16, float32 or float64"):
        pb.describe_ba_problem(bad)
    assert pb.default_ba_options()["use_inner_iterations"] == 1 and pb.default_ka_options()["parameter_tolerance"] == 1e-5


@pytest.mark.gpu
def test_ba_run_through_the_pybind_binding_equals_the_ctypes_mirror():
    prob, _ = synthetic.make_ba_scene(n_cams=5, n_points=40, track_len=4, channels=16, seed=3)
    ic = _capi.default_interp()
    prob.refs = _engine.refs_compute(prob, ic)[0]
    a, b = prob.copy(), prob.copy()
    s1 = _engine.ba_run(a, ic, _capi.default_ba_options(max_num_iterations=6))
    ctx = pb.Context(-1)
    d = dict(cam_model=b.cam_model, cam_params=b.cam_params, cam_const_mask=b.cam_const_mask, qvec=b.qvec, tvec=b.tvec,
             img_cam=b.img_cam, pose_const=b.pose_const, tvec_const_mask=b.tvec_const_mask, xyz=b.xyz,
             point_const=b.point_const, obs_img=b.obs_img, obs_pt=b.obs_pt, patches=b.patches, corner=b.corner,
             scale=b.scale, refs=b.refs)
    refs2, src2 = pb.refs_compute(ctx, dict(d, refs=None))
    assert np.abs(refs2 - prob.refs).max() < 1e-12
    s2 = pb.ba_run(ctx, d, {}, {"max_num_iterations": 6})
    assert s2["num_successful_steps"] == s1["num_successful_steps"]
    assert abs(s2["final_cost"] - s1["final_cost"]) <= 1e-8 * s1["final_cost"]        # same library, same IR: only the
    assert np.abs(b.xyz - a.xyz).max() < 1e-7 and np.abs(b.qvec - a.qvec).max() < 1e-7   # atomics' summation order differs
    with pytest.raises(ValueError, match="unknown solver option"):
        pb.ba_run(ctx, d, {}, {"max_iterations": 6})


def test_problem_builder_through_pybind_equals_the_mirror():
    """BundleOptimizer::SetUp + Parameterize reached from C++ (pxr_problem_build): same arrays as the ctypes mirror"""
    import recon_util
    from pixsfm._pixsfm import _bundle_adjustment as ba
    from pixsfm._pixsfm._features import FeatureView
    rec, fm, _, _ = recon_util.make_reconstruction(n_cams=6, n_points=40, track_len=4, channels=16, seed=4)
    ids = sorted(rec.images)
    setup = ba.BundleAdjustmentSetup(); setup.add_images(ids[:4]); setup.set_constant_pose(ids[0]); setup.set_constant_tvec(ids[1], [0])
    for pid in sorted(rec.points3D)[:10]:
        setup.add_variable_point(pid)
    options = ba.BundleOptimizerOptions(refine_extra_params=False, min_track_length=2)
    prob, ir = ba.build_problem(rec, FeatureView(fm.fset(0), rec), setup, options)
    out = pb.build_problem(rec.as_arrays(),
                           {"image_ids": ids[:4], "const_pose_ids": [ids[0]], "const_tvec_ids": [ids[1]], "const_tvec_masks": np.array([1], np.uint8),
                            "var_point_ids": sorted(rec.points3D)[:10]},
                           {"refine_extra_params": 0, "min_track_length": 2})
    assert list(out["image_ids"]) == ir.image_ids and list(out["point_ids"]) == ir.point_ids and list(out["camera_ids"]) == ir.camera_ids
    for name in ("obs_img", "obs_pt", "img_cam", "pose_const", "tvec_const_mask", "point_const", "cam_const_mask"):
        assert np.array_equal(out[name], getattr(prob, name)), name
    assert [tuple(t) for t in zip(out["obs_image_id"], out["obs_point2D_idx"], out["obs_point3D_id"])] == list(ir.obs)
    with pytest.raises(ValueError, match="setup image"):
        pb.build_problem(rec.as_arrays(), {"image_ids": [999]}, {})
