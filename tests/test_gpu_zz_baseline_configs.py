"""Parity at the SHAPES of BASELINE.json's configs[0] and configs[1] (the sacre_coeur cases; synthetic features, since
the S2DNet weights are not available offline):
  configs[0]  keypoint adjustment, 10 images / ~8 000 matched keypoints, 16-channel single-level features
  configs[1]  featuremetric BA, 10 cameras / 1 856 points / 8 076 observations (ragged tracks), 128 channels -> the
              DENSE_SCHUR regime of bundle_optimizer.h:184-185
Both against the oracle solving the same IR.  (configs[2] is tests/test_gpu_full_size.py + bench.py, configs[3] the
triangulation case of tests/test_gpu_edge_cases.py and the cost-map tests.)"""
import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic
from ka_util import make_ka_problem

pytestmark = pytest.mark.gpu


def test_config0_shape_keypoint_adjustment_16_channels():
    prob, _, lab = make_ka_problem(n_images=10, n_tracks=2000, track_len=4, channels=16, seed=40, kp_sigma=0.8,
                                    bound=4.0, max_per_problem=50)
    assert len(prob.keypoints) == 8000 and prob.n_problems >= 150
    ic = _capi.default_interp(); so = _capi.default_ka_options()
    p_cpu, p_gpu = prob.copy(), prob.copy()
    c0, c1 = O.ka_solve(p_cpu, ic, so)
    s = _engine.ka_run(p_gpu, ic, so)
    assert abs(s["initial_cost"] - c0) <= 1e-9 * c0
    assert abs(s["final_cost"] - c1) <= 1e-5 * c1 and c1 < 0.7 * c0
    d = np.abs(p_gpu.keypoints - p_cpu.keypoints).max(axis=1)
    assert np.quantile(d, 0.999) < 1e-5 and d.max() < 1e-3          # one packed problem = one trust region: allow a rare late step
    roots = lab["roots"].astype(bool)
    assert np.array_equal(p_gpu.keypoints[roots], prob.keypoints[roots])


def _ragged(prob, n_keep, rng):
    """drop observations at random (every point keeps at least two) until n_keep are left"""
    n_pts = len(prob.xyz)
    keep = np.ones(prob.n_obs, bool)
    per_point = np.bincount(prob.obs_pt, minlength=n_pts)
    for o in rng.permutation(prob.n_obs):
        if keep.sum() == n_keep:
            break
        if per_point[prob.obs_pt[o]] > 2:
            keep[o] = False
            per_point[prob.obs_pt[o]] -= 1
    idx = np.where(keep)[0]
    return _capi.BAProblem(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask,
                           qvec=prob.qvec, tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const,
                           tvec_const_mask=prob.tvec_const_mask, xyz=prob.xyz, point_const=prob.point_const,
                           obs_img=prob.obs_img[idx], obs_pt=prob.obs_pt[idx], patches=np.ascontiguousarray(prob.patches[idx]),
                           corner=prob.corner[idx], scale=prob.scale[idx])


def test_config1_shape_bundle_adjustment_dense_schur():
    full, gt = synthetic.make_ba_scene(n_cams=10, n_points=1856, track_len=5, channels=128, seed=41)
    prob = _ragged(full, 8076, np.random.default_rng(41))
    assert prob.n_obs == 8076 and len(prob.xyz) == 1856
    lens = np.bincount(prob.obs_pt)
    assert lens.min() >= 2 and lens.max() == 5 and len(set(lens)) > 2
    ic = _capi.default_interp()
    refs_gpu, src_gpu = _engine.refs_compute(prob, ic)
    refs_cpu, src_cpu = O.refs_compute(prob, ic)
    ties = lens == 2                                        # two observations are equally far from their mean
    assert np.array_equal(src_gpu[~ties], src_cpu[~ties]) and np.abs(refs_gpu[~ties] - refs_cpu[~ties]).max() < 1e-12
    prob.refs = refs_cpu
    so = _capi.default_ba_options(max_num_iterations=12)    # defaults: Cauchy(0.25), inner iterations, solver by #images
    p_cpu, p_gpu = prob.copy(), prob.copy()
    s_cpu = O.ba_solve(p_cpu, ic, so)
    s_gpu = _engine.ba_run(p_gpu, ic, so)
    assert abs(s_gpu["initial_cost"] - s_cpu["initial_cost"]) <= 1e-10 * s_cpu["initial_cost"]
    for ig, ir in list(zip(s_gpu["iterations"], s_cpu["iterations"]))[:4]:
        assert ig["step_is_successful"] == ir["step_is_successful"]
        assert abs(ig["cost"] - ir["cost"]) <= 1e-6 * abs(ir["cost"])
    assert s_cpu["final_cost"] < s_cpu["initial_cost"]
    assert abs(s_gpu["final_cost"] - s_cpu["final_cost"]) <= 1e-5 * s_cpu["final_cost"]
    # the reference's own tolerance for comparing two BA results is 1e-4 (bundle_optimizer_test.cc:52)
    assert np.abs(p_gpu.qvec - p_cpu.qvec).max() < 1e-4 and np.abs(p_gpu.tvec - p_cpu.tvec).max() < 1e-4
    assert np.abs(p_gpu.xyz - p_cpu.xyz).max() < 1e-4
    assert np.abs(p_gpu.cam_params[:, 0] / p_cpu.cam_params[:, 0] - 1).max() < 1e-4


def test_device_patch_gather_matches_the_host_extraction():
    """pxr_extract_patches (dense CNN map -> patch slab on the device) against the numpy step it replaces
    (features/extractor.py: dense_to_fmap), host and device-resident inputs, both memory layouts."""
    from pixsfm.features.extractor import cut_patches, dense_to_fmap, dense_to_fmap_on_device, patch_corners
    rng = np.random.default_rng(7)
    C, H, W, ps = 128, 96, 96, 16
    chw = rng.normal(size=(C, H, W)).astype(np.float32)
    # 30 patches of 16 x 16 px are smaller than the 96 x 96 map: the reference's "sparse pays off" rule keeps the sparse
    # branch (features/extractor.py:182-187); out-of-image keypoints exercise the corner clamp (:192-193)
    n_kp = 30
    assert n_kp * ps * ps < H * W
    kps = rng.uniform(-10, 400, (n_kp, 2))
    corners = patch_corners(kps, np.array([W / 384.0, H / 384.0]), ps, (W, H))
    fm_want = dense_to_fmap(chw, (384, 384), kps, patch_size=ps)
    assert fm_want.is_sparse
    want = fm_want.patches                                                         # fp16, L2-normalised
    got = _engine.extract_patches(chw, corners, ps, l2_normalize=True, out_dtype=np.float16, to_host=True)
    assert got.shape == want.shape == (n_kp, ps, ps, C) and got.dtype == np.float16
    diff = np.abs(got.astype(np.float32) - want.astype(np.float32))
    assert diff.max() <= 2.0 ** -11 and np.mean(got == want) > 0.99              # <= 1 fp16 ulp at |v| < 1, almost all equal
    # no normalisation, same dtype: a pure gather, bit for bit; [H,W,C] layout too
    hwc = np.ascontiguousarray(np.moveaxis(chw, 0, -1))
    exact = cut_patches(hwc, corners, ps)
    assert np.array_equal(_engine.extract_patches(chw, corners, ps, False, np.float32, True, to_host=True), exact)
    assert np.array_equal(_engine.extract_patches(hwc, corners, ps, False, np.float32, False, to_host=True), exact)
    assert np.array_equal(_engine.extract_patches(hwc.astype(np.float64), corners, ps, False, np.float64, False, to_host=True), exact)
    # device-resident input: the whole map as one 96 x 96 "patch" stays on the device and is gathered from there
    whole = _engine.extract_patches(hwc, np.zeros((1, 2), np.int32), H, False, np.float32, False)
    assert whole.shape == (1, H, W, C)
    assert np.array_equal(_engine.extract_patches(whole, corners, ps, False, np.float32, False, to_host=True), exact)
    with pytest.raises(ValueError):
        _engine.extract_patches(chw, np.array([[W - ps + 1, 0]], np.int32), ps, to_host=True)      # leaves the map
    assert _engine.extract_patches(chw, np.zeros((0, 2), np.int32), ps, to_host=True).shape == (0, ps, ps, C)
    # FeatureMap built on the device: same metadata as the host one, patches device-resident
    ids = list(range(1000, 1000 + n_kp))
    fm_dev = dense_to_fmap_on_device(chw, (384, 384), kps, ids, patch_size=ps)
    fm_host = dense_to_fmap(chw, (384, 384), kps, ids, patch_size=ps)
    assert fm_dev.is_sparse and fm_host.is_sparse
    assert fm_dev.point2D_ids == fm_host.point2D_ids and np.array_equal(fm_dev.corners, fm_host.corners)
    assert np.array_equal(fm_dev.scale, fm_host.scale) and fm_dev.patches.shape == fm_host.patches.shape
    # semi-dense keypoints: sparse does not pay off -> both paths keep ONE dense patch (kDenseId), as tensor_to_fmap does
    many = rng.uniform(0, 384, (300, 2))
    fd, fh = dense_to_fmap_on_device(chw, (384, 384), many, patch_size=ps), dense_to_fmap(chw, (384, 384), many, patch_size=ps)
    assert not fd.is_sparse and not fh.is_sparse and fd.patches.shape == fh.patches.shape == (1, H, W, C)
    assert fd.point2D_ids == fh.point2D_ids


def test_patch_interpolator_on_the_device_matches_the_oracle():
    """features.PatchInterpolator (features/bindings.cc:276-292) through pxr_interpolate_descriptors, fp16 / fp32 / fp64
    patches, image and local coordinates"""
    from pixsfm import features
    rng = np.random.default_rng(12)
    for dtype in (np.float16, np.float32, np.float64):
        data = rng.normal(size=(16, 16, 128)).astype(dtype)
        patch = features.FeaturePatch(data, (100, 40), (0.5, 0.25))
        for l2 in (True, False):
            pi = features.PatchInterpolator({"l2_normalize": l2})
            for xy in ([213.3, 188.9], [205.0, 170.0], [228.7, 219.1]):
                uv = patch.to_pixel_coordinates(np.array(xy))
                want = O.pixel_interp(data, uv[1], uv[0], l2)[0]
                assert np.abs(pi.interpolate(patch, xy)[0] - want).max() < 1e-12
                assert np.abs(pi.interpolate_local(patch, uv)[0] - want).max() < 1e-11
                assert pi.interpolate_nodes(patch, xy).shape == (1, 128)


def _multilevel_scene(n_cams, n_points, track_len, channels, level_scales, seed):
    """configs[3]-shaped inputs: one reconstruction and a FeatureManager with one FeatureSet per level.  Level l holds, per
    observation, a patch of the point's feature field rendered at resolution `level_scales[l]` of the image (S2DNet's
    levels sit at 1, 1/4, 1/16 of the image, reference models/s2dnet.py:65,92-98), the way features_from_reconstruction
    delivers them: corner and scale in that level's pixel grid (features/extractor.py:192-193)."""
    from pixsfm import features
    from recon_util import make_reconstruction
    rec, fm0, prob, gt = make_reconstruction(n_cams=n_cams, n_points=n_points, track_len=track_len, channels=channels, seed=seed)
    fm = features.FeatureManager([channels] * len(level_scales), prob.patches.dtype)
    per_image = {i: [o for o in range(prob.n_obs) if int(prob.obs_img[o]) == i] for i in range(n_cams)}
    for lvl, sc in enumerate(level_scales):
        size = int(round(1000 * sc))
        patches, corners, _ = synthetic.render_patches(gt["xy_true"] * sc, prob.obs_pt, n_points, channels, 16, seed + 17 * lvl,
                                                       0.01, prob.patches.dtype, image_size=max(size, 40))
        fset = features.FeatureSet(channels, prob.patches.dtype)
        for i in range(n_cams):
            obs = per_image[i]
            fset.emplace("image%03d.jpg" % i, features.FeatureMap(np.ascontiguousarray(patches[obs]), list(range(len(obs))),
                                                                  corners[obs], {"scale": (sc, sc), "is_sparse": True}))
        fm.fsets[lvl] = fset
    return rec, fm


def test_config3_shape_multilevel_costmap_triangulation_ba():
    """BASELINE configs[3]: ETH3D-courtyard-shaped triangulation refinement — 38 images (the courtyard scene's size; the
    reference tree only names the scene, eval/eth3d/config.py:7-8), multi-level features processed coarse to fine,
    strategy `costmaps`, poses and intrinsics fixed, 10 iterations per level (configs/pixsfm_eth3d.yaml:
    refine_focal_length / refine_extra_params / refine_extrinsics: false; bundle_adjustment/main.py:218-286).
    The mirror's three refinements against the oracle doing references -> cost maps -> cost-map BA level by level."""
    import copy
    from pixsfm import bundle_adjustment as ba_pkg, features
    from pixsfm._pixsfm import _bundle_adjustment as ba
    scales = (1.0, 0.25, 0.0625)
    rec, fm = _multilevel_scene(n_cams=38, n_points=700, track_len=5, channels=128, level_scales=scales, seed=43)
    rec_ref = copy.deepcopy(rec)
    conf = {"strategy": "costmaps",
            "optimizer": {"refine_focal_length": False, "refine_extra_params": False, "refine_extrinsics": False,
                          "solver": {"max_num_iterations": 10}}}
    adj = ba_pkg.BundleAdjuster.create(conf)
    out = adj.refine_multilevel(rec, fm)
    assert len(out["summary"]) == 3 and len(out["costmaps"]) == 3
    # nothing but the points may move
    for i in rec.images:
        assert np.array_equal(rec.images[i].tvec, rec_ref.images[i].tvec)
        assert np.abs(rec.images[i].qvec - rec_ref.images[i].qvec / np.linalg.norm(rec_ref.images[i].qvec)).max() < 1e-15
    for c in rec.cameras:
        assert np.array_equal(rec.cameras[c].params, rec_ref.cameras[c].params)
    # the oracle, level by level in the adjuster's order (coarse to fine: reverse index order, util/misc.py:19-23)
    setup = ba_pkg.default_problem_setup(rec_ref)
    options = ba.BundleOptimizerOptions(refine_focal_length=False, refine_extra_params=False, refine_extrinsics=False)
    ic = _capi.default_interp()
    ic_cm = _capi.default_interp(); ic_cm.l2_normalize = 0
    so = _capi.default_ba_options(max_num_iterations=10)     # the adjuster's defaults: inner iterations on, Cauchy(0.25)
    for k, lvl in enumerate((2, 1, 0)):
        fview = features.FeatureView(fm.fset(lvl), rec_ref)
        prob_r, ir_r = ba.build_problem(rec_ref, fview, None, None, None, for_references=set(rec_ref.points3D.keys()))
        prob_r.refs = O.refs_compute(prob_r, ic, iters=100)[0]
        cm = O.costmaps_compute(prob_r)
        prob, ir = ba.build_problem(rec_ref, fview, setup, options, None)
        assert prob.n_obs == prob_r.n_obs and np.all(prob.pose_const == 1) and np.all(prob.cam_const_mask & 0xF == 0xF)
        p = prob.with_patches(cm)
        s = O.ba_solve(p, ic_cm, so)
        summ = out["summary"][k]
        assert abs(summ.initial_cost - s["initial_cost"]) <= 1e-5 * max(s["initial_cost"], 1e-12)
        assert abs(summ.final_cost - s["final_cost"]) <= 1e-4 * max(s["final_cost"], 1e-12) + 1e-9
        ba.write_back(rec_ref, p, ir)
    moved = [np.abs(rec.points3D[q].xyz - rec_ref.points3D[q].xyz).max() for q in rec.points3D]
    assert max(moved) < 1e-5
