"""Host-side steps right before the hot path: COLMAP camera projection in numpy (util/cameras.py, against the oracle's
projection for all seven models), patch placement and cutting from dense feature maps (features/extractor.py, against
plain loops), and the PixSfM driver fed by a DenseFeatureExtractor."""
import numpy as np
import pytest

import oracle_lib as O
from pixsfm import features
from pixsfm._pixsfm import _capi
from pixsfm.features.extractor import DenseFeatureExtractor, cut_patches, dense_to_fmap, patch_corners
from pixsfm.util import cameras, synthetic

MODELS = {0: [1200.0, 500.0, 480.0], 1: [1200.0, 1150.0, 500.0, 480.0], 2: [1200.0, 500.0, 480.0, 0.08],
          3: [1200.0, 500.0, 480.0, 0.08, -0.02], 4: [1200.0, 1150.0, 500.0, 480.0, 0.08, -0.02, 1e-3, -2e-3],
          5: [1200.0, 1150.0, 500.0, 480.0, 0.05, -0.01, 2e-3, -1e-3],
          6: [1200.0, 1150.0, 500.0, 480.0, 0.08, -0.02, 1e-3, -2e-3, 5e-3, 0.02, -0.01, 3e-3]}


@pytest.mark.parametrize("model", sorted(MODELS))
def test_world_to_image_matches_the_oracle_projection(model):
    prob, _ = synthetic.make_ba_scene(n_cams=3, n_points=25, track_len=3, channels=8, ps=8, seed=model, shared_camera=True)
    n_cam = len(prob.cam_model)
    prob.cam_model[:] = model
    prob.cam_params[:] = 0.0
    prob.cam_params[:, :len(MODELS[model])] = MODELS[model]
    prob.refs = np.zeros((len(prob.xyz), 8))
    xy = O.ba_evaluate(prob, _capi.default_interp(), _capi.default_ba_options())["xy"]
    for i in range(len(prob.qvec)):
        m = prob.obs_img == i
        got = cameras.world_to_image(model, prob.cam_params[prob.img_cam[i]], prob.qvec[i], prob.tvec[i], prob.xyz[prob.obs_pt[m]])
        assert np.abs(got - xy[m]).max() < 1e-9
    assert n_cam >= 1
    with pytest.raises(ValueError):
        cameras.normalized_to_image(model, MODELS[model][:-1], np.zeros((1, 2)))
    names = {v: k for k, v in _capi.CAMERA_MODEL_IDS.items()}
    assert np.array_equal(cameras.normalized_to_image(names[model], MODELS[model], [[0.1, -0.2]]),
                          cameras.normalized_to_image(model, MODELS[model], [[0.1, -0.2]]))


def test_fisheye_at_the_optical_axis():
    xy = cameras.normalized_to_image(5, MODELS[5], [[0.0, 0.0], [1e-300, 0.0]])
    assert np.allclose(xy, [[500.0, 480.0]] * 2)


def test_patch_corners_and_cutting_against_loops():
    rng = np.random.default_rng(0)
    H, W, C, ps = 40, 60, 5, 8
    dense = rng.normal(size=(H, W, C)).astype(np.float32)
    kps = np.concatenate([rng.uniform(-5, 130, (200, 2)), [[0.0, 0.0], [119.9, 79.9], [3.9, 3.9], [8.0, 8.0]]])
    scale = np.array([W / 120.0, H / 80.0])
    corners = patch_corners(kps, scale, ps, (W, H))
    for k, kp in enumerate(kps):
        want = [int(kp[0] * scale[0] - ps / 2.0), int(kp[1] * scale[1] - ps / 2.0)]     # int(): towards zero, like astype
        want = [min(max(want[0], 0), W - ps - 1), min(max(want[1], 0), H - ps - 1)]
        assert list(corners[k]) == want
    assert corners.dtype == np.int32 and corners.min() >= 0
    patches = cut_patches(dense, corners, ps)
    assert patches.shape == (len(kps), ps, ps, C) and patches.flags["C_CONTIGUOUS"]
    for k in (0, 17, len(kps) - 1):
        x0, y0 = corners[k]
        assert np.array_equal(patches[k], dense[y0:y0 + ps, x0:x0 + ps])


def test_dense_to_fmap_sparse_dense_and_normalisation():
    rng = np.random.default_rng(1)
    chw = rng.normal(size=(16, 50, 70)).astype(np.float32)
    kps = rng.uniform(20, 250, (30, 2))
    fm = dense_to_fmap(chw, (280, 200), kps, list(range(100, 130)), patch_size=8)
    assert fm.is_sparse and fm.size() == 30 and fm.shape == (8, 8, 16) and fm.dtype == np.float16
    assert np.allclose(fm.scale, [70 / 280, 50 / 200]) and fm.point2D_ids[:2] == [100, 101] and fm.has_point2D(129)
    norms = np.linalg.norm(fm.patches.astype(np.float32), axis=-1)
    assert np.abs(norms - 1).max() < 2e-3                                   # unit descriptors, fp16 rounding
    k = 7
    x0, y0 = fm.corners[k]
    want = np.moveaxis(chw, 0, -1)[y0:y0 + 8, x0:x0 + 8]
    want = (want / np.linalg.norm(want, axis=-1, keepdims=True)).astype(np.float16)
    assert np.array_equal(fm.patches[k], want)
    # the patches would be bigger than the map: it stays dense (one patch with the dense id, corner 0)
    many = rng.uniform(20, 250, (60, 2))
    dense = dense_to_fmap(chw, (280, 200), many, patch_size=8, l2_normalize=False, dtype=np.float32)
    assert not dense.is_sparse and dense.size() == 1 and dense.point2D_ids == [features.kDenseId]
    assert np.array_equal(dense.patches[0], np.moveaxis(chw, 0, -1)) and dense.has_point2D(12345)
    assert np.array_equal(dense_to_fmap(np.moveaxis(chw, 0, -1), (280, 200), many, patch_size=8, l2_normalize=False,
                                        dtype=np.float32, channels_first=False).patches, dense.patches)
    with pytest.raises(RuntimeError, match="without any keypoints"):
        dense_to_fmap(chw, (280, 200))
    with pytest.raises(ValueError, match="do not match"):
        dense_to_fmap(chw, (280, 200), kps, [1, 2, 3])
    assert not dense_to_fmap(chw, (280, 200), sparse=False).is_sparse


def test_extractor_places_patches_at_projections_and_feeds_the_driver(monkeypatch):
    """PixSfM(conf, extractor=...) without a ready feature manager: patches are cut around the projected 3D points,
    and the adjusters run on them (the oracle stands in for the device, as in test_mirror_with_oracle_solver)."""
    import test_mirror_with_oracle_solver as H
    from pixsfm._pixsfm import _engine
    from pixsfm.refine_colmap import PixSfM
    from recon_util import make_reconstruction
    monkeypatch.setattr(_engine, "ba_run", H._ba_run)
    monkeypatch.setattr(_engine, "refs_compute", H._refs_compute)
    rec = make_reconstruction(n_cams=4, n_points=30, track_len=3, channels=16, seed=9)[0]
    rng = np.random.default_rng(3)
    maps = {im.name: [rng.normal(size=(16, 250, 250)).astype(np.float32), rng.normal(size=(16, 125, 125)).astype(np.float32)]
            for im in rec.images.values()}
    ex = DenseFeatureExtractor(lambda name: maps[name], lambda name: (1000, 1000), {"patch_size": 8})
    fm = ex.features_from_reconstruction(rec, "unused")
    assert fm.num_levels == 2 and fm.fset(0).channels == 16
    image = next(iter(rec.images.values()))
    fmap = fm.fset(0).fmap(image.name)
    seen = [k for k, p in enumerate(image.points2D) if p.has_point3D()]
    assert fmap.is_sparse and sorted(fmap.point2D_ids) == seen
    cam = rec.cameras[image.camera_id]
    k = seen[0]
    proj = cameras.world_to_image(cam.model_id, cam.params, image.qvec, image.tvec, [rec.points3D[image.points2D[k].point3D_id].xyz])[0]
    assert list(fmap.corners[fmap.local_index(k)]) == list(patch_corners([proj], fmap.scale, 8, (250, 250))[0])
    out_rec, ba_data, fm2 = PixSfM({"BA": {"optimizer": {"solver": {"max_num_iterations": 3}}}}, extractor=ex).run_ba(rec, "unused")
    assert len(ba_data["summary"]) == 2 and fm2.num_levels == 2
    assert all(s.final_cost <= s.initial_cost for s in ba_data["summary"])
    with pytest.raises(ValueError, match="unknown extractor options"):
        DenseFeatureExtractor(None, None, {"patchsize": 8})


def test_extractor_feature_cache(tmp_path):
    """use_cache / overwrite_cache / load_cache_on_init as in the reference (features/extractor.py:46-49, extract.py:72-147):
    a fresh extraction goes to the cache file and comes back through it; an existing file is taken as it is."""
    from pixsfm._pixsfm._features import LazyFeatureMap
    from recon_util import make_reconstruction
    rec = make_reconstruction(n_cams=3, n_points=20, track_len=3, channels=8, seed=11)[0]
    rng = np.random.default_rng(4)
    maps = {im.name: [rng.normal(size=(8, 250, 250)).astype(np.float32)] for im in rec.images.values()}
    calls = []

    def model(name):
        calls.append(name)
        return maps[name]

    plain = DenseFeatureExtractor(model, lambda name: (1000, 1000), {"patch_size": 8}).features_from_reconstruction(rec, "unused")
    n_plain = len(calls)
    ex = DenseFeatureExtractor(model, lambda name: (1000, 1000), {"patch_size": 8, "use_cache": True})
    with pytest.raises(RuntimeError, match="no path given"):
        ex.features_from_reconstruction(rec, "unused")
    cache = tmp_path / "features.h5"
    fm = ex.features_from_reconstruction(rec, "unused", cache_path=cache)
    assert cache.exists() and len(calls) == 2 * n_plain
    for name in plain.fset(0).keys():
        a, b = plain.fset(0).fmap(name), fm.fset(0).fmap(name)
        assert isinstance(b, LazyFeatureMap) and not b.is_loaded          # load_cache_on_init off: metadata only
        assert a.point2D_ids == b.point2D_ids and np.array_equal(a.corners, b.corners) and np.array_equal(a.patches, b.patches)
    again = ex.features_from_reconstruction(rec, "unused", cache_path=cache)
    assert len(calls) == 2 * n_plain and sorted(again.fset(0).keys()) == sorted(plain.fset(0).keys())      # read, not extracted
    filled = DenseFeatureExtractor(model, lambda name: (1000, 1000), {"patch_size": 8, "use_cache": True, "load_cache_on_init": True,
                                                                     "overwrite_cache": True})
    fm3 = filled.features_from_reconstruction(rec, "unused", cache_path=cache)
    assert len(calls) == 3 * n_plain and not isinstance(next(iter(fm3.fset(0)._maps.values())), LazyFeatureMap)


def test_low_memory_preset_configures_the_extractor_cache(tmp_path, monkeypatch):
    """configs/low_memory.yaml: dense_features.use_cache / overwrite_cache / load_cache_on_init=false, patch_size 8 — the block
    reaches the extractor, `refine_reconstruction` puts the cache next to its output (refine_colmap.py:120-145) and the
    adjustment runs on the lazily filled manager."""
    import test_mirror_with_oracle_solver as H
    from pixsfm._pixsfm import _engine
    from pixsfm._pixsfm._features import LazyFeatureMap
    from pixsfm.refine_colmap import PixSfM
    from recon_util import make_reconstruction
    monkeypatch.setattr(_engine, "ba_run", H._ba_run)
    monkeypatch.setattr(_engine, "refs_compute", H._refs_compute)
    rec = make_reconstruction(n_cams=4, n_points=30, track_len=3, channels=16, seed=12)[0]
    rng = np.random.default_rng(6)
    maps = {im.name: [rng.normal(size=(16, 250, 250)).astype(np.float32)] for im in rec.images.values()}
    ex = DenseFeatureExtractor(lambda name: maps[name], lambda name: (1000, 1000))
    conf = {"dense_features": dict(sparse=True, dtype="half", use_cache=True, overwrite_cache=True, load_cache_on_init=False, patch_size=8),
            "BA": {"optimizer": {"solver": {"max_num_iterations": 2}}}}
    sfm = PixSfM(conf, extractor=ex)
    assert ex.conf["patch_size"] == 8 and ex.conf["use_cache"] and not ex.conf["load_cache_on_init"]
    (tmp_path / "in").mkdir()
    rec.write(str(tmp_path / "in"))
    out_rec, ba_data, fm = sfm.refine_reconstruction(tmp_path / "out", tmp_path / "in", "unused")
    assert (tmp_path / "out" / "features_featuremaps_sparse.h5").exists()
    maps0 = list(fm.fset(0)._maps.values())
    assert all(isinstance(m, LazyFeatureMap) for m in maps0) and maps0[0].shape == (8, 8, 16)
    assert not any(m.is_loaded for m in maps0)                   # the views of the adjustment have let go of the patches
    assert ba_data["summary"][0].final_cost <= ba_data["summary"][0].initial_cost


def test_extract_module_functions():
    """pixsfm.extract: the reference's free functions (extract.py:22-222) over the extractor object"""
    from pixsfm import extract
    from recon_util import make_reconstruction
    rec = make_reconstruction(n_cams=3, n_points=20, track_len=3, channels=8, seed=13)[0]
    rng = np.random.default_rng(5)
    maps = {im.name: [rng.normal(size=(8, 250, 250)).astype(np.float32)] for im in rec.images.values()}
    ex = DenseFeatureExtractor(lambda name: maps[name], lambda name: (1000, 1000), {"patch_size": 8})
    a = extract.features_from_reconstruction(ex, rec, "unused")
    b = ex.features_from_reconstruction(rec, "unused")
    for name in b.fset(0).keys():
        assert a.fset(0).fmap(name).point2D_ids == b.fset(0).fmap(name).point2D_ids
        assert np.array_equal(a.fset(0).fmap(name).patches, b.fset(0).fmap(name).patches)
    kps = {"x": np.arange(12.0).reshape(6, 2)}
    k, ids, n = extract.get_keypoints_and_ids("x", kps, {"x": [4, 1]})
    assert n == 2 and ids == [4, 1] and np.array_equal(k, kps["x"][[4, 1]])
    assert extract.get_keypoints_and_ids("x", None, None) == (None, None, 0)
    assert extract.estimate_required_memory(ex, None, ["x"], kps, {"x": [4, 1]}, use_cache=True) == 0
