"""pxr_problem_build (csrc/pxr_problem.cu: BundleOptimizer::SetUp + Parameterize in C++, reference
bundle_adjustment/src/bundle_optimizer.h:139-165,247-442; GetVisibleObservations, reference_extractor.h:171-205) against
the per-object Python restatement `build_problem_py` — every array of the problem IR must be identical — on the
parameterisation cases of the reference's bundle_optimizer_test.cc (as tests/test_mirror_setup.py restates them) and on
random setups; plus the array-backed reconstruction."""
import copy

import numpy as np
import pytest

import recon_util
from pixsfm._pixsfm import _bundle_adjustment as ba
from pixsfm._pixsfm._features import FeatureView
from pixsfm.util.colmap_types import ArrayReconstruction


def _same(a, b):
    pa, ia = a
    pb, ib = b
    assert ia.image_ids == ib.image_ids and ia.camera_ids == ib.camera_ids and ia.point_ids == ib.point_ids
    assert list(ia.obs) == list(ib.obs)
    if not list(ia.obs):
        assert pa.n_obs == pb.n_obs == 0
        return
    for name in ("cam_model", "cam_params", "cam_const_mask", "qvec", "tvec", "img_cam", "pose_const", "tvec_const_mask", "xyz",
                 "point_const", "obs_img", "obs_pt", "obs_patch", "corner", "scale"):
        x, y = getattr(pa, name), getattr(pb, name)
        assert x.dtype == y.dtype and np.array_equal(x, y), name
    assert len(pa.patch_blocks) == len(pb.patch_blocks) and all(x is y for x, y in zip(pa.patch_blocks, pb.patch_blocks))
    assert ia.slab_offsets == ib.slab_offsets


def _scene(seed=0, **kw):
    args = dict(n_cams=6, n_points=40, track_len=4, channels=16, seed=seed)
    args.update(kw)
    rec, fm, _, _ = recon_util.make_reconstruction(**args)
    return rec, fm.fset(0)


def _both(rec, fview, setup, options, **kw):
    r1, r2 = copy.deepcopy(rec), copy.deepcopy(rec)
    return ba.build_problem(r1, fview, setup, options, **kw), ba.build_problem_py(r2, fview, setup, options, **kw)


@pytest.mark.parametrize("case", range(8))
def test_cpp_builder_equals_the_python_restatement(case):
    rec, fset = _scene(seed=case)
    fview = FeatureView(fset, rec)
    rng = np.random.default_rng(case)
    setup = ba.BundleAdjustmentSetup()
    image_ids = sorted(rec.images)
    chosen = image_ids if case % 2 == 0 else image_ids[: len(image_ids) // 2 + 1]
    setup.add_images(chosen)
    if case in (1, 3, 5):      # points seen from images outside the setup (AddPointToProblem)
        pts = sorted(rec.points3D)
        for pid in pts[: len(pts) // 2]:
            setup.add_variable_point(pid)
        for pid in pts[len(pts) // 2: len(pts) // 2 + 5]:
            setup.add_constant_point(pid)
    if case >= 2:
        setup.set_constant_pose(chosen[0])
        setup.set_constant_tvec(chosen[1], [int(rng.integers(0, 3))])
    if case in (4, 6):
        setup.set_constant_camera(rec.images[chosen[0]].camera_id)
    options = ba.BundleOptimizerOptions(refine_focal_length=case % 3 != 0, refine_extra_params=case % 2 == 0,
                                        refine_principal_point=case == 5, refine_extrinsics=case != 7,
                                        min_track_length=[-1, 2, 3, 5][case % 4])
    _same(*_both(rec, fview, setup, options))


def test_reference_observation_lists_and_missing_patches():
    rec, fset = _scene(seed=3)
    fview = FeatureView(fset, rec)
    ids = set(sorted(rec.points3D)[::2])
    _same(*_both(rec, fview, None, None, references=None, for_references=ids))
    # drop a feature map: its observations disappear from the list (a warning, reference_extractor.h:187-191)
    name = rec.images[sorted(rec.images)[1]].name
    del fset._maps[name]
    a, b = _both(rec, fview, None, None, references=None, for_references=ids)
    _same(a, b)
    assert all(o[0] != sorted(rec.images)[1] for o in a[1].obs)


def test_errors_travel_as_value_errors():
    rec, fset = _scene(seed=1)
    fview = FeatureView(fset, rec)
    setup = ba.BundleAdjustmentSetup(); setup.add_images(sorted(rec.images))
    pid = sorted(rec.points3D)[0]
    rec.points3D[pid].track.elements.pop(0)          # a 2D point now claims a track element that is gone
    with pytest.raises(ValueError, match="Failed to register track element"):
        ba.build_problem(rec, fview, setup, ba.BundleOptimizerOptions())


def test_array_reconstruction_builds_the_same_problem_and_takes_the_write_back():
    rec, fset = _scene(seed=5, n_points=60)
    fview = FeatureView(fset, rec)
    A = rec.as_arrays()
    image_ids = sorted(rec.images); point_ids = sorted(rec.points3D); camera_ids = sorted(rec.cameras)
    arec = ArrayReconstruction(image_ids, [rec.images[i].name for i in image_ids], A["image_camera_id"],
                               [rec.images[i].qvec for i in image_ids], [rec.images[i].tvec for i in image_ids],
                               A["p2d_begin"], A["p2d_point3D_id"], camera_ids, A["camera_model"],
                               [rec.cameras[c].params for c in camera_ids], point_ids, [rec.points3D[p].xyz for p in point_ids],
                               A["track_begin"], A["track_image_id"], A["track_point2D_idx"])
    setup = ba.BundleAdjustmentSetup(); setup.add_images(image_ids); setup.set_constant_pose(image_ids[0])
    options = ba.BundleOptimizerOptions()
    pa, ia = ba.build_problem(arec, FeatureView(fset, arec), setup, options)
    pb, ib = ba.build_problem_py(copy.deepcopy(rec), fview, setup, options)
    _same((pa, ia), (pb, ib))
    pa.xyz += 1.0; pa.qvec[:] = 0.5; pa.cam_params[:, 0] = 7.0
    ba.write_back(arec, pa, ia)
    assert np.allclose(arec.xyz[np.searchsorted(arec.point3D_id, ia.point_ids)], pa.xyz)
    assert np.allclose(arec.qvec, 0.5) and all(c[0] == 7.0 for c in arec.cam_params)
