"""Turns a synthetic BA scene (flat IR) into the reference-style objects: a COLMAP-like Reconstruction and a
FeatureManager with one FeatureMap per image, the way features_from_reconstruction would deliver them
(reference pixsfm/extract.py:153-194, features/extractor.py:175-201)."""
import numpy as np

from pixsfm import features
from pixsfm.util import colmap_types as ct
from pixsfm.util import synthetic


def make_reconstruction(n_cams=6, n_points=60, track_len=4, channels=128, seed=0, **kw):
    prob, gt = synthetic.make_ba_scene(n_cams=n_cams, n_points=n_points, track_len=track_len, channels=channels,
                                       seed=seed, **kw)
    rec = ct.Reconstruction()
    n_cam_models = len(prob.cam_model)
    for c in range(n_cam_models):
        rec.add_camera(ct.Camera(c + 1, "SIMPLE_RADIAL", 1000, 1000, prob.cam_params[c, :4]))
    per_image = {i: [] for i in range(n_cams)}   # obs indices per image, in point order
    for o in range(prob.n_obs):
        per_image[int(prob.obs_img[o])].append(o)
    for i in range(n_cams):
        img = ct.Image(i + 1, "image%03d.jpg" % i, int(prob.img_cam[i]) + 1, prob.qvec[i], prob.tvec[i],
                       [ct.Point2D(gt["xy_true"][o]) for o in per_image[i]])
        rec.add_image(img)
    for p in range(n_points):
        rec.add_point3D(p + 1, prob.xyz[p])
    fset = features.FeatureSet(channels, prob.patches.dtype)
    for i in range(n_cams):
        obs = per_image[i]
        for local, o in enumerate(obs):
            rec.add_observation(int(prob.obs_pt[o]) + 1, i + 1, local)
        fmap = features.FeatureMap(np.ascontiguousarray(prob.patches[obs]), list(range(len(obs))), prob.corner[obs],
                                   {"scale": prob.scale[obs[0]], "is_sparse": True})
        fset.emplace("image%03d.jpg" % i, fmap)
    fm = features.FeatureManager([channels], prob.patches.dtype)
    fm.fsets[0] = fset
    return rec, fm, prob, gt
