"""Query refinement (localization) on the GPU against the oracle: descriptor interpolation, nearest references, query
keypoint adjustment and query bundle adjustment (reference pixsfm/localization/src/*.h, localization/main.py)."""
import copy

import numpy as np
import pytest

import oracle_lib as O
from pixsfm import bundle_adjustment as ba_pkg, features, localization as loc_pkg
from pixsfm._pixsfm import _bundle_adjustment as ba
from pixsfm._pixsfm import _capi, _localization as loc
from pixsfm.util import colmap_types as ct
from recon_util import make_reconstruction

pytestmark = pytest.mark.gpu


def _scene():
    rec, fm, prob, gt = make_reconstruction(n_cams=6, n_points=60, track_len=4, channels=128, seed=15)
    # references of the mapped model, with all observations kept (QueryLocalizer default keep_observations=True)
    labels = ba_pkg.find_problem_labels(rec, 10)
    refs = ba.ReferenceExtractor({"iters": 100, "keep_observations": True}, {}).run(labels, rec, fm.fset(0))
    # the "query" = image 3 of the scene: its keypoints (perturbed), its feature map, the 3D points it sees
    qid = 3
    img = rec.images[qid]
    fmap = fm.fset(0).fmap(img.name)
    p2D_idxs = [i for i, p in enumerate(img.points2D) if p.has_point3D()]
    p3D_ids = [img.points2D[i].point3D_id for i in p2D_idxs]
    kps = np.array([img.points2D[i].xy for i in p2D_idxs], np.float64)
    return rec, fm, refs, qid, fmap, p2D_idxs, p3D_ids, kps


def test_descriptor_interpolation_and_nearest_references():
    rec, fm, refs, qid, fmap, p2D_idxs, p3D_ids, kps = _scene()
    desc = loc.interpolate_descriptors(fmap, p2D_idxs, kps, {})
    for k in (0, 7, len(kps) - 1):
        li = fmap.local_index(p2D_idxs[k])
        uv = (kps[k] * fmap.scale - 0.5 - fmap.corners[li])
        want = O.pixel_interp(fmap.patches[li], uv[1], uv[0], l2_normalize=True)[0]
        assert np.abs(desc[k] - want).max() < 1e-12
    nearest = loc.find_nearest_references(fmap, refs, kps, p3D_ids, {}, p2D_idxs)
    assert len(nearest) == len(kps)
    for k, pid in enumerate(p3D_ids):      # first minimum of the squared distances, as nearest_references.h:38-47
        obs = np.array([o.reshape(-1) for o in refs[pid].observations])
        assert len(obs) == 4
        d = ((obs - desc[k]) ** 2).sum(1)
        assert np.array_equal(nearest[k].reshape(-1), obs[int(np.argmin(d))])


@pytest.mark.parametrize("target", ["nearest", "robust_mean", "all_observations"])
def test_query_keypoint_adjuster_matches_oracle(target):
    rec, fm, refs, qid, fmap, p2D_idxs, p3D_ids, kps = _scene()
    rng = np.random.default_rng(3)
    kps = kps + rng.normal(0, 1.0, kps.shape)
    if target == "nearest":
        # exclude the query's own observation, otherwise the target is the keypoint's current descriptor
        r = []
        q = loc.interpolate_descriptors(fmap, p2D_idxs, kps, {})
        for k, pid in enumerate(p3D_ids):
            obs = [o.reshape(-1) for o, el in zip(refs[pid].observations, rec.points3D[pid].track.elements) if el.image_id != qid]
            d = [((o - q[k]) ** 2).sum() for o in obs]
            r.append(obs[int(np.argmin(d))].reshape(1, -1))
    elif target == "robust_mean":
        r = [refs[p] for p in p3D_ids]
        for x in r:
            x.observations = []                       # Reference without kept observations -> its descriptor
    else:
        r = [[o for o in refs[p].observations] for p in p3D_ids]
    qka = loc_pkg.QueryKeypointAdjuster({"optimizer": {"bound": 4.0}})
    prob, used = qka.solver.build_problem(kps.copy(), fmap, r, patch_idxs=p2D_idxs)
    p_cpu = prob.copy()
    c0, c1 = O.ka_solve(p_cpu, _capi.default_interp(), qka.solver.solver_options())
    refined = kps.copy()
    qka.refine(refined, fmap, r, point2D_idxs=p2D_idxs)
    s = qka.solver.summary()
    assert abs(s.initial_cost - c0) <= 1e-9 * c0 and abs(s.final_cost - c1) <= 1e-6 * c1 and c1 < c0
    assert np.abs(refined[used] - p_cpu.keypoints).max() < 1e-5
    assert np.abs(refined - kps).max() > 1e-2


@pytest.mark.parametrize("refine_focal", [False, True])
def test_query_bundle_adjuster_matches_oracle(refine_focal):
    rec, fm, refs, qid, fmap, p2D_idxs, p3D_ids, kps = _scene()
    img = rec.images[qid]
    cam = copy.deepcopy(rec.cameras[img.camera_id])
    rng = np.random.default_rng(5)
    qvec = img.qvec.copy() + rng.normal(0, 2e-4, 4); qvec /= np.linalg.norm(qvec)
    tvec = img.tvec.copy() + rng.normal(0, 2e-3, 3)
    points3D = [rec.points3D[p].xyz.copy() for p in p3D_ids]
    r = [refs[p] for p in p3D_ids]
    for x in r:
        x.observations = []
    inliers = [k % 7 != 0 for k in range(len(p3D_ids))]
    qba = loc_pkg.QueryBundleAdjuster({"optimizer": {"refine_focal_length": refine_focal, "solver": {"max_num_iterations": 20}}})
    prob = qba.solver.build_problem(qvec.copy(), tvec.copy(), copy.deepcopy(cam), points3D, fmap, r, inliers, p2D_idxs)
    assert prob.n_obs == sum(inliers) and prob.point_const.all()
    p_cpu = prob.copy()
    s_o = O.ba_solve(p_cpu, _capi.default_interp(), qba.solver.solver_options())
    q2, t2, cam2 = qvec.copy(), tvec.copy(), copy.deepcopy(cam)
    assert qba.refine(q2, t2, cam2, points3D, fmap, r, inliers=inliers, point2D_idxs=p2D_idxs)
    s = qba.solver.summary()
    assert s.num_iterations == s_o["num_iterations"]
    assert abs(s.final_cost - s_o["final_cost"]) <= 1e-6 * s_o["final_cost"] and s_o["final_cost"] < s_o["initial_cost"]
    assert np.abs(q2 - p_cpu.qvec[0]).max() < 1e-6 and np.abs(t2 - p_cpu.tvec[0]).max() < 1e-6
    assert np.abs(q2 - qvec).max() > 1e-6
    if refine_focal:
        assert abs(cam2.params[0] / p_cpu.cam_params[0, 0] - 1) < 1e-6 and cam2.params[0] != cam.params[0]
    else:
        assert np.array_equal(cam2.params, cam.params)


def test_query_localizer_runs_qka_pnp_qba():
    rec, fm, refs, qid, fmap, p2D_idxs, p3D_ids, kps = _scene()
    img = rec.images[qid]
    cam = copy.deepcopy(rec.cameras[img.camera_id])
    rng = np.random.default_rng(9)
    start = {"success": True, "qvec": img.qvec.copy(), "tvec": img.tvec + rng.normal(0, 2e-3, 3), "inliers": [True] * len(kps)}
    seen = {}

    def pnp(points2D, points3D, camera):      # stands in for pycolmap.absolute_pose_estimation
        seen["kps"] = points2D.copy()
        return dict(start, qvec=start["qvec"].copy(), tvec=start["tvec"].copy())

    ql = loc_pkg.QueryLocalizer(rec, {"target_reference": "robust_mean"}, references=[refs], pose_estimator=pnp)
    noisy = kps + rng.normal(0, 0.7, kps.shape)
    all_kps = np.zeros((max(p2D_idxs) + 1, 2)); all_kps[list(p2D_idxs)] = noisy     # the query image's keypoint array
    out = ql.localize(all_kps, p2D_idxs, p3D_ids, cam, query_fmaps=[fmap])
    assert len(out["inliers"]) == len(kps) and out["num_inliers"] == sum(out["inliers"])
    assert out["success"] and np.abs(seen["kps"] - noisy).max() > 1e-2      # QKA moved the keypoints before PnP
    assert np.abs(out["tvec"] - start["tvec"]).max() > 1e-6                  # QBA refined the PnP pose
    # reprojection of the mapped points with the refined pose stays within a pixel of the refined keypoints
    from pixsfm.util import synthetic
    xy = synthetic.project_simple_radial(np.asarray(cam.params), out["qvec"], out["tvec"],
                                         np.array([rec.points3D[p].xyz for p in p3D_ids]))
    assert np.median(np.linalg.norm(xy - out["keypoints"], axis=1)) < 1.0
