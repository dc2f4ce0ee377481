"""pixsfm.localization.main — the reference's query refinement surface (pixsfm/localization/main.py:20-330): helper
functions, QueryKeypointAdjuster, QueryBundleAdjuster and a QueryLocalizer whose PnP step is pluggable (the reference
calls pycolmap.absolute_pose_estimation, which is not part of this package)."""
import numpy as np

from .. import defaults
from .._pixsfm import _localization as loc
from ..util.conf import merge, to_ctr
from ..util.refine import level_order, optimizer_options

resolve_level_indices = level_order          # the reference's name (util/misc.py:19-23)
to_optim_ctr = optimizer_options             # the reference's name (util/misc.py:30-36)


def find_feature_inliers(p2Ds, fmap, references, interpolation_config, thresh=-1, point2D_idxs=None):
    """a correspondence whose reference is ONE descriptor (ndarray) is an outlier when the descriptor interpolated at
    its keypoint is farther than `thresh` (L2) from it; other reference kinds are not tested; thresh < 0 switches the
    test off (reference main.py:20-36)"""
    n = len(p2Ds)
    if thresh < 0.0:
        return [True] * n
    idxs = range(n) if point2D_idxs is None else point2D_idxs
    desc = loc.interpolate_descriptors(fmap, list(idxs), np.asarray(p2Ds, np.float64), interpolation_config)
    return [not isinstance(ref, np.ndarray) or bool(np.linalg.norm(d - ref.reshape(-1)) <= thresh)
            for d, ref in zip(desc, references)]


def find_unique_inliers(idxs, pre_inliers=None):
    """keep the first admissible correspondence of every keypoint index (reference main.py:53-63)"""
    keep, taken = [], set()
    for k, idx in enumerate(idxs):
        ok = (pre_inliers is None or bool(pre_inliers[k])) and idx not in taken
        if ok:
            taken.add(idx)
        keep.append(ok)
    return keep


def find_unique_min_by_group(errors, idxs, pre_inliers=None):
    """per keypoint index keep the admissible correspondence with the smallest error (first on ties; main.py:66-78)"""
    if len(idxs) != len(errors):
        raise ValueError("errors and idxs must have the same length")
    best = {}
    for k, (group, err) in enumerate(zip(idxs, errors)):
        if pre_inliers is not None and not pre_inliers[k]:
            continue
        if group not in best or err < best[group][1]:
            best[group] = (k, err)
    keep = [False] * len(idxs)
    for k, _ in best.values():
        keep[k] = True
    return keep


class QueryKeypointAdjuster:
    """keypoints of one query image against fixed reference descriptors (reference main.py:81-167)"""
    default_conf = defaults.query_keypoint_adjustment()

    def __init__(self, conf=None, callbacks=()):
        self.conf = merge(self.default_conf, conf or {})
        self.solver = loc.QueryKeypointOptimizer(optimizer_options(self.conf.optimizer, list(callbacks)),
                                                 to_ctr(self.conf.interpolation))

    def refine(self, pnp_points2D, fmap, references, point2D_idxs=None):
        inliers = find_feature_inliers(pnp_points2D, fmap, references, to_ctr(self.conf.interpolation),
                                       thresh=self.conf.feature_inlier_thresh, point2D_idxs=point2D_idxs)
        if self.conf.stack_correspondences:
            return self.refine_stacked(pnp_points2D, fmap, references, point2D_idxs, inliers=inliers)
        self.solver = loc.QueryKeypointOptimizer(self.solver.options, self.solver.interp)      # optimizers are one-shot
        self.solver.run(pnp_points2D, fmap, references, patch_idxs=point2D_idxs, inliers=inliers)

    def refine_multilevel(self, pnp_points2D, query_fmaps, references, point2D_idxs=None):
        for level in level_order(self.conf.level_indices, len(query_fmaps)):
            self.refine(pnp_points2D, query_fmaps[level], references[level], point2D_idxs=point2D_idxs)

    def refine_stacked(self, pnp_points2D, fmap, references, point2D_idxs, inliers=None):
        """correspondences that share a keypoint are stacked into ONE keypoint with several targets"""
        if point2D_idxs is None:
            raise ValueError("point2D_idxs must not be None in stacked QKA.")
        slot_of, slots = {}, []                       # keypoint index -> position among the distinct keypoints
        for p2D_idx in point2D_idxs:
            if p2D_idx not in slot_of:
                slot_of[p2D_idx] = len(slots)
                slots.append(p2D_idx)
        where = [slot_of[p] for p in point2D_idxs]
        stacked_kps = np.zeros((len(slots), 2))
        targets = [[] for _ in slots]
        admitted = [False] * len(slots)
        for k, ref in enumerate(references):
            if not isinstance(ref, np.ndarray):
                raise ValueError("Stacked QKA requires a np.ndarray reference for each 2D-3D correspondence. "
                                 "Consider setting target_references='nearest'.")
            stacked_kps[where[k]] = pnp_points2D[k]
            targets[where[k]].append(ref)
            admitted[where[k]] = admitted[where[k]] or inliers is None or bool(inliers[k])
        self.solver.run(stacked_kps, fmap, targets, patch_idxs=slots, inliers=None if inliers is None else admitted)
        for k, slot in enumerate(where):
            pnp_points2D[k] = stacked_kps[slot]


class QueryBundleAdjuster:
    """pose (and optionally intrinsics) of one query image against fixed 3D points (reference main.py:170-259)"""
    default_conf = defaults.query_bundle_adjustment()

    def __init__(self, conf=None, callbacks=()):
        self.conf = merge(self.default_conf, conf or {})
        self.solver = loc.QueryBundleOptimizer(optimizer_options(self.conf.optimizer, list(callbacks)),
                                               to_ctr(self.conf.interpolation))

    def refine(self, qvec, tvec, camera, points3D, fmap, references, inliers=None, point2D_idxs=None):
        return self.solver.run(qvec, tvec, camera, points3D, fmap, references, inliers=inliers, patch_idxs=point2D_idxs)

    def refine_multilevel(self, qvec, tvec, camera, points3D, fmaps, references, inliers=None, point2D_idxs=None):
        if len(fmaps) != len(references):
            raise ValueError("one reference list per feature level is required")
        for level in level_order(self.conf.level_indices, len(fmaps)):
            self.refine(qvec, tvec, camera, points3D, fmaps[level], references[level], inliers=inliers,
                        point2D_idxs=point2D_idxs)


def compute_reprojection_errors(pnp_points2D, pnp_points3D, qvec, tvec, camera):
    """pixel distance between every 2D point and the projection of its 3D point (reference main.py:96-102)"""
    from ..util.cameras import world_to_image
    xyz = np.array([np.asarray(getattr(p, "xyz", p), np.float64) for p in pnp_points3D]).reshape(-1, 3)
    proj = world_to_image(camera.model_id, np.asarray(camera.params, np.float64), np.asarray(qvec, np.float64),
                          np.asarray(tvec, np.float64), xyz)
    return [float(np.linalg.norm(proj[i] - np.asarray(p2D, np.float64))) for i, p2D in enumerate(pnp_points2D)]


def find_unique_min_reproj_inliers(pnp_points3D_id, qvec, tvec, camera, pnp_points2D, rec, pre_inliers=None, point2D_idxs=None):
    """of the correspondences that share a 3D point, then of those that share a keypoint, keep the one with the smallest
    reprojection error under the PnP pose (reference main.py:81-93)"""
    p3Ds = [rec.points3D[p3D_id] for p3D_id in pnp_points3D_id]
    errors = compute_reprojection_errors(pnp_points2D, p3Ds, qvec, tvec, camera)
    inliers = pre_inliers
    for idxs in (pnp_points3D_id, point2D_idxs):
        if idxs is None:
            continue
        inliers = find_unique_min_by_group(errors, idxs, pre_inliers=inliers)
    return inliers


class QueryLocalizer:
    """QKA -> PnP -> QBA for one query image (reference localization/main.py:262-537), with the reference's constructor
    and `localize` signatures.  Holds the reconstruction and one {point3D_id: Reference} map per feature level: given
    (`references=`), or extracted at construction from `dense_features` (a FeatureManager, or the path of a feature cache)
    or from features the `extractor` computes for `image_dir`.  The absolute-pose step is pycolmap.absolute_pose_estimation
    when pycolmap imports; `pose_estimator(points2D [N,2], points3D [N,3], camera) -> dict(success, qvec, tvec, inliers)`
    replaces it (pycolmap is outside this package)."""
    default_conf = defaults.query_localizer()

    def __init__(self, reconstruction, conf=None, dense_features=None, image_dir=None, references=None, extractor=None,
                 pose_estimator=None):
        from pathlib import Path
        from .. import bundle_adjustment as ba_pkg
        from .._pixsfm import _bundle_adjustment as ba
        if isinstance(conf, (str, Path)):
            from ..refine_colmap import _load_conf
            conf = _load_conf(conf)
        if isinstance(conf, dict) and isinstance(conf.get("localization"), dict):
            conf = conf["localization"]
        self.conf = merge(self.default_conf, conf or {})
        if self.conf.QKA.stack_correspondences and self.conf.target_reference not in ("nearest", "robust_mean"):
            raise ValueError("Stacked QKA requires a np.ndarray reference for each 2D-3D correspondence. Consider setting "
                             "target_references to 'nearest' or 'robust_mean'.")
        self.query_keypoint_adjuster = QueryKeypointAdjuster(to_ctr(self.conf.QKA))
        self.query_bundle_adjuster = QueryBundleAdjuster(to_ctr(self.conf.QBA))
        self.extractor = extractor
        self.pose_estimator = pose_estimator
        self.reference_extractor = ba.ReferenceExtractor(to_ctr(self.conf.references), to_ctr(self.conf.interpolation))
        self.target_reference_funcs = {"nearest": self.get_nearest_references, "robust_mean": self.get_robust_mean_references,
                                       "all_observations": self.get_all_references, "full": self.get_full_references}
        if self.conf.target_reference not in self.target_reference_funcs:
            raise ValueError("unknown target_reference %r" % (self.conf.target_reference,))
        self.get_query_references = self.target_reference_funcs[self.conf.target_reference]
        self.references = references
        if self.references is None and (self.conf.QKA.apply or self.conf.QBA.apply):
            cache_path = None
            if isinstance(dense_features, (str, Path)):
                cache_path = Path(dense_features)
                if cache_path.exists():
                    from ..features.store_features import load_features_from_cache
                    dense_features = load_features_from_cache(cache_path)
                else:
                    dense_features = None
            if dense_features is None:
                if image_dir is None or self.extractor is None:
                    raise ValueError("references (one {point3D_id: Reference} map per level), dense_features, or an extractor "
                                     "together with image_dir are required")
                dense_features = self.extractor.features_from_reconstruction(reconstruction, image_dir, cache_path=cache_path)
            labels = ba_pkg.find_problem_labels(reconstruction, self.conf.max_tracks_per_problem)
            self.references = [self.reference_extractor.run(labels, reconstruction, dense_features.fset(i))
                               for i in range(dense_features.num_levels)]
        self.reconstruction = reconstruction

    # ---- the reference descriptors a query's correspondences are compared with, one list per level (main.py:499-537)
    def get_nearest_references(self, pnp_points3D_id, query_fmaps, pnp_points2D, patch_idxs):
        return [loc.find_nearest_references(query_fmaps[level], refs, pnp_points2D, pnp_points3D_id,
                                            to_ctr(self.conf.interpolation), patch_idxs)
                for level, refs in enumerate(self.references)]

    def get_robust_mean_references(self, pnp_points3D_id, *args):
        return [[np.asarray(refs[p3D_id].descriptor, np.float64) for p3D_id in pnp_points3D_id] for refs in self.references]

    def get_all_references(self, pnp_points3D_id, *args):
        out = []
        for refs in self.references:
            level = []
            for p3D_id in pnp_points3D_id:
                if not len(refs[p3D_id].observations):
                    raise RuntimeError("Missing descriptors for observations.\nAssure that references.keep_observations==True.")
                level.append([np.asarray(o) for o in refs[p3D_id].observations])
            out.append(level)
        return out

    def get_full_references(self, pnp_points3D_id, *args):
        return [[refs[p3D_id] for p3D_id in pnp_points3D_id] for refs in self.references]

    def _query_features(self, image_path, keypoints, required_kp_ids):
        """the patches of the query image around the keypoints that take part in a correspondence (main.py:432-439)"""
        if self.extractor is None:
            raise ValueError("query_fmaps or an extractor (with image_path) are required")
        name = str(image_path)
        manager = self.extractor.features_from_image_list(None, [name], {name: keypoints}, {name: required_kp_ids})
        return [manager.fset(level).fmap(name) for level in range(manager.num_levels)]

    def _absolute_pose(self, pnp_points2D, pnp_points3D, query_camera, pose_estimator):
        estimator = pose_estimator or self.pose_estimator
        if estimator is not None:
            return estimator(pnp_points2D, np.array(pnp_points3D), query_camera)
        try:
            import pycolmap
        except ImportError:
            raise ValueError("a pose_estimator callable is required (pycolmap.absolute_pose_estimation is not importable)")
        return pycolmap.absolute_pose_estimation(pnp_points2D, pnp_points3D, query_camera,
                                                 estimation_options=to_ctr(self.conf.PnP.estimation),
                                                 refinement_options=to_ctr(self.conf.PnP.refinement))

    def localize(self, keypoints, pnp_point2D_idxs, pnp_points3D_id, query_camera, image_path=None, query_fmaps=None,
                 pose_estimator=None):
        """reference main.py:414-497.  `keypoints` [K,2] are the query image's keypoints, correspondence k pairs keypoint
        `pnp_point2D_idxs[k]` with 3D point `pnp_points3D_id[k]`.  Returns the pose dict of the PnP step with the refined
        pose, `inliers` / `num_inliers` recomputed from it, and (an addition) `keypoints`: the refined 2D points of the
        correspondences."""
        if len(pnp_point2D_idxs) != len(pnp_points3D_id):
            raise ValueError("one 3D point id per 2D index is required")
        if image_path is None and query_fmaps is None:
            raise ValueError("image_path or query_fmaps are required")
        if len(pnp_point2D_idxs) == 0:
            return {"success": False}
        pnp_point2D_idxs = [int(i) for i in pnp_point2D_idxs]
        pnp_points3D = [np.asarray(self.reconstruction.points3D[p3D_id].xyz, np.float64) for p3D_id in pnp_points3D_id]
        keypoints = np.array(keypoints, dtype=np.float64)
        require_feats = bool(self.conf.QKA.apply or self.conf.QBA.apply)
        if query_fmaps is None and require_feats:
            query_fmaps = self._query_features(image_path, keypoints, sorted(set(pnp_point2D_idxs)))
        pnp_points2D = keypoints[pnp_point2D_idxs]           # a copy: QKA refines the correspondences' points
        query_references = None
        if require_feats:
            query_references = self.get_query_references(pnp_points3D_id, query_fmaps, pnp_points2D, pnp_point2D_idxs)
            if len(query_fmaps) != len(query_references):
                raise ValueError("one reference map per feature level is required")
        if self.conf.QKA.apply:
            self.query_keypoint_adjuster.refine_multilevel(pnp_points2D, query_fmaps, query_references,
                                                           point2D_idxs=pnp_point2D_idxs)
        pose_dict = self._absolute_pose(pnp_points2D, pnp_points3D, query_camera, pose_estimator)
        if not pose_dict.get("success", False):
            return pose_dict
        inliers = pose_dict.get("inliers", [True] * len(pnp_points3D))
        if self.conf.unique_inliers:                          # None / False: keep what PnP says
            if self.conf.unique_inliers == "random":
                inliers = find_unique_inliers(pnp_points3D_id, pre_inliers=inliers)
            elif self.conf.unique_inliers == "min_error":
                inliers = find_unique_min_reproj_inliers(pnp_points3D_id, pose_dict["qvec"], pose_dict["tvec"], query_camera,
                                                         pnp_points2D, self.reconstruction, pre_inliers=inliers,
                                                         point2D_idxs=pnp_point2D_idxs)
            else:
                from .. import logger
                logger.warning("Unknown unique_inlier method %s.", self.conf.unique_inliers)
        if self.conf.QBA.apply:
            self.query_bundle_adjuster.refine_multilevel(pose_dict["qvec"], pose_dict["tvec"], query_camera, pnp_points3D,
                                                         query_fmaps, query_references, inliers=list(inliers),
                                                         point2D_idxs=pnp_point2D_idxs)
        errors = compute_reprojection_errors(pnp_points2D, pnp_points3D, pose_dict["qvec"], pose_dict["tvec"], query_camera)
        max_error = self.conf.PnP.estimation.ransac.max_error
        pose_dict["inliers"] = [err < max_error for err in errors]
        pose_dict["num_inliers"] = sum(pose_dict["inliers"])
        pose_dict["keypoints"] = pnp_points2D
        return pose_dict
