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


class QueryLocalizer:
    """QKA -> PnP -> QBA for one query image (reference localization/main.py:262-537).  Holds the reconstruction and
    one {point3D_id: Reference} map per feature level.  The absolute-pose estimator is a callable
    `pose_estimator(points2D [N,2], points3D [N,3], camera) -> dict(success, qvec, tvec, inliers)` — the reference uses
    pycolmap.absolute_pose_estimation, which is outside this package."""
    default_conf = defaults.query_localizer()

    def __init__(self, reconstruction, conf=None, references=None, pose_estimator=None):
        self.conf = merge(self.default_conf, conf or {})
        self.reconstruction = reconstruction
        if references is None:
            raise ValueError("references (one {point3D_id: Reference} map per level) are required")
        self.references = references
        self.pose_estimator = pose_estimator

    def _target_references(self, level, fmap, keypoints, point3D_ids, point2D_idxs):
        refs = self.references[level]
        if self.conf.target_reference == "nearest":
            return loc.find_nearest_references(fmap, refs, keypoints, point3D_ids, to_ctr(self.conf.interpolation), point2D_idxs)
        if self.conf.target_reference == "robust_mean":
            return [refs[p] for p in point3D_ids]
        if self.conf.target_reference == "all_observations":
            return [[np.asarray(o) for o in refs[p].observations] for p in point3D_ids]
        raise ValueError("unknown target_reference %r" % (self.conf.target_reference,))

    def localize(self, pnp_points2D, pnp_point3D_ids, query_camera, query_fmaps, pnp_point2D_idxs=None, pose_estimator=None):
        keypoints = np.ascontiguousarray(pnp_points2D, np.float64).copy()
        levels = resolve_level_indices(self.conf.QKA.level_indices, len(query_fmaps))
        if self.conf.QKA.apply:
            qka = QueryKeypointAdjuster(to_ctr(self.conf.QKA))
            for level in levels:
                refs = self._target_references(level, query_fmaps[level], keypoints, pnp_point3D_ids, pnp_point2D_idxs)
                qka.refine(keypoints, query_fmaps[level], refs, point2D_idxs=pnp_point2D_idxs)
        estimator = pose_estimator or self.pose_estimator
        if estimator is None:
            raise ValueError("a pose_estimator callable is required (pycolmap.absolute_pose_estimation is not bundled)")
        points3D = [np.asarray(self.reconstruction.points3D[p].xyz, np.float64) for p in pnp_point3D_ids]
        pose = estimator(keypoints, np.array(points3D), query_camera)
        if not pose.get("success", False):
            return pose
        if self.conf.QBA.apply:
            inliers = list(pose.get("inliers", [True] * len(points3D)))
            if self.conf.unique_inliers == "min_error" and pnp_point2D_idxs is not None:
                inliers = find_unique_inliers(pnp_point2D_idxs, pre_inliers=inliers)
            qba = QueryBundleAdjuster(to_ctr(self.conf.QBA))
            for level in resolve_level_indices(self.conf.QBA.level_indices, len(query_fmaps)):
                refs = self._target_references(level, query_fmaps[level], keypoints, pnp_point3D_ids, pnp_point2D_idxs)
                qba.refine(pose["qvec"], pose["tvec"], query_camera, points3D, query_fmaps[level], refs, inliers=inliers,
                           point2D_idxs=pnp_point2D_idxs)
        pose["keypoints"] = keypoints
        return pose
