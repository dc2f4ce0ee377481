"""pixsfm.localization.main — the reference's query refinement surface (pixsfm/localization/main.py:20-330): helper
functions, QueryKeypointAdjuster, QueryBundleAdjuster and a QueryLocalizer whose PnP step is pluggable (the reference
calls pycolmap.absolute_pose_estimation, which is not part of this package)."""
from collections import defaultdict
from copy import deepcopy

import numpy as np

from .. import features, logger
from .._pixsfm import _localization as loc
from ..base import interpolation_default_conf, solver_default_conf
from ..util.conf import merge, to_ctr


def resolve_level_indices(level_indices, n_levels):      # util/misc.py:19-23
    if level_indices in [None, "all"]:
        return list(reversed(range(n_levels)))
    return list(level_indices)


def to_optim_ctr(cfg, callbacks):
    conf = to_ctr(cfg)
    conf["solver"]["callbacks"] = callbacks
    return conf


def find_feature_inliers(p2Ds, fmap, references, interpolation_config, thresh=-1, point2D_idxs=None):
    inliers = [True] * len(p2Ds)
    if thresh < 0.0:
        return inliers
    idxs = list(range(len(p2Ds))) if point2D_idxs is None else list(point2D_idxs)
    desc = loc.interpolate_descriptors(fmap, idxs, np.asarray(p2Ds, np.float64), interpolation_config)
    for i in range(len(p2Ds)):
        if isinstance(references[i], np.ndarray):
            if np.linalg.norm(desc[i] - np.asarray(references[i]).reshape(-1)) > thresh:
                inliers[i] = False
    return inliers


def find_unique_inliers(idxs, pre_inliers=None):
    unique_inliers = [False] * len(idxs)
    found = set()
    for i, idx in enumerate(idxs):
        if pre_inliers is not None and not pre_inliers[i]:
            continue
        if idx not in found:
            found.add(idx)
            unique_inliers[i] = True
    return unique_inliers


def find_unique_min_by_group(errors, idxs, pre_inliers=None):
    assert len(idxs) == len(errors)
    if pre_inliers is None:
        pre_inliers = [True] * len(idxs)
    errs_by_group = defaultdict(list)
    for i, (gid, err) in enumerate(zip(idxs, errors)):
        if pre_inliers[i]:
            errs_by_group[gid].append((i, err))
    min_errors = [min(vals, key=lambda item: item[1])[0] for vals in errs_by_group.values()]
    unique = np.array([False] * len(idxs))
    unique[min_errors] = True
    return list(unique)


class QueryKeypointAdjuster:
    default_conf = {
        'apply': True,
        'feature_inlier_thresh': -1,
        'interpolation': interpolation_default_conf,
        'level_indices': None,
        'stack_correspondences': False,
        'optimizer': {
            'loss': {'name': 'trivial', 'params': []},
            'solver': {**solver_default_conf, 'parameter_tolerance': 1e-05},
            'print_summary': False,
            'bound': 4.0
        }
    }

    def __init__(self, conf=None, callbacks=()):
        self.conf = merge(deepcopy(self.default_conf), conf or {})
        self.solver = loc.QueryKeypointOptimizer(to_optim_ctr(self.conf.optimizer, list(callbacks)), to_ctr(self.conf.interpolation))

    def refine(self, pnp_points2D, fmap, references, point2D_idxs=None):
        qka_inliers = find_feature_inliers(pnp_points2D, fmap, references, to_ctr(self.conf.interpolation),
                                           thresh=self.conf.feature_inlier_thresh, point2D_idxs=point2D_idxs)
        if self.conf.stack_correspondences:
            self.refine_stacked(pnp_points2D, fmap, references, point2D_idxs, inliers=qka_inliers)
        else:
            self.solver = loc.QueryKeypointOptimizer(self.solver.options, self.solver.interp)
            self.solver.run(pnp_points2D, fmap, references, patch_idxs=point2D_idxs, inliers=qka_inliers)

    def refine_multilevel(self, pnp_points2D, query_fmaps, references, point2D_idxs=None):
        for l_idx in resolve_level_indices(self.conf.level_indices, len(query_fmaps)):
            self.refine(pnp_points2D, query_fmaps[l_idx], references[l_idx], point2D_idxs=point2D_idxs)

    def refine_stacked(self, pnp_points2D, fmap, references, point2D_idxs, inliers=None):
        if point2D_idxs is None:
            raise ValueError("point2D_idxs must not be None in stacked QKA.")
        unique_p2D_idxs = list(set(point2D_idxs))
        old_to_new, unique_kps = [], [None for _ in unique_p2D_idxs]
        for idx, p2D_idx in enumerate(point2D_idxs):
            new_idx = unique_p2D_idxs.index(p2D_idx)
            old_to_new.append(new_idx)
            unique_kps[new_idx] = pnp_points2D[idx]
        unique_kps = np.array(unique_kps, dtype=np.float64)
        stacked_refs = [[] for _ in unique_p2D_idxs]
        for idx, query_ref in enumerate(references):
            if not isinstance(query_ref, np.ndarray):
                raise ValueError("Stacked QKA requires a np.ndarray reference for each 2D-3D correspondence. "
                                 "Consider setting target_references='nearest'.")
            stacked_refs[old_to_new[idx]].append(query_ref)
        # NB the reference passes per-correspondence `inliers` against the stacked keypoints; kept: only the length matters
        st_inl = None if inliers is None else [any(inliers[i] for i in range(len(point2D_idxs)) if old_to_new[i] == k)
                                              for k in range(len(unique_p2D_idxs))]
        self.solver.run(unique_kps, fmap, stacked_refs, patch_idxs=unique_p2D_idxs, inliers=st_inl)
        for i, _ in enumerate(point2D_idxs):
            pnp_points2D[i] = unique_kps[old_to_new[i]]


class QueryBundleAdjuster:
    default_conf = {
        'apply': True,
        'interpolation': interpolation_default_conf,
        'level_indices': None,
        'optimizer': {
            'loss': {'name': 'cauchy', 'params': [0.25]},
            'solver': {**solver_default_conf},
            'print_summary': False,
            'refine_focal_length': False,
            'refine_principal_point': False,
            'refine_extra_params': False,
        }
    }

    def __init__(self, conf=None, callbacks=()):
        self.conf = merge(deepcopy(self.default_conf), conf or {})
        self.solver = loc.QueryBundleOptimizer(to_optim_ctr(self.conf.optimizer, list(callbacks)), to_ctr(self.conf.interpolation))

    def refine(self, qvec, tvec, camera, points3D, fmap, references, inliers=None, point2D_idxs=None):
        return self.solver.run(qvec, tvec, camera, points3D, fmap, references, inliers=inliers, patch_idxs=point2D_idxs)

    def refine_multilevel(self, qvec, tvec, camera, points3D, fmaps, references, inliers=None, point2D_idxs=None):
        assert len(fmaps) == len(references)
        for level in resolve_level_indices(self.conf.level_indices, len(fmaps)):
            self.refine(qvec, tvec, camera, points3D, fmaps[level], references[level], inliers=inliers, point2D_idxs=point2D_idxs)


class QueryLocalizer:
    """QKA -> PnP -> QBA for one query image (reference localization/main.py:262-537).  Holds the reconstruction and
    one {point3D_id: Reference} map per feature level.  The absolute-pose estimator is a callable
    `pose_estimator(points2D [N,2], points3D [N,3], camera) -> dict(success, qvec, tvec, inliers)` — the reference uses
    pycolmap.absolute_pose_estimation, which is outside this package."""
    default_conf = {
        "interpolation": interpolation_default_conf,
        "target_reference": "nearest",
        "unique_inliers": "min_error",
        "QKA": {**QueryKeypointAdjuster.default_conf},
        "QBA": {**QueryBundleAdjuster.default_conf},
    }

    def __init__(self, reconstruction, conf=None, references=None, pose_estimator=None):
        self.conf = merge(deepcopy(self.default_conf), conf or {})
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
