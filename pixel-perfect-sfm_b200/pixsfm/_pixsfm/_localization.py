"""Mirror of `pixsfm._pixsfm._localization` (pixsfm/localization/bindings.cc; src/single_query_keypoint_optimizer.h,
single_query_bundle_optimizer.h, query_refinement_options.h, nearest_references.h): query keypoint adjustment and
query bundle adjustment against fixed reference descriptors.  Same kernels as KA / BA:
  * QKA  = pxr_ka_run in query mode (every edge = keypoint vs fixed descriptor; block-diagonal normal equations,
           all keypoints of the query in one trust region like the reference's single ceres::Problem)
  * QBA  = pxr_ba_run with one free image, every 3D point constant (one IR "point" per (correspondence, reference))
"""
import ctypes as C

import numpy as np

from .. import logger
from . import _capi, _engine
from ._base import InterpolationConfig
from ._bundle_adjustment import _DictOptions, _Summary, solver_options_from
from ._features import DevicePatches, Reference


class QueryKeypointOptimizerOptions(_DictOptions):
    """query_refinement_options.h:54-98 (C++ defaults; the Python layer passes parameter_tolerance 1e-5, bound 4)"""
    _defaults = dict(print_summary=True, bound=-1.0, loss=lambda: {"name": "trivial", "params": []},
                     solver=lambda: {"function_tolerance": 0.0, "gradient_tolerance": 0.0, "parameter_tolerance": 1.0e-4,
                                     "max_num_iterations": 100, "max_num_consecutive_invalid_steps": 10})


class QueryBundleOptimizerOptions(_DictOptions):
    """query_refinement_options.h:7-52"""
    _defaults = dict(print_summary=True, refine_focal_length=False, refine_principal_point=False, refine_extra_params=False,
                     loss=lambda: {"name": "cauchy", "params": [0.25]},
                     solver=lambda: {"function_tolerance": 0.0, "gradient_tolerance": 0.0, "parameter_tolerance": 1.0e-5,
                                     "max_num_iterations": 100, "max_linear_solver_iterations": 200,
                                     "max_num_consecutive_invalid_steps": 10, "max_consecutive_nonmonotonic_steps": 10})


def _host_patches(fmap):
    if isinstance(fmap.patches, DevicePatches):
        raise ValueError("query refinement takes host (numpy) feature maps")
    return fmap.patches


def _descriptors_of(ref, channels):
    """the reference descriptors one correspondence is compared with (single_query_*_optimizer.h RunQuery overloads):
    ndarray -> itself; list of ndarrays -> all of them; Reference -> its observations if kept, else its descriptor"""
    if isinstance(ref, Reference):
        obs = getattr(ref, "observations", None)
        if obs:
            return [np.asarray(o, np.float64).reshape(-1)[:channels] for o in obs]
        return [np.asarray(ref.descriptor, np.float64).reshape(-1)[:channels]]
    if isinstance(ref, np.ndarray):
        return [np.asarray(ref, np.float64).reshape(-1)[:channels]]
    return [np.asarray(r, np.float64).reshape(-1)[:channels] for r in ref]


def interpolate_descriptors(fmap, patch_idxs, xys, interpolation_config=None):
    """PatchInterpolator.interpolate_nodes for one node, batched: descriptor of fmap's patch patch_idxs[i] at xys[i]"""
    interp = interpolation_config if isinstance(interpolation_config, InterpolationConfig) else InterpolationConfig(interpolation_config or {})
    interp.validate_for_device()
    patches = _host_patches(fmap)
    ctx = _capi.default_context()
    xys = np.ascontiguousarray(xys, np.float64).reshape(-1, 2)
    item_patch = np.array([fmap.local_index(i) for i in patch_idxs], np.int64)
    scales = np.ascontiguousarray(np.tile(fmap.scale, (fmap.size(), 1)), np.float64)
    out = np.zeros((len(xys), fmap.channels))
    ic = _capi.default_interp(interp.l2_normalize, interp.use_float_simd)
    _capi.check(ctx.lib.pxr_interpolate_descriptors(
        ctx.handle, patches.ctypes.data_as(C.c_void_p), C.c_int64(fmap.size()), _capi.DTYPE_IDS[patches.dtype],
        int(patches.shape[1]), int(patches.shape[2]), int(patches.shape[3]), fmap.corners.ctypes.data_as(C.c_void_p),
        scales.ctypes.data_as(C.c_void_p), C.c_double(1.0), C.c_int64(len(xys)), item_patch.ctypes.data_as(C.c_void_p),
        xys.ctypes.data_as(C.c_void_p), C.byref(ic), out.ctypes.data_as(C.c_void_p)))
    return out


def find_nearest_references(query_fmap, references, keypoints, point3D_ids, interpolation_config, patch_idxs=None):
    """localization/src/nearest_references.h:20-53: per correspondence the stored observation descriptor of the 3D
    point that is closest (squared L2) to the query descriptor interpolated at the keypoint; first minimum wins."""
    keypoints = np.asarray(keypoints, np.float64).reshape(-1, 2)
    idxs = list(range(len(keypoints))) if patch_idxs is None else list(patch_idxs)
    q = interpolate_descriptors(query_fmap, idxs, keypoints, interpolation_config)
    out = []
    for i, pid in enumerate(point3D_ids):
        ref = references[pid]
        obs = getattr(ref, "observations", None)
        if not obs:
            raise ValueError("Missing observations in references. Extract references with option keep_observations=True.")
        obs = [np.asarray(o, np.float64).reshape(-1) for o in obs]
        d = [float(((o - q[i]) ** 2).sum()) for o in obs]
        out.append(obs[int(np.argmin(d))].reshape(1, -1).copy())
    return out


class QueryKeypointOptimizer:
    """_localization.QueryKeypointOptimizer(options, interpolation).run(keypoints, fmap, references, patch_idxs=None,
    inliers=None) -> bool; `keypoints` ([N,2] float64) is refined IN PLACE."""

    def __init__(self, options, interpolation_config):
        self.options = options if isinstance(options, QueryKeypointOptimizerOptions) else QueryKeypointOptimizerOptions(options)
        self.interp = interpolation_config if isinstance(interpolation_config, InterpolationConfig) else InterpolationConfig(interpolation_config)
        self._summary = None

    def build_problem(self, keypoints, fmap, references, patch_idxs=None, inliers=None):
        """-> (KAProblem in query mode | None, indices of the keypoints it holds)"""
        if keypoints.dtype != np.float64 or keypoints.ndim != 2 or keypoints.shape[1] != 2:
            raise ValueError("keypoints must be a [N,2] float64 array")
        n = len(keypoints)
        if len(references) != n:
            raise ValueError("references.size() != keypoints.rows()")       # THROW_CHECK_EQ
        if patch_idxs is not None and len(patch_idxs) != n:
            raise ValueError("patch_idxs.size() != keypoints.rows()")
        self.interp.validate_for_device()
        if len(self.interp.nodes) != 1 or fmap.channels not in (8, 16, 32, 64, 128):
            raise ValueError("Unsupported dimensions (CHANNELS,N_NODES).")
        patches = _host_patches(fmap)
        es, ed, refs, used = [], [], [], []
        for i in range(n):
            if inliers is not None and not inliers[i]:
                continue
            descs = _descriptors_of(references[i], fmap.channels)
            for dsc in descs:
                es.append(len(used)); ed.append(len(refs)); refs.append(dsc)
            if descs:
                used.append(i)
        if not es:
            return None, used                                               # problem->NumResiduals() == 0
        kp_patch = np.array([fmap.local_index(i if patch_idxs is None else patch_idxs[i]) for i in used], np.int64)
        kps = np.ascontiguousarray(keypoints[used])
        prob = _capi.KAProblem(keypoints=kps, kp_const=np.zeros(len(used), np.uint8), edge_src=es, edge_dst=ed,
                               edge_weight=None, edge_problem=np.zeros(len(es), np.int32), n_problems=1, patches=patches,
                               corner=fmap.corners, scale=np.tile(fmap.scale, (fmap.size(), 1)), kp_patch=kp_patch,
                               bound=float(self.options.bound), patches_are_sparse=fmap.is_sparse, ref_desc=np.array(refs))
        return prob, used

    def solver_options(self):
        return solver_options_from(self.options.loss, self.options.solver, _capi.default_ka_options())

    def run(self, keypoints, fmap, references, patch_idxs=None, inliers=None):
        prob, used = self.build_problem(keypoints, fmap, references, patch_idxs, inliers)
        if prob is None:
            return False
        so = self.solver_options()
        ic = _capi.default_interp(self.interp.l2_normalize, self.interp.use_float_simd)
        s = _engine.ka_run(prob, ic, so)
        keypoints[used] = prob.keypoints
        nres = max(1, len(prob.edge_src) * fmap.channels)
        self._summary = _Summary(s, num_residuals_reduced=nres)
        logger.info("QKA Time: %.4gs, cost change: %.6g --> %.6g", s["total_time_s"], np.sqrt(s["initial_cost"] / nres),
                    np.sqrt(s["final_cost"] / nres))
        return True

    def summary(self):
        return self._summary


class QueryBundleOptimizer:
    """_localization.QueryBundleOptimizer(options, interpolation).run(qvec, tvec, camera, points3D, fmap, references,
    inliers=None, patch_idxs=None) -> bool; qvec / tvec / camera.params are refined IN PLACE."""

    def __init__(self, options, interpolation_config):
        self.options = options if isinstance(options, QueryBundleOptimizerOptions) else QueryBundleOptimizerOptions(options)
        self.interp = interpolation_config if isinstance(interpolation_config, InterpolationConfig) else InterpolationConfig(interpolation_config)
        self._summary = None

    def build_problem(self, qvec, tvec, camera, points3D, fmap, references, inliers=None, patch_idxs=None):
        n = len(points3D)
        if len(references) != n:
            raise ValueError("references.size() != points3D.size()")
        if patch_idxs is not None and len(patch_idxs) != n:
            raise ValueError("patch_idxs.size() != points3D.size()")
        self.interp.validate_for_device()
        if len(self.interp.nodes) != 1 or fmap.channels not in (8, 16, 32, 64, 128, 256):
            raise ValueError("Unsupported dimensions (CHANNELS,N_NODES).")
        patches = _host_patches(fmap)
        xyz, refs, obs_patch = [], [], []
        for i in range(n):
            if inliers is not None and not inliers[i]:
                continue
            pi = fmap.local_index(i if patch_idxs is None else patch_idxs[i])
            for dsc in _descriptors_of(references[i], fmap.channels):   # one residual block per reference descriptor
                xyz.append(np.asarray(points3D[i], np.float64).reshape(3)); refs.append(dsc); obs_patch.append(pi)
        if not xyz:
            return None
        m = len(xyz)
        model = int(camera.model_id)
        focal, pp, extra = _capi.CAMERA_PARAM_GROUPS[model]
        o = self.options
        if not (o.refine_focal_length or o.refine_principal_point or o.refine_extra_params):
            mask = 0xFFFFFFFF                                            # ParameterizeQuery: constant camera
        else:
            mask = (0 if o.refine_focal_length else focal) | (0 if o.refine_principal_point else pp) | (0 if o.refine_extra_params else extra)
        prob = _capi.BAProblem(cam_model=[model], cam_params=[np.asarray(camera.params, np.float64)], cam_const_mask=[mask],
                               qvec=np.asarray(qvec, np.float64).reshape(1, 4), tvec=np.asarray(tvec, np.float64).reshape(1, 3),
                               img_cam=[0], pose_const=[0], tvec_const_mask=[0], xyz=np.array(xyz), point_const=np.ones(m, np.uint8),
                               obs_img=np.zeros(m, np.int32), obs_pt=np.arange(m, dtype=np.int64), patches=patches,
                               corner=fmap.corners, scale=np.tile(fmap.scale, (fmap.size(), 1)), refs=np.array(refs),
                               obs_patch=np.array(obs_patch, np.int64))
        return prob

    def solver_options(self):
        return solver_options_from(self.options.loss, self.options.solver,
                                   _capi.default_ba_options(use_inner_iterations=0, parameter_tolerance=1e-5))

    def run(self, qvec, tvec, camera, points3D, fmap, references, inliers=None, patch_idxs=None):
        prob = self.build_problem(qvec, tvec, camera, points3D, fmap, references, inliers, patch_idxs)
        if prob is None:
            return False
        so = self.solver_options()
        ic = _capi.default_interp(self.interp.l2_normalize, self.interp.use_float_simd)
        s = _engine.ba_run(prob, ic, so)
        np.asarray(qvec).reshape(-1)[:] = prob.qvec[0]
        np.asarray(tvec).reshape(-1)[:] = prob.tvec[0]
        camera.params[:] = prob.cam_params[0, :len(camera.params)]
        nres = max(1, s["num_residuals"])
        self._summary = _Summary(s, num_residuals_reduced=nres)
        logger.info("QBA Time: %.4gs, cost change: %.6g --> %.6g", s["total_time_s"], np.sqrt(s["initial_cost"] / nres),
                    np.sqrt(s["final_cost"] / nres))
        return True

    def summary(self):
        return self._summary
