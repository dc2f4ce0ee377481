"""Default option values of the reference's Python layer, in one place.  VALUES are the reference's (a drop-in must
start from the same options: base/main.py:1-22, keypoint_adjustment/main.py:60-83,206-250,
bundle_adjustment/main.py:30-62,218-240); each function returns a fresh nested dict."""
from copy import deepcopy

INTERPOLATION = dict(nodes=[[0.0, 0.0]], mode="BICUBIC", l2_normalize=True, ncc_normalize=False, use_float_simd=False)

# ceres::Solver::Options fields the reference sets for every optimizer
SOLVER = dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0, max_num_iterations=100,
              max_linear_solver_iterations=200, max_num_consecutive_invalid_steps=10,
              max_consecutive_nonmonotonic_steps=10, use_inner_iterations=False, use_nonmonotonic_steps=False,
              update_state_every_iteration=False, minimizer_progress_to_stdout=False, num_threads=-1)

CAUCHY = dict(name="cauchy", params=[0.25])


def interpolation():
    return deepcopy(INTERPOLATION)


def solver(**overrides):
    return dict(deepcopy(SOLVER), **overrides)


def keypoint_adjustment(**optimizer_extras):
    optimizer = dict(loss=deepcopy(CAUCHY), solver=solver(parameter_tolerance=1.0e-5, num_threads=1), print_summary=False,
                     bound=4.0, num_threads=-1, **optimizer_extras)
    return dict(strategy="featuremetric", apply=True, interpolation=interpolation(), level_indices=None,
                max_kps_per_problem=50, optimizer=optimizer, split_in_subproblems=True)


def bundle_adjustment():
    optimizer = dict(loss=deepcopy(CAUCHY), solver=solver(use_inner_iterations=True), print_summary=False,
                     refine_focal_length=True, refine_principal_point=False, refine_extra_params=True, refine_extrinsics=True)
    references = dict(loss=deepcopy(CAUCHY), iters=100, keep_observations=False, compute_offsets3D=False, num_threads=-1)
    return dict(strategy="feature_reference", apply=True, interpolation=interpolation(), level_indices=None,
                max_tracks_per_problem=10, optimizer=optimizer, references=references)


def costmaps():
    return dict(loss=dict(name="trivial", params=[]), as_gradientfield=True, compute_cross_derivative=False, num_threads=-1)


def query_keypoint_adjustment():
    optimizer = dict(loss=dict(name="trivial", params=[]), solver=solver(parameter_tolerance=1e-05), print_summary=False,
                     bound=4.0)
    return dict(apply=True, feature_inlier_thresh=-1, interpolation=interpolation(), level_indices=None,
                stack_correspondences=False, optimizer=optimizer)


def query_bundle_adjustment():
    optimizer = dict(loss=deepcopy(CAUCHY), solver=solver(), print_summary=False, refine_focal_length=False,
                     refine_principal_point=False, refine_extra_params=False)
    return dict(apply=True, interpolation=interpolation(), level_indices=None, optimizer=optimizer)


def query_localizer():
    # localization/main.py:262-299 (dense_features: the options of the extractor handed in, see features/extractor.py)
    return dict(dense_features={}, overwrite_features_sparse=None, interpolation=interpolation(), target_reference="nearest",
                unique_inliers="min_error",
                references=dict(loss=dict(name="cauchy", params=[0.25]), iters=100, keep_observations=True, num_threads=-1),
                max_tracks_per_problem=50, QKA=query_keypoint_adjustment(),
                PnP=dict(estimation=dict(ransac=dict(max_error=12)), refinement={}), QBA=query_bundle_adjustment())
