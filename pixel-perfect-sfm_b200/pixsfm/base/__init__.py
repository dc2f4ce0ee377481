"""pixsfm.base — reference pixsfm/base/main.py:1-22 + the `_base` bindings."""
from .._pixsfm._base import (Graph, FeatureNode, Match, InterpolationConfig, InterpolatorType,  # noqa: F401
                             compute_track_labels, compute_score_labels, compute_root_labels, count_track_edges)

Map_NameKeypoints = dict  # image name -> [N,2] float64 array (base/bindings.cc:18,116)

interpolation_default_conf = {
    'nodes': [[0.0, 0.0]],
    'mode': 'BICUBIC',
    'l2_normalize': True,
    'ncc_normalize': False,
    "use_float_simd": False
}

solver_default_conf = {
    'function_tolerance': 0.0,
    'gradient_tolerance': 0.0,
    'parameter_tolerance': 0.0,
    'minimizer_progress_to_stdout': False,
    'max_num_iterations': 100,
    'max_linear_solver_iterations': 200,
    'max_num_consecutive_invalid_steps': 10,
    'max_consecutive_nonmonotonic_steps': 10,
    'use_inner_iterations': False,
    'use_nonmonotonic_steps': False,
    'update_state_every_iteration': False,
    'num_threads': -1,
}
