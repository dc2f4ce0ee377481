"""pixsfm.base — the `_base` bindings (graph, labelling, interpolation config) and the two default option blocks the
reference keeps in pixsfm/base/main.py."""
from .. import defaults as _defaults
from .._pixsfm._base import (Graph, FeatureNode, Match, InterpolationConfig, InterpolatorType,  # noqa: F401
                             compute_track_labels, compute_score_labels, compute_root_labels, count_track_edges,
                             count_edges_AB)

Map_NameKeypoints = dict  # image name -> [N,2] float64 array (base/bindings.cc:18,116)

interpolation_default_conf = _defaults.interpolation()
solver_default_conf = _defaults.solver()
