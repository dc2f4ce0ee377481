from .main import (KeypointAdjuster, FeatureMetricKeypointAdjuster, TopologicalReferenceKeypointAdjuster,  # noqa: F401
                   find_problem_labels, build_matching_graph, extract_patchdata_from_graph)
from .._pixsfm._keypoint_adjustment import (KeypointAdjustmentSetup, KeypointOptimizerOptions,  # noqa: F401
                                            FeatureMetricKeypointOptimizer, TopologicalReferenceKeypointOptimizer,
                                            TopologicalReferenceKeypointOptimizerOptions)
