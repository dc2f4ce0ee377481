from .main import (KeypointAdjuster, FeatureMetricKeypointAdjuster, TopologicalReferenceKeypointAdjuster,  # noqa: F401
                   find_problem_labels)
from .._pixsfm._keypoint_adjustment import (KeypointAdjustmentSetup, KeypointOptimizerOptions,  # noqa: F401
                                            FeatureMetricKeypointOptimizer, TopologicalReferenceKeypointOptimizer,
                                            TopologicalReferenceKeypointOptimizerOptions)
