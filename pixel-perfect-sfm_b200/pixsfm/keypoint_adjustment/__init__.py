from .main import KeypointAdjuster, FeatureMetricKeypointAdjuster, find_problem_labels  # noqa: F401
from .._pixsfm._keypoint_adjustment import (KeypointAdjustmentSetup, KeypointOptimizerOptions,  # noqa: F401
                                            FeatureMetricKeypointOptimizer)
