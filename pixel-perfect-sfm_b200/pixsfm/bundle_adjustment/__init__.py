from .main import (BundleAdjuster, FeatureReferenceBundleAdjuster, CostMapBundleAdjuster, default_problem_setup,  # noqa: F401
                   find_problem_labels)
from .._pixsfm._bundle_adjustment import (BundleAdjustmentSetup, BundleOptimizerOptions, ReferenceConfig,  # noqa: F401
                                          ReferenceExtractor, FeatureReferenceBundleOptimizer, CostMapConfig,
                                          CostMapExtractor, CostMapBundleOptimizer)
