"""pixsfm.features — the feature containers consumed by the hot path (reference pixsfm/features/bindings.cc:38-300).
CNN feature extraction (features/extractor.py, models/) is unchanged reference territory and out of scope."""
from .._pixsfm._features import (FeaturePatch, FeatureMap, FeatureSet, FeatureView, FeatureManager,  # noqa: F401
                                 Reference, kDenseId)

Map_IdReference = dict
