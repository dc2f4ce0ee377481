"""pixsfm.residuals — the `_residuals` cost-functor factories (reference: pixsfm/residuals/__init__.py re-exports
pixsfm._pixsfm._residuals)."""
from .._pixsfm._residuals import (FeatureReferenceCostFunctor, FeatureReferenceConstantPoseCostFunctor,  # noqa: F401
                                  FeatureMetricCostFunctor, GeometricCostFunctor, GeometricConstantPoseCostFunctor)
