"""tsfresh_b200 -- B200-native implementation of tsfresh's feature-extraction hot path.

Public surface = the reference's for this path: `extract_features` and the FC-parameter presets."""
from .extraction import extract_features  # noqa: F401
from .settings import ComprehensiveFCParameters, EfficientFCParameters, MinimalFCParameters  # noqa: F401

__all__ = ["extract_features", "ComprehensiveFCParameters", "EfficientFCParameters", "MinimalFCParameters"]
