"""tsfresh_b200 -- B200-native implementation of tsfresh's feature-extraction hot path.

Public surface = the reference's for this path: `extract_features`, the FC-parameter presets and the
imputation helpers that `extract_features(impute_function=...)` takes."""
from .extraction import _do_extraction_on_chunk, do_extraction_on_chunks, extract_features  # noqa: F401
from .dataframe_functions import (get_range_values_per_column, impute, impute_dataframe_range,  # noqa: F401
                                  impute_dataframe_zero)
from .feature_selection import calculate_relevance_table, select_features  # noqa: F401
from .rolling import RolledTimeSeries, roll_time_series  # noqa: F401
from .settings import (ComprehensiveFCParameters, EfficientFCParameters, MinimalFCParameters,  # noqa: F401
                       from_columns)

__all__ = ["extract_features", "do_extraction_on_chunks", "impute", "impute_dataframe_zero", "impute_dataframe_range",
           "get_range_values_per_column", "roll_time_series", "RolledTimeSeries", "select_features", "calculate_relevance_table", "from_columns", "ComprehensiveFCParameters", "EfficientFCParameters", "MinimalFCParameters"]
