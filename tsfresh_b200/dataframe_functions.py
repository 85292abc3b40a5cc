"""Device-side counterparts of the reference's imputation helpers
(tsfresh/utilities/dataframe_functions.py:49-212): same names, arguments, in-place behaviour and errors.

    impute(df)                         +inf -> column max, -inf -> column min, NaN -> column median (finite values);
                                       a column without any finite value is filled with 0
    impute_dataframe_zero(df)          every non-finite value -> 0
    impute_dataframe_range(df, ...)    the same replacement with caller-supplied values
    get_range_values_per_column(df)    (col_to_max, col_to_min, col_to_median)

The matrix goes to the GPU as one row-major float64 block; `extract_features(impute_function=impute)` skips that
round trip and imputes the feature matrix while it is still on the device (TSFX_FLAG_IMPUTE).
"""
import warnings

import numpy as np
import pandas as pd

from . import _lib


def _matrix(df):
    m = df.to_numpy(dtype=np.float64, copy=True)
    return np.ascontiguousarray(m)


def _write_back(df, m):
    # in place, like the reference (DataFrame.where(..., inplace=True)); all columns end up float64
    for j, c in enumerate(df.columns):
        df[c] = m[:, j]
    return df


def _ctx(device=None):
    from .extraction import get_context
    return get_context(device)


def impute(df_impute, device=None):
    """dataframe_functions.py:49-78."""
    if len(df_impute) == 0:
        return df_impute
    m = _matrix(df_impute)
    stats = _ctx(device).impute(m, _lib.IMPUTE_RANGE)
    _warn_non_finite(df_impute, m, stats)
    return _write_back(df_impute, m)


def impute_dataframe_zero(df_impute, device=None):
    """dataframe_functions.py:81-101."""
    if len(df_impute) == 0:
        return df_impute
    m = _matrix(df_impute)
    _ctx(device).impute(m, _lib.IMPUTE_ZERO)
    return _write_back(df_impute, m)


def impute_dataframe_range(df_impute, col_to_max, col_to_min, col_to_median, device=None):
    """dataframe_functions.py:104-167 (same ValueErrors)."""
    if len(df_impute) == 0:
        return df_impute
    columns = df_impute.columns
    if (not set(columns) <= set(col_to_median.keys()) or not set(columns) <= set(col_to_max.keys())
            or not set(columns) <= set(col_to_min.keys())):
        raise ValueError("Some of the dictionaries col_to_median, col_to_max, col_to_min contains more or less keys "
                         "than the column names in df")
    if (np.any(~np.isfinite(list(col_to_median.values()))) or np.any(~np.isfinite(list(col_to_min.values())))
            or np.any(~np.isfinite(list(col_to_max.values())))):
        raise ValueError("Some of the dictionaries col_to_median, col_to_max, col_to_min contains non finite values "
                         "to replace")
    stats = np.array([[col_to_min[c] for c in columns], [col_to_max[c] for c in columns],
                      [col_to_median[c] for c in columns]], dtype=np.float64)
    m = _matrix(df_impute)
    _ctx(device).impute(m, _lib.IMPUTE_GIVEN, col_stats=stats)
    return _write_back(df_impute, m)


def get_range_values_per_column(df, device=None):
    """dataframe_functions.py:170-212: three dicts column -> finite max / min / median (0 without finite values)."""
    m = _matrix(df)
    stats = _ctx(device).impute(m, _lib.IMPUTE_STATS)
    _warn_non_finite(df, m, stats, check_input=True)
    cols = df.columns
    return dict(zip(cols, stats[1])), dict(zip(cols, stats[0])), dict(zip(cols, stats[2]))


def _warn_non_finite(df, m, stats, check_input=False):
    # the reference warns about columns without any finite value (:194-201); after RANGE they are all-zero and had
    # min = max = median = 0, so look at the statistics (cheap: 3 x cols)
    zero = (stats[0] == 0.0) & (stats[1] == 0.0) & (stats[2] == 0.0)
    if zero.any():
        src = df.to_numpy(dtype=np.float64)[:, zero]
        bad = ~np.isfinite(src).any(axis=0)
        if bad.any():
            warnings.warn("The columns {} did not have any finite values. Filling with zeros.".format(
                df.columns[np.where(zero)[0][bad]].values), RuntimeWarning)
