"""FC-parameter dictionaries: the drop-in configuration surface of the hot path.

Mirrors the reference's settings objects (tsfresh/feature_extraction/settings.py:133-343): a mapping
`{calculator_name: None | [param_dict, ...]}` whose iteration order defines the output column order
(extraction.py:339-378).  The registry below lists the calculators in the order the reference's
`ComprehensiveFCParameters.__init__` discovers them (definition order of feature_calculators.py for
parameter-free calculators, then the parameter grid of settings.py:165-280).

A plain dict (or the reference's own settings object) works equally well as input to
`tsfresh_b200.extract_features`; these classes only provide the same three presets.
"""
from collections import UserDict
from itertools import product

# (name, flags) in reference definition order.  flags: m = "minimal" attribute, h = "high_comp_cost",
# t = needs a pd.DatetimeIndex (extraction.py:349-358: skipped with a warning otherwise).
_PARAMETER_FREE = [
    ("variance_larger_than_standard_deviation", ""), ("has_duplicate_max", ""), ("has_duplicate_min", ""),
    ("has_duplicate", ""), ("sum_values", "m"), ("abs_energy", ""), ("mean_abs_change", ""),
    ("mean_change", ""), ("mean_second_derivative_central", ""), ("median", "m"), ("mean", "m"),
    ("length", "m"), ("standard_deviation", "m"), ("variation_coefficient", ""), ("variance", "m"),
    ("skewness", ""), ("kurtosis", ""), ("root_mean_square", "m"), ("absolute_sum_of_changes", ""),
    ("longest_strike_below_mean", ""), ("longest_strike_above_mean", ""), ("count_above_mean", ""),
    ("count_below_mean", ""), ("last_location_of_maximum", ""), ("first_location_of_maximum", ""),
    ("last_location_of_minimum", ""), ("first_location_of_minimum", ""),
    ("percentage_of_reoccurring_values_to_all_values", ""),
    ("percentage_of_reoccurring_datapoints_to_all_datapoints", ""), ("sum_of_reoccurring_values", ""),
    ("sum_of_reoccurring_data_points", ""), ("ratio_value_number_to_time_series_length", ""),
    ("sample_entropy", "h"), ("maximum", "m"), ("absolute_maximum", "m"), ("minimum", "m"),
    ("benford_correlation", ""),
]

HIGH_COMP_COST = {"sample_entropy", "approximate_entropy"}
MINIMAL = {n for n, f in _PARAMETER_FREE if "m" in f}
NEEDS_DATETIME_INDEX = {"linear_trend_timewise"}


def _parameter_grid():
    """The hand-written grid of settings.py:165-280 (same literals, same order)."""
    five = [{"attr": a} for a in ("pvalue", "rvalue", "intercept", "slope", "stderr")]
    return {
        "time_reversal_asymmetry_statistic": [{"lag": lag} for lag in range(1, 4)],
        "c3": [{"lag": lag} for lag in range(1, 4)],
        "cid_ce": [{"normalize": True}, {"normalize": False}],
        "symmetry_looking": [{"r": r * 0.05} for r in range(20)],
        "large_standard_deviation": [{"r": r * 0.05} for r in range(1, 20)],
        "quantile": [{"q": q} for q in (0.1, 0.2, 0.3, 0.4, 0.6, 0.7, 0.8, 0.9)],
        "autocorrelation": [{"lag": lag} for lag in range(10)],
        "agg_autocorrelation": [{"f_agg": s, "maxlag": 40} for s in ("mean", "median", "var")],
        "partial_autocorrelation": [{"lag": lag} for lag in range(10)],
        "number_cwt_peaks": [{"n": n} for n in (1, 5)],
        "number_peaks": [{"n": n} for n in (1, 3, 5, 10, 50)],
        "binned_entropy": [{"max_bins": b} for b in (10,)],
        "index_mass_quantile": [{"q": q} for q in (0.1, 0.2, 0.3, 0.4, 0.6, 0.7, 0.8, 0.9)],
        "cwt_coefficients": [{"widths": width, "coeff": coeff, "w": w}
                             for width in [(2, 5, 10, 20)] for coeff in range(15) for w in (2, 5, 10, 20)],
        "spkt_welch_density": [{"coeff": c} for c in (2, 5, 8)],
        "ar_coefficient": [{"coeff": c, "k": k} for c in range(10 + 1) for k in (10,)],
        "change_quantiles": [{"ql": ql, "qh": qh, "isabs": b, "f_agg": f}
                             for ql in (0.0, 0.2, 0.4, 0.6, 0.8) for qh in (0.2, 0.4, 0.6, 0.8, 1.0)
                             for b in (False, True) for f in ("mean", "var") if ql < qh],
        "fft_coefficient": [{"coeff": k, "attr": a}
                            for a, k in product(("real", "imag", "abs", "angle"), range(100))],
        "fft_aggregated": [{"aggtype": s} for s in ("centroid", "variance", "skew", "kurtosis")],
        "value_count": [{"value": v} for v in (0, 1, -1)],
        "range_count": [{"min": -1, "max": 1}, {"min": -1e12, "max": 0}, {"min": 0, "max": 1e12}],
        "approximate_entropy": [{"m": 2, "r": r} for r in (0.1, 0.3, 0.5, 0.7, 0.9)],
        "friedrich_coefficients": [{"coeff": c, "m": 3, "r": 30} for c in range(3 + 1)],
        "max_langevin_fixed_point": [{"m": 3, "r": 30}],
        "linear_trend": list(five),
        "agg_linear_trend": [{"attr": attr, "chunk_len": i, "f_agg": f}
                             for attr in ("rvalue", "intercept", "slope", "stderr")
                             for i in (5, 10, 50) for f in ("max", "min", "mean", "var")],
        "augmented_dickey_fuller": [{"attr": "teststat"}, {"attr": "pvalue"}, {"attr": "usedlag"}],
        "number_crossing_m": [{"m": 0}, {"m": -1}, {"m": 1}],
        "energy_ratio_by_chunks": [{"num_segments": 10, "segment_focus": i} for i in range(10)],
        "ratio_beyond_r_sigma": [{"r": x} for x in (0.5, 1, 1.5, 2, 2.5, 3, 5, 6, 7, 10)],
        "linear_trend_timewise": [dict(d) for d in five],
        "count_above": [{"t": 0}],
        "count_below": [{"t": 0}],
        "lempel_ziv_complexity": [{"bins": x} for x in (2, 3, 5, 10, 100)],
        "fourier_entropy": [{"bins": x} for x in (2, 3, 5, 10, 100)],
        "permutation_entropy": [{"tau": 1, "dimension": x} for x in (3, 4, 5, 6, 7)],
        "query_similarity_count": [{"query": None, "threshold": 0.0}],
        # settings.py:272-278 writes one dict literal with the key repeated three times; Python keeps
        # the last one, so only number_of_maxima=7 exists.  (matrix_profile is dropped by the
        # reference when its optional dependency is missing, settings.py:282-292; it is not provided.)
        "mean_n_absolute_max": [{"number_of_maxima": 7}],
    }


class ComprehensiveFCParameters(UserDict):
    """All 75 calculators / 788 features (783 without a DatetimeIndex); settings.py:133-294."""

    def __init__(self):
        d = {name: None for name, _ in _PARAMETER_FREE}
        d.update(_parameter_grid())
        super().__init__(d)


class MinimalFCParameters(ComprehensiveFCParameters):
    """Only the calculators tagged "minimal" (settings.py:297-320)."""

    def __init__(self):
        super().__init__()
        for k in [k for k in self.data if k not in MINIMAL]:
            del self.data[k]


class EfficientFCParameters(ComprehensiveFCParameters):
    """Everything except the "high_comp_cost" calculators (settings.py:323-343)."""

    def __init__(self):
        super().__init__()
        for k in [k for k in self.data if k in HIGH_COMP_COST]:
            del self.data[k]


# ------------------------------------------------------------------ column names -> settings (settings.py:23-83)
def get_config_from_string(parts):
    """Inverse of the feature-name grammar (utilities/string_manipulation.py:10-44): `parts` is a column name split
    on "__"; everything after <kind>, <calculator> is `<parameter name>_<python literal>`.  None without parameters."""
    import ast

    import numpy as np
    if len(parts) <= 2:
        return None
    config = {}
    for piece in parts[2:]:
        key, value = piece.rsplit("_", 1)
        low = value.lower()
        if low == "nan":
            config[key] = np.nan
        elif low == "-inf":
            config[key] = -np.inf
        elif low == "inf":
            config[key] = np.inf
        else:
            config[key] = ast.literal_eval(value)
    return config


def from_columns(columns, columns_to_ignore=None):
    """`kind_to_fc_parameters` that extracts exactly the features named in `columns` (the reference's
    tsfresh.feature_extraction.settings.from_columns, settings.py:23-83: typically the columns a feature selection
    kept).  Same errors: TypeError for a non-string column, ValueError for a name without "__" or an unknown
    calculator."""
    from .plan import CALC
    ignore = set(columns_to_ignore or [])
    kind_to_fc_parameters = {}
    for col in columns:
        if col in ignore:
            continue
        if not isinstance(col, str):
            raise TypeError("Column name {} should be a string or unicode".format(col))
        parts = col.split("__")
        if len(parts) == 1:
            raise ValueError("Splitting of columnname {} resulted in only one part.".format(col))
        kind, feature_name = parts[0], parts[1]
        per_kind = kind_to_fc_parameters.setdefault(kind, {})
        if "TSFX_" + feature_name.upper() not in CALC or feature_name == "const_nan":
            raise ValueError("Unknown feature name {}".format(feature_name))
        config = get_config_from_string(parts)
        if config:
            per_kind.setdefault(feature_name, [])
            if per_kind[feature_name] is None:
                per_kind[feature_name] = []
            per_kind[feature_name].append(config)
        else:
            per_kind[feature_name] = None
    return kind_to_fc_parameters
