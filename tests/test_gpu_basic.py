"""GPU parity, kernel group BASIC (moments / counts / order-dependent streams) through the C ABI."""
import numpy as np
import pytest

from oracle.extract import NOISE_FLOOR
from tests.helpers import gpu_vs_oracle, synthetic_series
from tsfresh_b200.settings import ComprehensiveFCParameters, MinimalFCParameters

pytestmark = pytest.mark.gpu

BASIC = [
    "variance_larger_than_standard_deviation", "has_duplicate_max", "has_duplicate_min", "sum_values", "abs_energy",
    "mean_abs_change", "mean_change", "mean_second_derivative_central", "mean", "length", "standard_deviation",
    "variation_coefficient", "variance", "skewness", "kurtosis", "root_mean_square", "absolute_sum_of_changes",
    "longest_strike_below_mean", "longest_strike_above_mean", "count_above_mean", "count_below_mean",
    "last_location_of_maximum", "first_location_of_maximum", "last_location_of_minimum",
    "first_location_of_minimum", "maximum", "absolute_maximum", "minimum", "benford_correlation",
    "time_reversal_asymmetry_statistic", "c3", "cid_ce", "large_standard_deviation", "autocorrelation",
    "agg_autocorrelation", "partial_autocorrelation", "number_peaks", "binned_entropy", "index_mass_quantile",
    "value_count", "range_count", "linear_trend", "agg_linear_trend", "number_crossing_m",
    "energy_ratio_by_chunks", "ratio_beyond_r_sigma", "count_above", "count_below", "query_similarity_count",
]


@pytest.fixture(scope="module")
def ctx():
    from tsfresh_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def basic_settings():
    full = ComprehensiveFCParameters()
    return {k: full[k] for k in full if k in BASIC}


def _report(bad):
    return "\n".join("row %d %s: gpu=%r oracle=%r" % b for b in bad[:40]) + "\n(%d mismatches)" % len(bad)


@pytest.mark.parametrize("kind", ["normal", "walk", "rounded"])
@pytest.mark.parametrize("length", [256, 100, 1024, 37])
def test_basic_group(ctx, kind, length):
    series = list(synthetic_series(7 + length, 48, length, kind))
    # "rounded" series are multiples of 0.5: sums that are exactly 0 occur, the reference returns rounding noise there
    bad, plan, got, want = gpu_vs_oracle(ctx, basic_settings(), series, atol=NOISE_FLOOR if kind == "rounded" else 0.0)
    assert not bad, _report(bad)


def test_minimal_config1(ctx):
    # BASELINE.json configs[0] shape: Minimal on 1000 x 128
    series = list(synthetic_series(42, 1000, 128))
    s = MinimalFCParameters()
    del s["median"]        # median lives in kernel group SORTED (tests/test_gpu_sorted.py)
    bad, *_ = gpu_vs_oracle(ctx, s, series)
    assert not bad, _report(bad)


def test_short_and_ragged(ctx):
    rng = np.random.default_rng(5)
    series = [rng.standard_normal(n).astype(np.float32) for n in (1, 2, 3, 4, 5, 8, 20, 31, 32, 33, 63, 64, 65, 200, 1500)]
    series += [np.zeros(10, np.float32), np.ones(7, np.float32), np.array([1, 1, 2, 2, 3, 3, 3], np.float32),
               np.array([5.0], np.float32), np.array([-1, 1] * 20, np.float32)]
    bad, *_ = gpu_vs_oracle(ctx, basic_settings(), series, atol=NOISE_FLOOR)        # constant / alternating / 1-sample series
    assert not bad, _report(bad)
