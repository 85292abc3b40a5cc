"""TEST INFRASTRUCTURE: the oracle's version of the per-series loop (_do_extraction_on_chunk,
extraction.py:308-386): for every series, for every settings entry, evaluate the restated calculator.
Column order/names come from tsfresh_b200.plan (checked against the reference in tests/golden)."""
import warnings

import numpy as np

from . import calculators

# columns whose reference value is boolean / integer / an exact ratio k/n: compared bit-exactly
EXACT_CALCULATORS = {
    "variance_larger_than_standard_deviation", "ratio_beyond_r_sigma", "large_standard_deviation",
    "symmetry_looking", "has_duplicate_max", "has_duplicate_min", "has_duplicate", "length",
    "longest_strike_below_mean", "longest_strike_above_mean", "count_above_mean", "count_below_mean",
    "last_location_of_maximum", "first_location_of_maximum", "last_location_of_minimum",
    "first_location_of_minimum", "percentage_of_reoccurring_values_to_all_values",
    "percentage_of_reoccurring_datapoints_to_all_datapoints", "ratio_value_number_to_time_series_length",
    "number_peaks", "index_mass_quantile", "number_cwt_peaks", "number_crossing_m", "value_count",
    "range_count", "count_above", "count_below", "lempel_ziv_complexity",
}


def is_exact_column(suffix):
    name = suffix.split("__")[0]
    return name in EXACT_CALCULATORS or suffix.startswith('augmented_dickey_fuller__attr_"usedlag"')


def oracle_rows(series, fc_parameters, skip=("linear_trend_timewise",), times=None):
    """series: iterable of 1-D arrays.  Returns float64 matrix [n_series x n_columns].
    times: per-series int64 nanosecond timestamps (the frame's DatetimeIndex); with them linear_trend_timewise is
    evaluated (extraction.py:349-361 skips it, with a warning, when the index is not a DatetimeIndex)."""
    rows = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            for k, x in enumerate(series):
                x = np.asarray(x, dtype=np.float64)
                row = []
                for name, params in fc_parameters.items():
                    if name == "linear_trend_timewise" and times is not None:
                        row.extend(float(v) for v in calculators.linear_trend_timewise(x, times[k], params))
                        continue
                    if name in skip:
                        continue
                    row.extend(calculators.evaluate(name, x, params))
                rows.append(row)
    return np.asarray(rows, dtype=np.float64)


NOISE_FLOOR = 1e-9


def compare(got, want, suffixes, rtol=1e-5, atol=0.0):
    """Returns a list of (row, column-suffix, got, want) mismatches under the parity definition of
    SURVEY.md section 8c: exact columns ==, float columns isclose(rtol, atol=0) with NaN == NaN, inf == inf.

    The default is the strict definition (atol = 0).  Measured on the B200 at the BASELINE shapes (2 048 x 256 Efficient,
    2 048 x 256 + 512 random walks Comprehensive, 512 x 1024, tests/test_gpu_shapes.py): NO column needs an absolute
    floor.  `atol=NOISE_FLOOR` is passed only by the tests that feed DEGENERATE series (constant, two-valued, 1..5
    samples, exact integers): there the mathematically exact answer is 0 (FFT bins of a constant series, the slope of
    a flat aggregate, ...) and the reference itself returns rounding noise of order 1e-16..1e-13 whose digits no
    other summation order reproduces.  The floor is never applied to the exact (bool / count / ratio) columns.  The
    phase of an FFT bin whose magnitude is below the floor (or exactly 0) is not compared."""
    bad = []
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    index = {s: i for i, s in enumerate(suffixes)}
    for c, suf in enumerate(suffixes):
        g, w = got[:, c], want[:, c]
        both_nan = np.isnan(g) & np.isnan(w)
        if is_exact_column(suf):
            ok = (g == w) | both_nan
        else:
            with np.errstate(all="ignore"):
                ok = np.isclose(g, w, rtol=rtol, atol=atol) | both_nan | ((g == w))
        if suf.startswith('fft_coefficient__attr_"angle"'):
            mag = index.get(suf.replace('"angle"', '"abs"'))
            if mag is not None:
                ok = ok | (np.abs(want[:, mag]) < max(atol, 1e-12))
        if 'attr_"stderr"' in suf:
            # scipy.stats.linregress on exactly two points: stderr = sqrt((1-r^2)*ssym/ssxm/0) is NaN when
            # r rounds to +-1 and inf when it rounds to 0.999..; the reference itself is rounding-chaotic
            # here (DESIGN.md "known reference instabilities"), so NaN and inf are treated as the same answer.
            ok = ok | ((np.isnan(g) | np.isinf(g)) & (np.isnan(w) | np.isinf(w)))
        for r in np.nonzero(~ok)[0]:
            bad.append((int(r), suf, float(g[r]), float(w[r])))
    return bad
