"""The oracle against the committed golden vectors (made from the unmodified reference by oracle/make_golden.py)
and against the reference's own known-answer tests
(/root/reference/tests/units/feature_extraction/test_feature_calculations.py, lines cited per case).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import calculators as oc
from oracle.extract import compare, oracle_rows
from tsfresh_b200.plan import Plan
from tsfresh_b200.settings import ComprehensiveFCParameters, EfficientFCParameters, MinimalFCParameters

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_series():
    z = np.load(os.path.join(G, "comprehensive.npz"))
    return [z["values"][b:b + n] for b, n in zip(z["begin"], z["length"])], z["reference"]


def test_oracle_reproduces_reference_golden_matrix():
    series, ref = golden_series()
    s = ComprehensiveFCParameters()
    got = oracle_rows([x.astype(np.float64) for x in series], s)
    assert got.shape == ref.shape == (12, 783)
    assert not compare(got, ref, Plan(s).suffixes, rtol=1e-9, atol=1e-12)


def test_column_names_match_reference():
    cols = json.load(open(os.path.join(G, "columns.json")))
    assert Plan(ComprehensiveFCParameters()).suffixes == cols["comprehensive"]
    assert Plan(EfficientFCParameters()).suffixes == cols["efficient"]
    assert Plan(MinimalFCParameters()).suffixes == cols["minimal"]
    assert len(cols["comprehensive"]) == 783 and len(cols["efficient"]) == 777 and len(cols["minimal"]) == 10
    assert len(cols["restated_columns"]) == 87


A = np.asarray


@pytest.mark.parametrize("name,x,params,want", [
    # (calculator, input, params, expected) -- test_feature_calculations.py line of the assertion
    ("abs_energy", [1, 1, 1], {}, 3),                                                   # :519
    ("abs_energy", [-1, 2, -3], {}, 14),                                                # :521
    ("cid_ce", [1, 1, 1], {"normalize": True}, 0),                                      # :527
    ("cid_ce", [-4.33, -1.33, 2.67], {"normalize": False}, 5),                          # :535
    ("mean_abs_change", [-2, 2, 5], {}, 3.5),                                           # :640
    ("mean_change", [-2, 2, 5], {}, 3.5),                                               # :649
    ("mean_second_derivative_central", [1, 3, 7], {}, 1),                               # :665 (2/2)
    ("variance_larger_than_standard_deviation", [-1, -1, 1, 1, 2], {}, True),           # :158
    ("large_standard_deviation", [-1, -1, 1, 1], {"r": 0.25}, True),                    # :169
    ("has_duplicate_max", [2.1, 0, 0, 2.1, 1.1], {}, True),                             # :203
    ("has_duplicate_min", [-2.1, 0, 0, -2.1, 1.1], {}, True),                           # :213
    ("has_duplicate", [2.1, 0, 0, 2.1, 1.1], {}, True),                                 # :222
    ("median", [1, 1, 2, 2], {}, 1.5),                                                  # :677
    ("skewness", [1, 2, 2, 3], {}, 0.0),                                                # :721
    ("longest_strike_below_mean", [1, 2, 1, 1, 1, 2, 2, 2], {}, 3),                     # :757
    ("longest_strike_above_mean", [1, 2, 1, 2, 1, 2, 2, 1], {}, 2),                     # :772
    ("count_above_mean", [1, 2, 1, 2, 1, 2], {}, 3),                                    # :786
    ("count_below_mean", [1, 2, 1, 2, 1, 2], {}, 3),                                    # :792
    ("last_location_of_maximum", [1, 2, 1, 2, 1], {}, 0.8),                             # :799
    ("first_location_of_maximum", [1, 2, 1, 2, 1], {}, 0.2),                            # :806
    ("last_location_of_minimum", [1, 2, 1, 2, 1], {}, 1.0),                             # :821
    ("first_location_of_minimum", [1, 2, 1, 2, 1], {}, 0.0),                            # :828
    ("ratio_beyond_r_sigma", [0, 1] * 10 + [10, 20, -30], {"r": 1}, 3.0 / 23),          # :151 (3 outliers of 23)
    ("number_peaks", [0, 1, 2, 1, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1], {"n": 1}, 2),          # :972
    ("number_peaks", [0, 1, 2, 1, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1], {"n": 3}, 1),          # :974
    ("number_crossing_m", [10, -10, 10, -10], {"m": 0}, 3),                             # :1635
    ("value_count", [1] * 10, {"value": 1}, 10),                                        # :1666
    ("range_count", list(range(10)), {"min": 1, "max": 1}, 0),                          # :1684
    ("lempel_ziv_complexity", [1, 1, 1], {"bins": 2}, 2.0 / 3),                         # :437
    ("lempel_ziv_complexity", [1, 1, 1, 1, 1, 1, 1], {"bins": 2}, 0.4285714285),      # :439-441
    ("lempel_ziv_complexity", [1, 1, 1, 2, 1, 1, 1], {"bins": 2}, 0.5714285714),      # :442-444
    ("lempel_ziv_complexity", [-1, 4.3, 5, 1, -4.5, 1, 5, 7, -3.4, 6], {"bins": 10}, 0.8),   # :446-448
    ("permutation_entropy", [4, 7, 9, 10, 6, 11, 3], {"dimension": 3, "tau": 1}, 1.054920167),   # :489-495
    ("count_above", [1] * 10, {"t": 1}, 1.0),                                           # :1985
    ("count_below", [1] * 10, {"t": 1}, 1.0),                                           # :1999
    ("mean_n_absolute_max", [12, 3], {"number_of_maxima": 1}, 12),                      # :1300
    ("sum_of_reoccurring_values", [1, 1, 2, 3, 4, 4], {}, 5),                           # :877 (1 + 4)
    ("sum_of_reoccurring_data_points", [1, 1, 2, 3, 4, 4], {}, 10),                     # :883
    ("percentage_of_reoccurring_values_to_all_values", [1, 1, 2, 3, 4], {}, 0.25),      # :840
    ("percentage_of_reoccurring_datapoints_to_all_datapoints", [1, 1, 2, 3, 4], {}, 0.4),   # :858
    ("ratio_value_number_to_time_series_length", [1, 1, 2, 3, 4], {}, 0.8),             # :894
])
def test_known_answers(name, x, params, want):
    got = oc.SIMPLE[name](A(x, dtype=float), **params)
    assert float(got) == pytest.approx(float(want), rel=1e-8, abs=1e-12)


def test_known_answers_combiners():
    # agg_autocorrelation on range(10): 0.77777777 (mean), -0.64983164983165 (var closed form) -- :262-280, places=4
    x = np.arange(10, dtype=float)
    got = oc.COMBINER["agg_autocorrelation"](x, [{"f_agg": "mean", "maxlag": 1}, {"f_agg": "mean", "maxlag": 10}])
    assert got[0] == pytest.approx(0.77777777, abs=1e-4) and got[1] == pytest.approx(-0.64983164983165, abs=1e-4)
    x = A([1.0, 2.0, -3.0])                       # closed form of :246-249
    want = 1 / np.var(x) * (((1 * 2 + 2 * (-3)) / 2 + (1 * -3)) / 2)
    for agg in ("mean", "median"):
        assert oc.COMBINER["agg_autocorrelation"](x, [{"f_agg": agg, "maxlag": 10}])[0] == pytest.approx(want, abs=1e-4)
    # partial_autocorrelation of an alternating series: lag 1 = -1 (:300-304)
    got = oc.COMBINER["partial_autocorrelation"](A([1.0, -1.0] * 50), [{"lag": 1}])
    assert got[0] == pytest.approx(-1.0, abs=1e-4)
    # ar_coefficient recovers (1, 2.5) for x_t = 2.5 x_{t-1} + 1 (:1084-1099)
    x = [1.0, 3.5]
    for _ in range(30):
        x.append(2.5 * x[-1] + 1)
    got = oc.COMBINER["ar_coefficient"](A(x), [{"k": 1, "coeff": 0}, {"k": 1, "coeff": 1}])
    assert got[0] == pytest.approx(1.0, rel=1e-3) and got[1] == pytest.approx(2.5, rel=1e-6)
    # fft_aggregated on [1]*10 + [0]*10-like shapes are covered by the golden matrix; spot value (:907-913)
    x = np.arange(10, dtype=float)
    got = oc.COMBINER["fft_aggregated"](x, [{"aggtype": s} for s in ("centroid", "variance", "skew", "kurtosis")])
    np.testing.assert_allclose(got, [1.135, 2.368, 1.249, 3.643], atol=1e-3)
    # energy_ratio_by_chunks on range(1, 13) in 3 chunks (:1751-1760): 0.0468, 0.2677, 0.6854
    got = oc.COMBINER["energy_ratio_by_chunks"](np.arange(1, 13, dtype=float), [{"num_segments": 3, "segment_focus": i} for i in range(3)])
    np.testing.assert_allclose(got, [30 / 650, 174 / 650, 446 / 650], rtol=1e-12)
    # linear_trend on a straight line (:1010-1020)
    got = oc.COMBINER["linear_trend"](np.arange(20, dtype=float) * 2 + 1, [{"attr": a} for a in ("pvalue", "rvalue", "intercept", "slope", "stderr")])
    np.testing.assert_allclose(got, [0, 1, 1, 2, 0], atol=1e-9)


def test_fixture80_end_to_end_values():
    """tests/units/feature_extraction/test_extraction.py:40-55: exact values on the reference's 80-row fixture."""
    z = np.load(os.path.join(G, "fixture80.npz"))
    cols = list(z["columns"])
    ref = z["reference"]
    assert list(z["index"]) == [10, 500]
    for name, want in (("a__maximum", [71, 77]), ("a__sum_values", [691, 1017]), ("a__abs_energy", [32211, 63167]),
                       ("b__sum_values", [757, 695]), ("b__minimum", [3, 1]), ("b__abs_energy", [36619, 35483]),
                       ("b__mean", [37.85, 34.75]), ("b__median", [39.5, 28.0])):
        np.testing.assert_allclose(ref[:, cols.index(name)], want)


def test_oracle_impute_reproduces_reference_golden():
    """tests/golden/impute.npz (oracle/make_golden_impute.py, unmodified reference) -- bit-exact."""
    from oracle import impute as oi
    g = np.load(os.path.join(G, "impute.npz"))
    m = g["input"]
    assert np.array_equal(oi.range_values(m), g["stats"])
    assert np.array_equal(oi.impute(m), g["imputed"])
    assert np.array_equal(oi.impute_zero(m), g["zero"])
    assert np.isfinite(g["imputed"]).all()
