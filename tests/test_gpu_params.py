"""GPU parity on NON-DEFAULT parameter values: the settings dict is user input, so every calculator is exercised
away from the ComprehensiveFCParameters grid (large lags, many bins, extreme quantiles, other autolag modes ...)."""
import numpy as np
import pytest

from tests.helpers import gpu_vs_oracle, synthetic_series

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from tsfresh_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def _report(bad):
    return "\n".join("row %d %s: gpu=%r oracle=%r" % b for b in bad[:40]) + "\n(%d mismatches)" % len(bad)


SETTINGS = {
    "basic": {
        "ratio_beyond_r_sigma": [{"r": r} for r in (0.0, 0.1, 0.25, 3.3, 100)],
        "large_standard_deviation": [{"r": r} for r in (0.0, 0.01, 0.33, 1.0, 2.0)],
        "cid_ce": [{"normalize": True}, {"normalize": False}],
        "autocorrelation": [{"lag": l} for l in (0, 1, 17, 63, 100, 199, 200, 201, 500)],
        "agg_autocorrelation": [{"f_agg": f, "maxlag": m} for f in ("mean", "median", "var", "std") for m in (1, 3, 64, 150, 400)],
        "partial_autocorrelation": [{"lag": l} for l in (0, 1, 2, 5, 20, 45)],
        "number_peaks": [{"n": n} for n in (1, 2, 7, 40, 99, 100, 300)],
        "binned_entropy": [{"max_bins": b} for b in (1, 2, 7, 100, 1000)],
        "index_mass_quantile": [{"q": q} for q in (0.0, 0.001, 0.5, 0.999, 1.0)],
        "value_count": [{"value": v} for v in (0.5, -0.5, 2, float("nan"))],
        "range_count": [{"min": -0.5, "max": 0.5}, {"min": 1, "max": -1}, {"min": -1e30, "max": 1e30}],
        "number_crossing_m": [{"m": m} for m in (0.3, -2.5, 100)],
        "count_above": [{"t": t} for t in (-1, 0.5, 10)],
        "count_below": [{"t": t} for t in (-1, 0.5, 10)],
        "time_reversal_asymmetry_statistic": [{"lag": l} for l in (1, 10, 99, 100, 150)],
        "c3": [{"lag": l} for l in (1, 10, 99, 100, 150)],
        "energy_ratio_by_chunks": [{"num_segments": s, "segment_focus": f} for s, f in ((1, 0), (3, 2), (7, 0), (7, 6), (64, 63))],
        "linear_trend": [{"attr": a} for a in ("slope", "pvalue")],
        "agg_linear_trend": [{"attr": a, "chunk_len": c, "f_agg": f} for a in ("pvalue", "slope", "stderr")
                             for c in (2, 3, 7, 66, 199, 1000) for f in ("max", "min", "mean", "var", "std")],
    },
    "sorted": {
        "symmetry_looking": [{"r": r} for r in (0.0, 0.013, 0.5, 1.0)],
        "quantile": [{"q": q} for q in (0.0, 0.01, 0.25, 0.5, 0.75, 0.999, 1.0)],
        "mean_n_absolute_max": [{"number_of_maxima": k} for k in (1, 2, 50, 199, 200, 1000)],
        "change_quantiles": [{"ql": ql, "qh": qh, "isabs": b, "f_agg": f} for ql, qh in ((0.0, 1.0), (0.1, 0.11), (0.3, 0.3), (0.9, 0.2), (0.45, 0.55))
                             for b in (False, True) for f in ("mean", "var", "std")],
        "friedrich_coefficients": [{"coeff": c, "m": 3, "r": r} for r in (2, 5, 10, 30, 100) for c in (0, 3, 4)],
        "max_langevin_fixed_point": [{"m": 3, "r": r} for r in (5, 30, 60)],
    },
    "spectral": {
        "fft_coefficient": [{"coeff": k, "attr": a} for a in ("real", "imag", "abs", "angle") for k in (0, 1, 50, 100, 101, 127, 128, 129, 500)],
        "fft_aggregated": [{"aggtype": s} for s in ("centroid", "variance", "skew", "kurtosis")],
        "spkt_welch_density": [{"coeff": c} for c in (0, 1, 64, 100, 128, 129, 1000)],
        "fourier_entropy": [{"bins": b} for b in (1, 4, 17, 500)],
        "cwt_coefficients": [{"widths": w, "coeff": c, "w": s} for w in ((1, 3), (2, 7, 30), (2, 5, 10, 20)) for s in w for c in (0, 7, 100, 199, 200, 300)],
    },
    "la": {
        "ar_coefficient": [{"coeff": c, "k": k} for k in (1, 2, 5, 20, 32) for c in (0, 1, k, k + 1)],
        "augmented_dickey_fuller": [{"attr": a, "autolag": al} for al in ("AIC", "BIC", None) for a in ("teststat", "pvalue", "usedlag", "nonsense")],
    },
    "entropy_seq": {
        "approximate_entropy": [{"m": 2, "r": r} for r in (0.0, 0.05, 0.2, 1.0, 1.5, 3.0, 10.0)],
        "sample_entropy": None,
        "lempel_ziv_complexity": [{"bins": b} for b in (1, 2, 4, 16, 64, 1000, 4, 7, 9)],
        "permutation_entropy": [{"tau": t, "dimension": d} for t in (1, 2, 5) for d in (2, 3, 6, 7, 8)],
        "number_cwt_peaks": [{"n": n} for n in (1, 2, 3, 8, 16)],
    },
}


@pytest.mark.parametrize("group", list(SETTINGS))
@pytest.mark.parametrize("kind,length", [("normal", 200), ("walk", 256), ("normal", 61)])
def test_non_default_parameters(ctx, group, kind, length):
    series = list(synthetic_series(500 + length, 24, length, kind))
    bad, *_ = gpu_vs_oracle(ctx, SETTINGS[group], series)
    assert not bad, _report(bad)


def test_very_wide_plan(ctx):
    """thousands of columns in one plan (the assemble pass sizes its shared-memory row buffer accordingly)"""
    settings = {"fft_coefficient": [{"coeff": k, "attr": a} for a in ("real", "imag", "abs", "angle") for k in range(1200)],
                "quantile": [{"q": q / 1000.0} for q in range(1001)]}
    series = list(synthetic_series(3, 6, 300))
    bad, plan, got, want = gpu_vs_oracle(ctx, settings, series)
    assert plan.n_cols == 4800 + 1001
    assert not bad, _report(bad)
