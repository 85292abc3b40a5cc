"""Device-side impute (tsfx_impute / TSFX_FLAG_IMPUTE) against the oracle restatement of
tsfresh/utilities/dataframe_functions.py:49-212 and against the golden vectors made from the unmodified reference.
Selection + one add/halve only, so every comparison is bit-exact."""
import os
import warnings

import numpy as np
import pandas as pd
import pytest

from oracle import impute as oi
from tests.helpers import synthetic_series
from tsfresh_b200 import (EfficientFCParameters, extract_features, get_range_values_per_column, impute,
                          impute_dataframe_range, impute_dataframe_zero)
from tsfresh_b200 import _lib
from tsfresh_b200.extraction import get_context

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dirty(seed, rows, cols):
    rng = np.random.default_rng(seed)
    m = rng.standard_normal((rows, cols)) * rng.uniform(0.01, 1e4, cols)
    m[rng.random((rows, cols)) < 0.1] = np.nan
    m[rng.random((rows, cols)) < 0.04] = np.inf
    m[rng.random((rows, cols)) < 0.04] = -np.inf
    if cols > 3:
        m[:, 1] = np.nan
        m[:, 2] = rng.standard_normal(rows)
        m[:, 3] = np.where(np.arange(rows) % 2 == 0, np.inf, -np.inf)
    return m


def test_golden_from_reference():
    g = np.load(os.path.join(G, "impute.npz"))
    ctx = get_context()
    m = g["input"].copy()
    stats = ctx.impute(m, _lib.IMPUTE_RANGE, all_medians=True)
    assert np.array_equal(m, g["imputed"])
    assert np.array_equal(stats, g["stats"])
    z = g["input"].copy()
    ctx.impute(z, _lib.IMPUTE_ZERO)
    assert np.array_equal(z, g["zero"])


@pytest.mark.parametrize("rows,cols", [(1, 1), (2, 5), (33, 32), (1000, 33), (5000, 783), (70000, 7)])
def test_range_matches_oracle(rows, cols):
    m = dirty(rows * 7 + cols, rows, cols)
    want, want_stats = oi.impute(m), oi.range_values(m)
    got = m.copy()
    stats = get_context().impute(got, _lib.IMPUTE_RANGE)
    assert np.array_equal(got, want)
    assert np.array_equal(stats[:2], want_stats[:2])
    has_nan = np.isnan(m).any(axis=0) | ~np.isfinite(m).any(axis=0)
    assert np.array_equal(stats[2][has_nan], want_stats[2][has_nan])      # medians only where one is needed
    assert np.isnan(stats[2][~has_nan]).all()
    only = m.copy()
    st2 = get_context().impute(only, _lib.IMPUTE_STATS)
    assert np.array_equal(only, m, equal_nan=True)                           # statistics only: matrix untouched
    assert np.array_equal(st2, want_stats)


def test_dataframe_functions_mirror_reference_api():
    m = dirty(11, 300, 9)
    cols = ["value__f%d" % i for i in range(m.shape[1])]
    df = pd.DataFrame(m.copy(), columns=cols)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = impute(df)
    assert out is df and np.array_equal(df.to_numpy(), oi.impute(m))
    assert any("did not have any finite values" in str(x.message) for x in w)
    df0 = pd.DataFrame(m.copy(), columns=cols)
    impute_dataframe_zero(df0)
    assert np.array_equal(df0.to_numpy(), oi.impute_zero(m))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cmax, cmin, cmed = get_range_values_per_column(pd.DataFrame(m.copy(), columns=cols))
    st = oi.range_values(m)
    assert [cmin[c] for c in cols] == list(st[0]) and [cmax[c] for c in cols] == list(st[1])
    assert [cmed[c] for c in cols] == list(st[2])
    # train statistics applied to another frame (impute_dataframe_range)
    m2 = dirty(12, 120, 9)
    df2 = pd.DataFrame(m2.copy(), columns=cols)
    impute_dataframe_range(df2, cmax, cmin, cmed)
    assert np.array_equal(df2.to_numpy(), oi.apply_range(m2, st))
    with pytest.raises(ValueError):
        impute_dataframe_range(pd.DataFrame(m2.copy(), columns=cols), {}, cmin, cmed)
    bad = dict(cmax)
    bad[cols[0]] = np.inf
    with pytest.raises(ValueError):
        impute_dataframe_range(pd.DataFrame(m2.copy(), columns=cols), bad, cmin, cmed)
    assert len(impute(pd.DataFrame(columns=cols, dtype=float))) == 0


def test_extract_features_with_device_impute():
    series = list(synthetic_series(5, 40, 64, "normal")) + [np.ones(64, np.float32), np.arange(3, dtype=np.float32)]
    df = pd.DataFrame({"id": np.concatenate([np.full(len(s), i) for i, s in enumerate(series)]),
                       "time": np.concatenate([np.arange(len(s)) for s in series]),
                       "value": np.concatenate(series).astype(np.float32)})
    s = EfficientFCParameters()
    raw = extract_features(df, column_id="id", column_sort="time", default_fc_parameters=s)
    assert not np.isfinite(raw.to_numpy()).all()                              # short / constant series leave NaNs
    X = extract_features(df, column_id="id", column_sort="time", default_fc_parameters=s, impute_function=impute)
    assert list(X.columns) == list(raw.columns) and X.index.equals(raw.index)
    assert np.array_equal(X.to_numpy(), oi.impute(raw.to_numpy()))
    assert np.isfinite(X.to_numpy()).all()
    # two kinds: imputed over the assembled frame
    df2 = df.assign(kind="a")
    dfb = df[df["id"] < 30].assign(kind="b")
    both = pd.concat([df2, dfb], ignore_index=True)
    raw2 = extract_features(both, column_id="id", column_sort="time", column_kind="kind", column_value="value",
                            default_fc_parameters=s)
    X2 = extract_features(both, column_id="id", column_sort="time", column_kind="kind", column_value="value",
                          default_fc_parameters=s, impute_function=impute)
    assert np.array_equal(X2.to_numpy(), oi.impute(raw2.to_numpy()))


def test_impute_flag_through_the_c_abi():
    ctx = get_context()
    from tsfresh_b200.extraction import _device_plan
    from tsfresh_b200.plan import Plan
    dp = _device_plan(ctx, Plan(EfficientFCParameters()))
    v = np.stack(list(synthetic_series(9, 300, 48, "walk"))).astype(np.float32)
    v[7] = 3.0
    raw = dp.extract_dense(v)
    got = dp.extract_dense(v, flags=_lib.FLAG_IMPUTE)
    assert np.array_equal(got, oi.impute(raw))
    begin = (np.arange(len(v)) * 48).astype(np.int64)
    length = np.full(len(v), 48, np.int32)
    got2 = dp.extract_csr(v.ravel(), begin, length, flags=_lib.FLAG_IMPUTE)
    assert np.array_equal(got2, oi.impute(raw))


def test_impute_errors():
    ctx = get_context()
    with pytest.raises(ValueError):
        ctx.impute(np.zeros((3, 3), np.float32))
    with pytest.raises(ValueError):
        ctx.impute(np.zeros((3, 3)), _lib.IMPUTE_GIVEN, col_stats=np.full((3, 3), np.nan))
