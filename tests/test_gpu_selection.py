"""Feature selection on the GPU (tsfresh_b200.feature_selection, csrc/tsfx_select.cu) against golden relevance tables of the
UNMODIFIED reference (tests/golden/selection.npz from oracle/make_golden_selection.py: relevance.py:31-322 with scipy's
mannwhitneyu / ks_2samp / fisher_exact / kendalltau) -- binary, multiclass and regression targets (with and without ties),
real / tied / binary / constant features."""
import os
import warnings

import numpy as np
import pandas as pd
import pytest

from tsfresh_b200.feature_selection import calculate_relevance_table, select_features

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "selection.npz"), allow_pickle=True)

CASES = {"yr": ("yr", {}), "yt": ("yt", {"hypotheses_independent": True}), "y2": ("y2", {}), "y2smir": ("y2", {"test_for_binary_target_real_feature": "smir"}),
         "y2indep": ("y2", {"hypotheses_independent": True, "fdr_level": 0.2}),
         "y3": ("y3", {"multiclass": True, "n_significant": 2})}


@pytest.mark.parametrize("size", ["small", "medium", "large"])
@pytest.mark.parametrize("case", list(CASES))
def test_relevance_table_matches_the_reference(size, case):
    ycol, kw = CASES[case]
    X = pd.DataFrame(Z[size + "_X"], columns=list(Z["columns"]))
    y = pd.Series(Z["%s_%s" % (size, ycol)], index=X.index)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t = calculate_relevance_table(X, y, ml_task="regression" if case in ("yr", "yt") else "classification", **kw)
    key = "%s_%s" % (size, case)
    assert list(t.columns) == list(Z[key + "_columns"])
    assert sorted(t.index) == sorted(Z[key + "_index"])
    ref = {c: pd.Series(Z[key + "_col_" + c], index=Z[key + "_index"]) for c in Z[key + "_columns"]}
    for c in t.columns:
        got, want = t[c], ref[c].reindex(t.index)
        if c.startswith("p_value"):
            np.testing.assert_allclose(got.to_numpy(dtype=float), want.to_numpy(dtype=float), rtol=1e-9, atol=1e-300, equal_nan=True)
        elif c.startswith("relevant") or c == "n_significant":
            assert list(got.astype(float)) == list(want.astype(float)), c
        else:
            assert list(got.astype(str)) == list(want.astype(str)), c
    if "p_value" in t.columns:                   # sorted by p-value, constant features (NaN) last -- as the reference
        p = t["p_value"].to_numpy(dtype=float)
        finite = p[~np.isnan(p)]
        assert (np.diff(finite) >= 0).all() and not np.isnan(p[:len(finite)]).any()


def test_select_features_and_errors():
    X = pd.DataFrame(Z["large_X"], columns=list(Z["columns"]))
    y = Z["large_y2"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sel = select_features(X, y)
    rel = pd.Series(Z["large_y2_col_relevant"], index=Z["large_y2_index"]).astype(bool)
    assert sorted(sel.columns) == sorted(rel.index[rel.to_numpy()])
    assert "constant" not in sel.columns and len(sel) == len(X)
    with pytest.raises(ValueError, match="NaN"):
        yn = np.random.default_rng(0).standard_normal(len(X))
        yn[3] = np.nan
        calculate_relevance_table(X, pd.Series(yn), ml_task="regression")
    Xn = X.copy()
    Xn.iloc[5, 3] = np.nan
    with pytest.raises(ValueError, match="NaN"):
        calculate_relevance_table(Xn, pd.Series(y), ml_task="classification")
    with pytest.raises(AssertionError):
        select_features(X, np.zeros(len(X), dtype=int))               # one class only
