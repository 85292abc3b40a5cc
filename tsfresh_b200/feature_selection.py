"""Feature selection on the feature matrix -- the step after the hot path.

Mirrors tsfresh/feature_selection/relevance.py:31-322 (`calculate_relevance_table`, `combine_relevance_tables`,
`infer_ml_task`, `get_feature_type`) and selection.py:17-200 (`select_features`) for CLASSIFICATION targets (binary and
multiclass): same arguments, same relevance table (index / columns / order / dtypes), same decisions.

The reference runs one scipy test per feature column in a Python loop (783 calls per class label, each sorting the column on
the CPU).  Here every column of the matrix is sorted ONCE on the GPU and the sufficient statistics of all tests come out of
one pass over the sorted column (tsfx_select_classification, csrc/tsfx_select.cu):

  real feature,   target binary:  Mann-Whitney U + tie term  (significance_tests.py:84-132, scipy.stats.mannwhitneyu)
                                  or the two-sample Kolmogorov-Smirnov statistic ('smir', scipy.stats.ks_2samp)
  binary feature, target binary:  the 2 x 2 contingency table  (significance_tests.py:43-81, scipy.stats.fisher_exact)

What is left for the host is O(n_features): the p-value of each statistic (the closed forms scipy itself evaluates: normal
tail with continuity and tie correction / exact U distribution for tiny untied samples, hypergeometric tail, Kolmogorov
distributions) and the Benjamini-Hochberg / Benjamini-Yekutieli step-up decision (statsmodels.stats.multitest.multipletests
in the reference, relevance.py:347-351).

REGRESSION targets (relevance.py:282-296) take tsfx_select_regression: real features need Kendall's tau -- the rows are
brought into (x, rank(y)) order by two stable device sorts, the discordant pairs are the strict inversions of the rank
sequence (bottom-up merge, one kernel per level), tie sums come from the runs; binary features need the two-sample KS of
the target between their two groups (one pass over the y-sorted rows).
"""
import math
import warnings
from functools import reduce

import numpy as np
import pandas as pd

from . import _lib
from .extraction import N_PROCESSES, get_context

TEST_FOR_BINARY_TARGET_BINARY_FEATURE = "fisher"        # tsfresh/defaults.py:15-20
TEST_FOR_BINARY_TARGET_REAL_FEATURE = "mann"
TEST_FOR_REAL_TARGET_BINARY_FEATURE = "mann"
TEST_FOR_REAL_TARGET_REAL_FEATURE = "kendall"
FDR_LEVEL = 0.05
HYPOTHESES_INDEPENDENT = False


# ------------------------------------------------------------------------------------------ p-values from statistics
def _norm_sf(z):
    return 0.5 * math.erfc(z / math.sqrt(2.0))


def _mwu_exact_sf(k, n1, n2):
    """P(U >= k) for the Mann-Whitney U of a sample of n1 against n2 without ties: coefficients of the Gaussian binomial
    prod_{i=1..n1} (1 - q^(n2+i)) / (1 - q^i) (what scipy's _MWU tabulates)."""
    n1, n2 = int(n1), int(n2)
    if n1 > n2:
        n1, n2 = n2, n1
    size = n1 * n2 + 1
    c = np.zeros(size, dtype=np.float64)
    c[0] = 1.0
    for i in range(1, n1 + 1):
        num = n2 + i
        if num < size:                               # multiply by (1 - q^num)
            c[num:] -= c[:-num].copy()
        for u in range(i, size):                     # divide by (1 - q^i)
            c[u] += c[u - i]
    total = c.sum()
    k = max(0, int(k))
    return float(c[k:].sum() / total) if k < size else 0.0


def mannwhitneyu_pvalue(U1, n1, n2, tie_term, distinct):
    """scipy.stats.mannwhitneyu(x_y1, x_y0, use_continuity=True, alternative="two-sided") p-value from its statistics
    (method="auto": exact for a sample of <= 8 without ties, else the normal approximation with tie correction)."""
    if n1 == 0 or n2 == 0:
        raise ValueError("`x` and `y` must be of nonzero size.")
    n = n1 + n2
    ties = distinct < n
    if (n1 <= 8 or n2 <= 8) and not ties:
        U = max(U1, n1 * n2 - U1)
        return float(min(1.0, max(0.0, 2.0 * _mwu_exact_sf(int(U), n1, n2))))
    mu = n1 * n2 / 2.0
    with np.errstate(all="ignore"):
        s = math.sqrt(n1 * n2 / 12.0 * ((n + 1) - tie_term / (n * (n - 1.0))))
        num = U1 - mu
        num -= 0.5 * np.sign(num)
        z = num / s if s != 0 else (math.nan if num == 0 else math.copysign(math.inf, num))
    if z != z:
        return math.nan
    return float(min(1.0, max(0.0, 2.0 * _norm_sf(abs(z)))))


def ks_2samp_pvalue(d, n1, n2):
    """scipy.stats.ks_2samp(..., alternative="two-sided", method="auto") p-value from the statistic: exact for samples of
    at most 10 000 rows (scipy's lattice-path count), else the asymptotic Kolmogorov distribution."""
    from scipy import stats
    n1, n2 = int(n1), int(n2)
    g = math.gcd(n1, n2)
    if max(n1, n2) <= 10000:
        try:
            from scipy.stats._stats_py import _attempt_exact_2kssamp
            ok, d2, prob = _attempt_exact_2kssamp(n1, n2, g, d, "two-sided")
            if ok:
                return float(np.clip(prob, 0, 1))
        except Exception:
            pass
    m, n = sorted([float(n1), float(n2)], reverse=True)
    en = m * n / (m + n)
    return float(np.clip(stats.distributions.kstwo.sf(d, np.round(en)), 0, 1))


def kendall_pvalue(n, dis, xtie, x0, x1, ntie, ytie, y0, y1):
    """scipy.stats.kendalltau(x, y, method="asymptotic").pvalue (tau-b) from its sufficient statistics: discordant pairs,
    the tie sums sum t(t-1)/2, sum t(t-1)(t-2), sum t(t-1)(2t+5) of x and y and the joint ties (scipy _stats_py.py)."""
    tot = n * (n - 1) // 2
    if xtie == tot or ytie == tot:
        return math.nan
    con_minus_dis = tot - xtie - ytie + ntie - 2 * dis
    m = n * (n - 1.0)
    var = (m * (2 * n + 5) - x1 - y1) / 18 + (2 * xtie * ytie) / m + x0 * y0 / (9 * m * (n - 2))
    z = con_minus_dis / math.sqrt(var)
    return math.erfc(abs(z) / math.sqrt(2.0))


def fisher_pvalue(n_y1_x1, n_y1_x0, n_y0_x1, n_y0_x0):
    from scipy import stats
    table = np.array([[n_y1_x1, n_y1_x0], [n_y0_x1, n_y0_x0]], dtype=np.int64)
    return float(stats.fisher_exact(table, alternative="two-sided")[1])


def benjamini_reject(pvals, alpha, independent):
    """statsmodels.stats.multitest.multipletests(pvals, alpha, "fdr_bh" | "fdr_by")[0] (relevance.py:347-351): the
    Benjamini-Hochberg step-up procedure; for arbitrary dependence the thresholds are divided by sum(1 / i)."""
    p = np.asarray(pvals, dtype=np.float64)
    m = len(p)
    if m == 0:
        return np.zeros(0, dtype=bool)
    order = np.argsort(p)
    ps = p[order]
    factor = np.arange(1, m + 1) / float(m)
    if not independent:
        factor = factor / np.sum(1.0 / np.arange(1, m + 1))
    reject = ps <= factor * alpha
    if reject.any():
        reject[:np.max(np.nonzero(reject)[0]) + 1] = True
    out = np.empty(m, dtype=bool)
    out[order] = reject
    return out


# ------------------------------------------------------------------------------------------ reference API
def infer_ml_task(y):
    """relevance.py:354-375"""
    if y.dtype.kind in np.typecodes["AllInteger"] or y.dtype == object or isinstance(y.dtype, pd.StringDtype):
        return "classification"
    return "regression"


def get_feature_type(feature_column):
    """relevance.py:398-414 (one column; calculate_relevance_table types all columns on the device)"""
    n_unique_values = len(set(np.asarray(feature_column)))
    return "constant" if n_unique_values == 1 else ("binary" if n_unique_values == 2 else "real")


def combine_relevance_tables(relevance_tables):
    """relevance.py:378-395"""
    def _combine(a, b):
        a.relevant |= b.relevant
        a.p_value = a.p_value.combine(b.p_value, min, 1)
        return a
    return reduce(_combine, relevance_tables)


def _table_for_label(features, types, stats, test_real, fdr_level, hypotheses_independent):
    """relevance.py:325-351 for one implicit binary target, from the device statistics [n_features, 8]"""
    real = [i for i, t in enumerate(types) if t == "real"]
    binary = [i for i, t in enumerate(types) if t == "binary"]
    p = {}
    for i in real:
        s = stats[i]
        if test_real == "mann":
            p[i] = mannwhitneyu_pvalue(s[3], s[1], s[2], s[4], s[6])
        elif test_real == "smir":
            p[i] = ks_2samp_pvalue(s[5], s[1], s[2])
        else:
            raise ValueError("Please use a valid entry for test_for_binary_target_real_feature. "
                             "Valid entries are 'mann' and 'smir'.")
    for i in binary:
        s = stats[i]
        p[i] = fisher_pvalue(int(s[3]), int(s[4]), int(s[5]), int(s[6]))
    idx = real + binary                                   # pd.concat([table_real, table_binary])
    table = pd.DataFrame({"feature": [features[i] for i in idx], "type": [types[i] for i in idx],
                          "p_value": [p[i] for i in idx]}, index=pd.Index([features[i] for i in idx], name="feature"))
    table["relevant"] = benjamini_reject(table.p_value.to_numpy(), fdr_level, hypotheses_independent)
    return table.sort_values("p_value")


def calculate_relevance_table(X, y, ml_task="auto", multiclass=False, n_significant=1, n_jobs=N_PROCESSES, show_warnings=False,
                              chunksize=None, test_for_binary_target_binary_feature=TEST_FOR_BINARY_TARGET_BINARY_FEATURE,
                              test_for_binary_target_real_feature=TEST_FOR_BINARY_TARGET_REAL_FEATURE,
                              test_for_real_target_binary_feature=TEST_FOR_REAL_TARGET_BINARY_FEATURE,
                              test_for_real_target_real_feature=TEST_FOR_REAL_TARGET_REAL_FEATURE, fdr_level=FDR_LEVEL,
                              hypotheses_independent=HYPOTHESES_INDEPENDENT, device=None):
    """GPU implementation of tsfresh.feature_selection.relevance.calculate_relevance_table (relevance.py:31-322)."""
    y = y.sort_index()
    X = X.sort_index()
    assert list(y.index) == list(X.index), "The index of X and y need to be the same"
    if ml_task not in ["auto", "classification", "regression"]:
        raise ValueError("ml_task must be one of: 'auto', 'classification', 'regression'")
    elif ml_task == "auto":
        ml_task = infer_ml_task(y)
    if multiclass:
        assert ml_task == "classification", "ml_task must be classification for multiclass problem"
        assert len(y.unique()) >= n_significant, "n_significant must not exceed the total number of classes"
        if len(y.unique()) <= 2:
            warnings.warn("Two or fewer classes, binary feature selection will be used (multiclass = False)")
            multiclass = False

    with warnings.catch_warnings():
        warnings.simplefilter("default" if show_warnings else "ignore")
        features = list(X.columns)
        M = np.ascontiguousarray(X.to_numpy(dtype=np.float64))
        ctx = get_context(device)
        try:
            if ml_task == "classification":
                labels = list(y.unique())                # order of appearance, as the reference's loop (relevance.py:248)
                codes = pd.Categorical(y, categories=labels).codes.astype(np.int32)
                stats = ctx.select_classification(M, codes, len(labels))      # [n_labels, n_features, 8]
                type_col = stats[0, :, 0]
            else:
                yv = np.ascontiguousarray(y.to_numpy(dtype=np.float64))
                if np.isnan(yv).any():
                    raise ValueError("Target y contains NaN values")
                rstats, ytail = ctx.select_regression(M, yv)                  # [n_features, 8], (ytie, y0, y1, n)
                type_col = rstats[:, 0]
        except ValueError as e:
            if "NaN" in str(e) and "Target" not in str(e):
                raise ValueError("Feature {} contains NaN values".format("matrix")) from None
            raise
        type_names = {0: "constant", 1: "binary", 2: "real"}
        types = [type_names[int(t)] for t in type_col]
        const = [i for i, t in enumerate(types) if t == "constant"]
        table_const = pd.DataFrame({"feature": [features[i] for i in const], "type": ["constant"] * len(const)},
                                   index=pd.Index([features[i] for i in const], name="feature"))
        table_const["p_value"] = np.nan
        table_const["relevant"] = False
        if not table_const.empty:
            warnings.warn("[test_feature_significance] Constant features: {}".format(", ".join(map(str, table_const.feature))),
                          RuntimeWarning)
        if len(table_const) == len(features):
            return table_const

        if ml_task == "regression":
            # relevance.py:282-296: Kendall's tau for real features, two-sample KS of the target for binary ones
            n = int(ytail[3])
            real = [i for i, t in enumerate(types) if t == "real"]
            binary = [i for i, t in enumerate(types) if t == "binary"]
            p = {}
            for i in real:
                s_ = rstats[i]
                p[i] = kendall_pvalue(n, int(s_[2]), int(s_[3]), s_[4], s_[5], int(s_[6]), int(ytail[0]), ytail[1], ytail[2])
            for i in binary:
                s_ = rstats[i]
                p[i] = ks_2samp_pvalue(s_[2], s_[3], s_[4])
            idx = real + binary
            relevance_table = pd.DataFrame({"feature": [features[i] for i in idx], "type": [types[i] for i in idx],
                                            "p_value": [p[i] for i in idx]}, index=pd.Index([features[i] for i in idx], name="feature"))
            relevance_table["relevant"] = benjamini_reject(relevance_table.p_value.to_numpy(), fdr_level, hypotheses_independent)
            relevance_table = relevance_table.sort_values("p_value")
            labels = []
        tables = []
        for k, label in enumerate(labels):
            tmp = _table_for_label(features, types, stats[k], test_for_binary_target_real_feature, fdr_level,
                                   hypotheses_independent)
            if multiclass:
                tmp = tmp.reset_index(drop=True)
                tmp.columns = tmp.columns.map(lambda x: (x + "_" + str(label) if x != "feature" and x != "type" else x))
            tables.append(tmp)
        if ml_task == "regression":
            pass
        elif multiclass:
            relevance_table = reduce(lambda left, right: pd.merge(left, right, on=["feature", "type"], how="outer"), tables)
            relevance_table["n_significant"] = relevance_table.filter(regex="^relevant_", axis=1).sum(axis=1)
            relevance_table["relevant"] = relevance_table["n_significant"] >= n_significant
            relevance_table.index = relevance_table["feature"]
        else:
            relevance_table = combine_relevance_tables(tables)

        if multiclass:
            for column in relevance_table.filter(regex="^relevant_", axis=1).columns:
                table_const[column] = False
            table_const["n_significant"] = 0
            table_const.drop(columns=["p_value"], inplace=True)
        relevance_table = pd.concat([relevance_table, table_const], axis=0)
        if sum(relevance_table["relevant"]) == 0:
            warnings.warn("No feature was found relevant for {} for fdr level = {} (which corresponds to the maximal percentage "
                          "of irrelevant features, consider using an higher fdr level or add other features."
                          .format(ml_task, fdr_level), RuntimeWarning)
    return relevance_table


def select_features(X, y, test_for_binary_target_binary_feature=TEST_FOR_BINARY_TARGET_BINARY_FEATURE,
                    test_for_binary_target_real_feature=TEST_FOR_BINARY_TARGET_REAL_FEATURE,
                    test_for_real_target_binary_feature=TEST_FOR_REAL_TARGET_BINARY_FEATURE,
                    test_for_real_target_real_feature=TEST_FOR_REAL_TARGET_REAL_FEATURE, fdr_level=FDR_LEVEL,
                    hypotheses_independent=HYPOTHESES_INDEPENDENT, n_jobs=N_PROCESSES, show_warnings=False, chunksize=None,
                    ml_task="auto", multiclass=False, n_significant=1, device=None):
    """GPU implementation of tsfresh.feature_selection.selection.select_features (selection.py:17-200)."""
    assert isinstance(X, pd.DataFrame), "Please pass features in X as pandas.DataFrame."
    assert isinstance(y, (pd.Series, np.ndarray)), "The type of target vector y must be one of: pandas.Series, numpy.ndarray"
    assert len(y) > 1, "y must contain at least two samples."
    assert len(X) == len(y), "X and y must contain the same number of samples."
    assert len(set(y)) > 1, "Feature selection is only possible if more than 1 label/class is provided"
    if isinstance(y, pd.Series) and set(X.index) != set(y.index):
        raise ValueError("Index of X and y must be identical if provided")
    if isinstance(y, np.ndarray):
        y = pd.Series(y, index=X.index)
    relevance_table = calculate_relevance_table(
        X, y, ml_task=ml_task, multiclass=multiclass, n_significant=n_significant, n_jobs=n_jobs, show_warnings=show_warnings,
        chunksize=chunksize, test_for_binary_target_real_feature=test_for_binary_target_real_feature, fdr_level=fdr_level,
        hypotheses_independent=hypotheses_independent, device=device)
    relevant_features = relevance_table[relevance_table.relevant].feature
    return X.loc[:, relevant_features]
