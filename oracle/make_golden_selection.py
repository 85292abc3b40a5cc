"""TEST INFRASTRUCTURE (build container only): golden relevance tables from the UNMODIFIED reference
(tsfresh.feature_selection.relevance.calculate_relevance_table, relevance.py:31-322; scipy tests; the Benjamini step is the
restated multipletests of oracle/thirdparty.py because statsmodels is not installable here) -> tests/golden/selection.npz."""
import os
import sys
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_X(rng, n):
    y2 = rng.integers(0, 2, n)
    cols = {}
    for k in range(6):
        cols["noise_%d" % k] = rng.standard_normal(n)
    for k in range(4):
        cols["signal_%d" % k] = rng.standard_normal(n) + (0.25 + 0.25 * k) * y2
    cols["counts"] = rng.integers(0, 7, n).astype(float) + y2 * rng.integers(0, 2, n)          # heavy ties
    cols["halves"] = np.round(rng.standard_normal(n) * 2) / 2
    cols["binary_rel"] = ((rng.random(n) < 0.3 + 0.4 * y2) * 1.0)
    cols["binary_irr"] = (rng.random(n) < 0.5) * 3.0 - 1.0                                     # two values, not 0/1
    cols["constant"] = np.full(n, 2.5)
    cols["neg_zero"] = np.where(rng.random(n) < 0.5, 0.0, -0.0) + (rng.random(n) < 0.2)       # -0.0 == 0.0: binary
    return pd.DataFrame(cols), y2


def main():
    from oracle import ref_shim
    ref_shim.load()
    from tsfresh.feature_selection.relevance import calculate_relevance_table
    rng = np.random.default_rng(7)
    out = {}
    for tag, n in (("small", 37), ("medium", 400), ("large", 3000)):
        X, y2 = make_X(rng, n)
        y3 = (y2 + (rng.random(n) < 0.3) * (1 + y2)).astype(np.int64) % 3
        out[tag + "_X"] = X.to_numpy()
        out[tag + "_y2"] = y2
        out[tag + "_y3"] = y3
        out["columns"] = np.array(list(X.columns))
        yr = (X["signal_1"].to_numpy() * 0.5 + 0.2 * X["binary_rel"].to_numpy() + rng.standard_normal(n))      # regression target
        yt = np.round(yr * 2) / 2                                                                               # ... with ties
        out[tag + "_yr"] = yr
        out[tag + "_yt"] = yt
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for ytag, y, kw in (("yr", yr, {}), ("yt", yt, {"hypotheses_independent": True}), ("y2", y2, {}), ("y2smir", y2, {"test_for_binary_target_real_feature": "smir"}),
                                ("y2indep", y2, {"hypotheses_independent": True, "fdr_level": 0.2}),
                                ("y3", y3, {"multiclass": True, "n_significant": 2})):
                task = "regression" if ytag in ("yr", "yt") else "classification"
                t = calculate_relevance_table(X, pd.Series(y, index=X.index), ml_task=task, n_jobs=0, **kw)
                key = "%s_%s" % (tag, ytag)
                out[key + "_index"] = np.array(list(t.index))
                out[key + "_columns"] = np.array(list(t.columns))
                for c in t.columns:
                    v = t[c].to_numpy()
                    out[key + "_col_" + c] = v.astype(np.float64) if v.dtype.kind in "fbiu" else v.astype(str)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "selection.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_index")})


if __name__ == "__main__":
    main()
