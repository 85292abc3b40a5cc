"""TEST INFRASTRUCTURE (build container only): golden vectors for linear_trend_timewise from the UNMODIFIED reference
(tsfresh.extract_features on a frame with a DatetimeIndex, feature_calculators.py:2274-2306) -> tests/golden/timewise.npz."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import ref_shim
    ref_shim.load()
    from tsfresh.feature_extraction import extract_features
    rng = np.random.default_rng(2024)
    lens = [2, 3, 5, 40, 256, 300]
    ids = np.concatenate([np.full(n, 10 + i) for i, n in enumerate(lens)])
    t = np.concatenate([np.sort(rng.integers(0, 10 ** 6, n) * 10 ** 9 + rng.integers(0, 10 ** 9, n)) for n in lens]) + 1_514_764_800 * 10 ** 9
    v = rng.standard_normal(len(ids)).astype(np.float32)
    fc = {"linear_trend_timewise": [{"attr": a} for a in ("pvalue", "rvalue", "intercept", "slope", "stderr")],
          "mean": None, "linear_trend": [{"attr": "slope"}]}
    df = pd.DataFrame({"id": ids, "value": v.astype(np.float64)}, index=pd.DatetimeIndex(t.astype("datetime64[ns]")))
    X = extract_features(df, column_id="id", default_fc_parameters=fc, n_jobs=0, disable_progressbar=True)
    np.savez(os.path.join(ROOT, "tests", "golden", "timewise.npz"), id=ids, t_ns=t, value=v, columns=np.array(list(X.columns)),
             index=np.asarray(X.index), reference=X.to_numpy())
    print(X)


if __name__ == "__main__":
    main()
