"""TEST INFRASTRUCTURE (build container only): generates tests/golden/ from the UNMODIFIED reference.

    python -m oracle.make_golden

Imports the reference from /root/reference through oracle/ref_shim.py (statsmodels / pywt replaced by the
restatements of oracle/thirdparty.py -- the 87 columns they feed are flagged in `restated_columns`) and
stores small input/output vectors that travel to the GPU box:

  tests/golden/comprehensive.npz   inputs (float32, ragged) + reference extract_features() matrix (783 columns)
  tests/golden/columns.json        column names of the three settings presets, in reference order
  tests/golden/roll.npz            reference roll_time_series() window ids for a small frame
  tests/golden/fixture80.npz       the reference's own 80-row test fixture (tests/fixtures.py:28-198) + outputs
"""
import json
import os
import warnings

import numpy as np
import pandas as pd

from . import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
RESTATED = ("agg_autocorrelation", "partial_autocorrelation", "augmented_dickey_fuller", "ar_coefficient", "cwt_coefficients")


def frame(series):
    ids = np.concatenate([np.full(len(s), i) for i, s in enumerate(series)])
    t = np.concatenate([np.arange(len(s)) for s in series])
    v = np.concatenate([np.asarray(s, np.float32).astype(np.float64) for s in series])
    return pd.DataFrame({"id": ids, "time": t, "value": v})


def main():
    ref_shim.load()
    from tsfresh.feature_extraction import extract_features, settings as rs
    from tsfresh.utilities.dataframe_functions import roll_time_series
    os.makedirs(OUT, exist_ok=True)
    warnings.simplefilter("ignore")

    rng = np.random.default_rng(20260922)
    lengths = [256, 256, 256, 256, 128, 100, 64, 37, 20, 9, 1024, 300]
    series = []
    for i, n in enumerate(lengths):
        x = rng.standard_normal(n)
        if i % 3 == 1:
            x = x.cumsum()                        # random walk (the reference's own benchmark generator)
        series.append(x.astype(np.float32))
    X = extract_features(frame(series), column_id="id", column_sort="time", default_fc_parameters=rs.ComprehensiveFCParameters(),
                         n_jobs=0, disable_progressbar=True)
    begin = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "comprehensive.npz"), values=np.concatenate(series), begin=begin,
                        length=np.asarray(lengths, np.int32), reference=X.to_numpy(dtype=np.float64))
    cols = {}
    for name, cls in (("comprehensive", rs.ComprehensiveFCParameters), ("efficient", rs.EfficientFCParameters),
                      ("minimal", rs.MinimalFCParameters)):
        Xs = extract_features(frame(series[:2]), column_id="id", column_sort="time", default_fc_parameters=cls(), n_jobs=0,
                              disable_progressbar=True)
        cols[name] = [c[len("value__"):] for c in Xs.columns]
    cols["restated_columns"] = [c for c in cols["comprehensive"] if c.split("__")[0] in RESTATED]
    json.dump(cols, open(os.path.join(OUT, "columns.json"), "w"), indent=0)

    # roll_time_series: windows of 8 rows, stride 3, on ragged series
    lens = [20, 9, 31]
    df = frame([rng.standard_normal(n).astype(np.float32) for n in lens])
    rolled = roll_time_series(df, column_id="id", column_sort="time", rolling_direction=3, max_timeshift=7, min_timeshift=7,
                              n_jobs=0, disable_progressbar=True)
    g = rolled.groupby("id", sort=False)
    ids = list(g.groups.keys())
    parent = np.array([i[0] for i in ids], dtype=np.int64)
    t_end = np.array([i[1] for i in ids], dtype=np.int64)
    first_time = g["time"].min().to_numpy().astype(np.int64)
    count = g.size().to_numpy().astype(np.int64)
    rolled2 = roll_time_series(df, column_id="id", column_sort="time", rolling_direction=3, max_timeshift=7, n_jobs=0,
                               disable_progressbar=True)
    g2 = rolled2.groupby("id", sort=False)
    ids2 = list(g2.groups.keys())
    np.savez_compressed(os.path.join(OUT, "roll.npz"), lens=np.asarray(lens, np.int32), parent=parent, t_end=t_end,
                        first_time=first_time, count=count,
                        parent_nomin=np.array([i[0] for i in ids2], dtype=np.int64),
                        t_end_nomin=np.array([i[1] for i in ids2], dtype=np.int64),
                        count_nomin=g2.size().to_numpy().astype(np.int64))

    # the reference's 80-row fixture (two ids x two kinds), values are small integers
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_fixtures", os.path.join(ref_shim.REFERENCE_ROOT, "tests", "fixtures.py"))
    fx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fx)
    case = fx.DataTestCase()
    df80 = case.create_test_data_sample()
    X80 = extract_features(df80, column_id="id", column_sort="sort", column_kind="kind", column_value="val",
                           default_fc_parameters=rs.EfficientFCParameters(), n_jobs=0, disable_progressbar=True)
    np.savez_compressed(os.path.join(OUT, "fixture80.npz"), id=df80["id"].to_numpy(np.int64), sort=df80["sort"].to_numpy(np.int64),
                        kind=df80["kind"].to_numpy().astype("U1"), val=df80["val"].to_numpy(np.float64),
                        index=X80.index.to_numpy(np.int64), columns=np.array(list(X80.columns)), reference=X80.to_numpy(np.float64))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
