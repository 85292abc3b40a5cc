"""TEST INFRASTRUCTURE (build container only): tests/golden/roll_cases.npz from the UNMODIFIED reference.

    python -m oracle.make_golden_roll

tsfresh.utilities.dataframe_functions.roll_time_series (dataframe_functions.py:376-603) on a small ragged frame for a
list of (rolling_direction, max_timeshift, min_timeshift) cases, both directions; stored per case as the sorted list of
(parent series, index of the id row, first row of the window, window length).
"""
import os
import warnings

import numpy as np
import pandas as pd

from . import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
LENS = [20, 9, 31, 1, 2]
CASES = [(-1, 7, 0), (-3, 7, 7), (-2, 4, 2), (-5, 100, 0), (-1, 100, 0), (3, 7, 0), (1, 100, 3), (4, 5, 5), (2, 1, 1)]


def reference_windows(lens, rd, mx, mn):
    ref_shim.load()
    from tsfresh.utilities.dataframe_functions import roll_time_series
    rng = np.random.default_rng(3)
    df = pd.DataFrame({"id": np.concatenate([np.full(n, i) for i, n in enumerate(lens)]),
                       "time": np.concatenate([np.arange(n) for n in lens]),
                       "value": rng.standard_normal(sum(lens))})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = roll_time_series(df, column_id="id", column_sort="time", rolling_direction=rd, max_timeshift=mx,
                             min_timeshift=mn, n_jobs=0, disable_progressbar=True)
    g = r.groupby("id", sort=False)
    first, size = g["time"].min(), g.size()
    return sorted((int(i[0]), int(i[1]), int(first[i]), int(size[i])) for i in g.groups.keys())


def main():
    out = {"lens": np.asarray(LENS, np.int32), "cases": np.asarray(CASES, np.int32)}
    for k, (rd, mx, mn) in enumerate(CASES):
        out["case%d" % k] = np.asarray(reference_windows(LENS, rd, mx, mn), np.int64).reshape(-1, 4)
    np.savez_compressed(os.path.join(OUT, "roll_cases.npz"), **out)
    print("wrote roll_cases.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
