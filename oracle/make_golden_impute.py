"""TEST INFRASTRUCTURE (build container only): tests/golden/impute.npz from the UNMODIFIED reference.

    python -m oracle.make_golden_impute

Runs tsfresh.utilities.dataframe_functions.impute / impute_dataframe_zero / get_range_values_per_column
(dataframe_functions.py:49-212) on a small matrix that exercises every branch: NaN, +inf, -inf, a column without any
finite value, clean columns, odd and even finite counts, repeated middle values.
"""
import os
import warnings

import numpy as np
import pandas as pd

from . import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def make_input(seed=5, rows=201, cols=14):
    rng = np.random.default_rng(seed)
    m = rng.standard_normal((rows, cols)) * rng.uniform(0.1, 50.0, cols)
    m[rng.random((rows, cols)) < 0.08] = np.nan
    m[rng.random((rows, cols)) < 0.03] = np.inf
    m[rng.random((rows, cols)) < 0.03] = -np.inf
    m[:, 3] = np.nan                       # no finite value at all
    m[:, 4] = rng.standard_normal(rows)    # clean column
    m[:, 5] = np.round(m[:, 5])            # ties around the median
    m[::2, 6] = np.nan                     # even / odd finite counts
    m[1::2, 7] = np.inf
    m[:, 8] = np.where(np.arange(rows) % 3 == 0, -np.inf, np.nan)      # only non-finite, mixed
    m[0, 9] = np.nan; m[1:, 9] = 2.5       # constant column with one NaN
    return m


def main():
    ref_shim.load()
    from tsfresh.utilities import dataframe_functions as rdf
    m = make_input()
    cols = ["c%d" % i for i in range(m.shape[1])]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cmax, cmin, cmed = rdf.get_range_values_per_column(pd.DataFrame(m.copy(), columns=cols))
        imputed = rdf.impute(pd.DataFrame(m.copy(), columns=cols)).to_numpy(np.float64)
        zero = rdf.impute_dataframe_zero(pd.DataFrame(m.copy(), columns=cols)).to_numpy(np.float64)
    stats = np.array([[float(cmin[c]) for c in cols], [float(cmax[c]) for c in cols], [float(cmed[c]) for c in cols]])
    np.savez_compressed(os.path.join(OUT, "impute.npz"), input=m, imputed=imputed, zero=zero, stats=stats)
    print("wrote impute.npz", m.shape, "non-finite left:", int((~np.isfinite(imputed)).sum()))


if __name__ == "__main__":
    main()
