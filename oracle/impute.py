"""CPU restatement of the reference's imputation helpers -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows tsfresh/utilities/dataframe_functions.py: get_range_values_per_column :170-212 (finite max / min / median per
column through numpy masked arrays; columns without any finite value -> 0), impute_dataframe_range :104-167,
impute :49-78, impute_dataframe_zero :81-101 -- on a plain float64 matrix instead of a DataFrame.
Pinned against the unmodified reference in tests/test_oracle_vs_reference.py and tests/golden/impute.npz.
"""
import numpy as np


def range_values(m):
    """-> stats[3, cols]: min | max | median of the finite values of every column (0 when there are none)."""
    m = np.asarray(m, dtype=np.float64)
    rows, cols = m.shape
    stats = np.zeros((3, cols))
    for c in range(cols):
        v = m[:, c]
        f = np.sort(v[np.isfinite(v)])
        if f.size == 0:
            continue                                  # :194-203
        stats[0, c], stats[1, c] = f[0], f[-1]
        lo, hi = f[(f.size - 1) // 2], f[f.size // 2]
        stats[2, c] = lo if lo == hi else (lo + hi) / 2.0      # np.ma.median: mean of the two middle values
    return stats


def apply_range(m, stats):
    out = np.array(m, dtype=np.float64, copy=True)
    for c in range(out.shape[1]):
        v = out[:, c]
        v[v == np.inf] = stats[1, c]
        v[v == -np.inf] = stats[0, c]
        v[np.isnan(v)] = stats[2, c]
    return out


def impute(m):
    return apply_range(m, range_values(m))


def impute_zero(m):
    out = np.array(m, dtype=np.float64, copy=True)
    out[~np.isfinite(out)] = 0.0
    return out
