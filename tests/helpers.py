import numpy as np

from oracle.extract import compare, oracle_rows
from tsfresh_b200.plan import Plan


def synthetic_series(seed, n_series, length, kind="normal"):
    rng = np.random.default_rng(seed)
    if kind == "normal":
        v = rng.standard_normal((n_series, length))
    elif kind == "walk":
        v = rng.standard_normal((n_series, length)).cumsum(axis=1)
    elif kind == "rounded":            # many ties / duplicates
        v = np.round(rng.standard_normal((n_series, length)) * 2) / 2
    else:
        raise ValueError(kind)
    return v.astype(np.float32)


def to_csr(series):
    lens = np.array([len(s) for s in series], dtype=np.int32)
    begin = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    values = np.concatenate([np.asarray(s, dtype=np.float32) for s in series]) if len(series) else np.zeros(0, np.float32)
    return values, begin, lens


def gpu_vs_oracle(ctx, settings, series, rtol=1e-5, atol=0.0):
    """Runs `settings` over the list of float32 series on the GPU (C-ABI, CSR path) and in the oracle;
    returns (mismatches, plan, gpu_matrix, oracle_matrix)."""
    from tsfresh_b200._lib import DevicePlan
    plan = Plan(settings)
    dp = DevicePlan(ctx, plan)
    try:
        values, begin, lens = to_csr(series)
        got = dp.extract_csr(values, begin, lens)
    finally:
        dp.close()
    want = oracle_rows([np.asarray(s, dtype=np.float32).astype(np.float64) for s in series], settings)
    return compare(got, want, plan.suffixes, rtol=rtol, atol=atol), plan, got, want
