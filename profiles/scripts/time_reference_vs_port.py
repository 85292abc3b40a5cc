#!/usr/bin/env python
"""Build container only (needs /root/reference): wall clock of the UNMODIFIED reference's
extract_features(n_jobs=cores) -- adapter, MultiprocessingDistributor, per-series loop, pivot -- next to the oracle port
that bench.py --impl reference times on the GPU box (the Python reference cannot travel).  BASELINE.md section 3.1.
Writes profiles/reference_vs_port_r2.json."""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ[k] = "1"


def main():
    from oracle import ref_shim
    tsfresh = ref_shim.load()
    from tsfresh.feature_extraction import ComprehensiveFCParameters, extract_features
    import bench
    cores = bench.usable_cores()
    S, L = int(os.environ.get("N_SERIES", 64 * cores)), 256
    rng = np.random.default_rng(42)
    x = rng.standard_normal((S, L)).astype(np.float32).astype(np.float64)
    df = pd.DataFrame({"id": np.repeat(np.arange(S), L), "time": np.tile(np.arange(L), S), "value": x.reshape(-1)})
    res = {}
    for n_jobs in (cores, 0):
        sub = df if n_jobs else df[df["id"] < max(8, S // cores)]
        ns = sub["id"].nunique()
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            X = extract_features(sub, column_id="id", column_sort="time", default_fc_parameters=ComprehensiveFCParameters(),
                                 n_jobs=n_jobs, disable_progressbar=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        res["reference_n_jobs_%d" % n_jobs] = {"series": int(ns), "seconds": best, "series_per_s": ns / best, "columns": int(X.shape[1])}
    port = bench.cpu_baseline(L, "comprehensive", target_seconds=float(S) / cores * 0.12, cores=cores)
    res["port_all_cores"] = port
    res["cores"] = cores
    res["ratio_port_over_reference_all_cores"] = port["value"] / res["reference_n_jobs_%d" % cores]["series_per_s"]
    res["note"] = ("unmodified reference through oracle/ref_shim.py (statsmodels / pywt calls answered by the restatements of "
                   "oracle/thirdparty.py); the reference returns 788 columns (5 linear_trend_timewise skipped as NaN without a DatetimeIndex)")
    json.dump(res, open(os.path.join(ROOT, "profiles", "reference_vs_port_r2.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
