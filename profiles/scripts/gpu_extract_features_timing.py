"""Not a test: times the user-facing extract_features(DataFrame) call on BASELINE.json configs[1]
(EfficientFCParameters, 100 000 series x 256) and prints where the host time goes."""
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tsfresh_b200 import EfficientFCParameters, extract_features  # noqa: E402

N, L = 100_000, 256
rng = np.random.default_rng(43)
df = pd.DataFrame({"id": np.repeat(np.arange(N), L), "time": np.tile(np.arange(L), N),
                   "value": rng.standard_normal(N * L).astype(np.float32)})
s = EfficientFCParameters()
for rep in range(3):
    t0 = time.perf_counter()
    X = extract_features(df, column_id="id", column_sort="time", default_fc_parameters=s)
    dt = time.perf_counter() - t0
    print("extract_features(df %d x %d, Efficient): %.3f s -> %.0f series/s, result %s" % (N, L, dt, N / dt, X.shape))
sh = df.sample(frac=1.0, random_state=0).reset_index(drop=True)
t0 = time.perf_counter()
X2 = extract_features(sh, column_id="id", column_sort="time", default_fc_parameters=s)
dt = time.perf_counter() - t0
print("same frame, rows shuffled (device sort): %.3f s -> %.0f series/s, equal=%s" % (dt, N / dt, np.array_equal(X.to_numpy(), X2.to_numpy(), equal_nan=True)))

# ---- stage (a) alone: long frame -> CSR on the device (tsfx_build_csr), host buffers in, nothing copied back
from tsfresh_b200 import _lib  # noqa: E402
from tsfresh_b200.extraction import get_context  # noqa: E402
import ctypes  # noqa: E402

ctx = get_context(0)
for name, frame in (("sorted rows", df), ("shuffled rows", sh)):
    ids = np.ascontiguousarray(frame["id"].to_numpy(np.int64))
    keys = np.ascontiguousarray(frame["time"].to_numpy(np.int64))
    vals = np.ascontiguousarray(frame["value"].to_numpy(np.float32))
    k = ctypes.c_int64(0)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        rc = ctx.lib.tsfx_build_csr(ctx.h, ids.ctypes.data, keys.ctypes.data, 0, vals.ctypes.data, len(ids), None, None, None, None, 0,
                                    ctypes.byref(k))
        ctx.sync()
        best = min(best, time.perf_counter() - t0)
        assert rc == 0 and k.value == N
    print("stage (a) tsfx_build_csr, %s: %d rows in %.3f s = %.1f M rows/s (%.2f GB/s of (id, time, value) input incl. pageable H2D)"
          % (name, len(ids), best, len(ids) / best / 1e6, len(ids) * 20 / best / 1e9))
