"""Not a test: times the user-facing extract_features(DataFrame) call on BASELINE.json configs[1]
(EfficientFCParameters, 100 000 series x 256) and prints where the host time goes."""
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, ".")
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
