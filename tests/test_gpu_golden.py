"""GPU path against the committed golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/): the 783-column ComprehensiveFCParameters matrix on 12 ragged series
and the reference's own 80-row test fixture through extract_features()."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle.extract import NOISE_FLOOR, compare
from tsfresh_b200 import ComprehensiveFCParameters, EfficientFCParameters, extract_features
from tsfresh_b200.plan import Plan

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_comprehensive_matches_reference_golden():
    from tsfresh_b200._lib import Context, DevicePlan
    z = np.load(os.path.join(G, "comprehensive.npz"))
    plan = Plan(ComprehensiveFCParameters())
    ctx = Context(0)
    dp = DevicePlan(ctx, plan)
    got = dp.extract_csr(z["values"], z["begin"], z["length"])
    dp.close()
    ctx.close()
    bad = compare(got, z["reference"], plan.suffixes, atol=NOISE_FLOOR)         # ragged fixture incl. 1..5-sample and constant series
    assert not bad, bad[:30]


def test_fixture80_through_extract_features():
    z = np.load(os.path.join(G, "fixture80.npz"))
    df = pd.DataFrame({"id": z["id"], "sort": z["sort"], "kind": z["kind"], "val": z["val"]})
    X = extract_features(df, column_id="id", column_sort="sort", column_kind="kind", column_value="val",
                         default_fc_parameters=EfficientFCParameters())
    assert list(X.columns) == list(z["columns"]) and list(X.index) == list(z["index"])
    suffixes = [c.split("__", 1)[1] for c in X.columns]
    bad = compare(X.to_numpy(), z["reference"], suffixes, atol=NOISE_FLOOR)     # 20 small integers per series
    # the fixture is 20 small integers per series: many exact ties.  permutation_entropy on a window that contains a tie is
    # implementation-defined in the reference (numpy's default argsort is unstable), so that column
    # is compared exactly where EVERY window of the series is tie-free, and skipped only on the others.
    def tie_free(row_id, kind, suffix):
        m = __import__("re").search(r"dimension_(\d+)__tau_(\d+)", suffix)
        D, tau = int(m.group(1)), int(m.group(2))
        sel = (z["id"] == row_id) & (z["kind"] == kind)
        x = z["val"][sel][np.argsort(z["sort"][sel], kind="stable")]
        wins = [x[i:i + D * tau:tau] for i in range(0, len(x) - (D - 1) * tau)]
        return all(len(set(w.tolist())) == len(w) for w in wins)

    kept, checked = [b for b in bad if not b[1].startswith("permutation_entropy")], 0
    for c, name in enumerate(X.columns):
        kind, suffix = name.split("__", 1)
        if not suffix.startswith("permutation_entropy"):
            continue
        for r, rid in enumerate(X.index):
            if tie_free(rid, kind, suffix):
                checked += 1
                g, w = X.iloc[r, c], z["reference"][r, c]
                assert (np.isnan(g) and np.isnan(w)) or np.isclose(g, w, rtol=1e-5, atol=0.0), (name, rid, g, w)
    assert checked >= 20          # every permutation_entropy cell of the fixture is tie-free window by window
    bad = kept
    assert not bad, bad[:30]
    for name, want in (("a__maximum", [71, 77]), ("a__sum_values", [691, 1017]), ("a__abs_energy", [32211, 63167]),
                       ("b__mean", [37.85, 34.75]), ("b__median", [39.5, 28.0])):      # test_extraction.py:40-55
        np.testing.assert_allclose(X[name].to_numpy(), want)
