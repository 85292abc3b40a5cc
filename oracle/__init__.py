"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy / scipy / pandas, float64) of the tsfresh feature-extraction hot path
(`_do_extraction_on_chunk` over the calculator registry,
/root/reference/tsfresh/feature_extraction/extraction.py:308-386 and feature_calculators.py:238-2521).

Nothing in the product path (`tsfresh_b200/`) may import this package.  Allowed importers:
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

Pinning status (see DESIGN.md "Oracle"):
  * 70 of 75 calculators are pinned against the UNMODIFIED reference code imported in the build
    container (oracle/ref_shim.py + oracle/make_golden.py -> tests/golden/*.npz) and against the
    reference's own known-answer tests (tests/test_oracle_known_answers.py).
  * agg_autocorrelation, partial_autocorrelation, ar_coefficient are pinned by the reference's
    known-answer tests through our restatement of the statsmodels routines (oracle/thirdparty.py).
  * augmented_dickey_fuller (teststat / pvalue) and cwt_coefficients values: PARITY UNPINNED --
    statsmodels / PyWavelets are not installable here and the reference holds no value tests for them.
"""
