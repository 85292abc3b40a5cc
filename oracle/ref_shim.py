"""TEST INFRASTRUCTURE (build container only): import the UNMODIFIED reference from /root/reference.

statsmodels / pywt / stumpy are not installable here, so empty stand-in modules are registered in
`sys.modules` before the import (recipe: SURVEY.md section 8c); the five calculators that call them
get the float64 restatements of oracle/thirdparty.py.  Everything else (70 of 75 calculators) is the
reference's own code running on the installed numpy/scipy/pandas.

/root/reference does not exist on the GPU box: only oracle/make_golden.py and the `not gpu` tests that
are skipped when the directory is absent may call `load()`.
"""
import os
import sys
import types
import unittest.mock

REFERENCE_ROOT = os.environ.get("TSFX_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tsfresh"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


def _missing(*a, **k):
    raise NotImplementedError("third-party routine not available in this image")


def load():
    """Returns the imported reference package `tsfresh` (cached in sys.modules)."""
    if "tsfresh" in sys.modules and getattr(sys.modules["tsfresh"], "__tsfx_shim__", False):
        return sys.modules["tsfresh"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from . import thirdparty as tp

    _stub("pywt", cwt=tp.cwt)
    _stub("stumpy").core = _stub("stumpy.core", mass=_missing, mass_absolute=_missing)
    _stub("statsmodels")
    _stub("statsmodels.tools")
    _stub("statsmodels.tsa")
    _stub("statsmodels.stats")
    _stub("statsmodels.tools.sm_exceptions", MissingDataError=tp.MissingDataError)
    _stub("statsmodels.tsa.ar_model", AutoReg=tp.AutoReg)
    _stub("statsmodels.tsa.stattools", acf=tp.acf, adfuller=tp.adfuller, pacf=tp.pacf)
    _stub("statsmodels.stats.multitest", multipletests=tp.multipletests)
    sys.modules.setdefault("mock", unittest.mock)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import tsfresh  # noqa: E402

    tsfresh.__tsfx_shim__ = True
    return tsfresh
