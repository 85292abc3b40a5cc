"""TEST INFRASTRUCTURE: float64 numpy restatements of the third-party routines the reference calls on
the hot path but which are NOT installable in this image (no network):

  statsmodels (setup.cfg:42, unpinned >=0.13; restated from the published 0.14 algorithms)
    acf       <- feature_calculators.py:429   (agg_autocorrelation)
    pacf      <- feature_calculators.py:490   (partial_autocorrelation, method="ld")
    adfuller  <- feature_calculators.py:521   (augmented_dickey_fuller, autolag="AIC")
    AutoReg   <- feature_calculators.py:1493  (ar_coefficient)
  PyWavelets (setup.cfg:44, unpinned; restated from the published 1.x algorithm)
    cwt       <- feature_calculators.py:1402  (cwt_coefficients, wavelet "mexh")

Pinning: acf / pacf / AutoReg reproduce the reference's own known-answer tests
(tests/units/feature_extraction/test_feature_calculations.py:238-344, 1077-1127).  adfuller's
teststat/pvalue and cwt values are PARITY UNPINNED (the reference has no value tests for them).
"""
import math

import numpy as np


class MissingDataError(Exception):
    """Stand-in for statsmodels.tools.sm_exceptions.MissingDataError."""


# ----------------------------------------------------------------------------- acovf / acf / pacf
def acovf_adjusted(x):
    """statsmodels.tsa.stattools.acovf(x, adjusted=True, demean=True, fft=False)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    xo = x - x.mean()
    full = np.correlate(xo, xo, "full")[n - 1:]
    return full / (n - np.arange(n))


def acf(x, adjusted=True, fft=False, nlags=None, **_):
    """statsmodels acf: acovf[:nlags+1] / acovf[0].  `fft` only changes how the same sums are formed."""
    if not adjusted:
        raise NotImplementedError("only adjusted=True is on the tsfresh path")
    x = np.asarray(x, dtype=np.float64)
    if np.isnan(x).any() or np.isinf(x).any():
        raise MissingDataError("exog contains inf or nans")
    avf = acovf_adjusted(x)
    if nlags is None:
        nlags = min(int(10 * np.log10(len(x))), len(x) - 1)
    return avf[: nlags + 1] / avf[0]


def levinson_durbin_pacf(acv, nlags):
    """statsmodels levinson_durbin(acv, nlags, isacov=True)[2] (the partial autocorrelations)."""
    phi = np.zeros((nlags + 1, nlags + 1))
    sig = np.zeros(nlags + 1)
    with np.errstate(all="ignore"):
        phi[1, 1] = acv[1] / acv[0]
        sig[1] = acv[0] - phi[1, 1] * acv[1]
        for k in range(2, nlags + 1):
            phi[k, k] = (acv[k] - np.dot(phi[1:k, k - 1], acv[1:k][::-1])) / sig[k - 1]
            for j in range(1, k):
                phi[j, k] = phi[j, k - 1] - phi[k, k] * phi[k - j, k - 1]
            sig[k] = sig[k - 1] * (1 - phi[k, k] ** 2)
    out = np.diag(phi).copy()
    out[0] = 1.0
    return out


def pacf(x, nlags=None, method="ld", **_):
    if method not in ("ld", "ldadjusted", "lda"):
        raise NotImplementedError("only method='ld' is on the tsfresh path")
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if nlags is None:
        nlags = min(int(10 * np.log10(n)), n // 2 - 1)
    if nlags >= n // 2:
        raise ValueError("Can only compute partial correlations for lags up to 50% of the sample size.")
    acv = acovf_adjusted(x)
    return levinson_durbin_pacf(acv, nlags)


# ----------------------------------------------------------------------------- adfuller
_TAU_STAR_C, _TAU_MIN_C, _TAU_MAX_C = -1.61, -18.83, 2.74
_TAU_C_SMALLP = (2.1659, 1.4412, 3.8269e-2)
_TAU_C_LARGEP = (1.7339, 9.3202e-1, -1.2745e-1, -1.0368e-2)


def norm_cdf(z):
    return 0.5 * math.erfc(-z / math.sqrt(2.0))


def mackinnonp_c(stat):
    """statsmodels.tsa.adfvalues.mackinnonp(stat, regression='c', N=1)."""
    if stat > _TAU_MAX_C:
        return 1.0
    if stat < _TAU_MIN_C:
        return 0.0
    coef = _TAU_C_SMALLP if stat <= _TAU_STAR_C else _TAU_C_LARGEP
    acc = 0.0
    for c in reversed(coef):
        acc = acc * stat + c
    return norm_cdf(acc)


def _ols(y, X):
    """OLS via pinv like statsmodels (params, ssr, (X'X)^+)."""
    pinv = np.linalg.pinv(X)
    beta = pinv @ y
    resid = y - X @ beta
    return beta, float(resid @ resid), pinv @ pinv.T


def _lagmat_both_in(d, maxlag):
    """lagmat(d[:,None], maxlag, trim='both', original='in'): col0 = d[t], col j = d[t-j], t=maxlag..len-1."""
    n = d.shape[0]
    return np.column_stack([d[maxlag - j: n - j] for j in range(maxlag + 1)])


def adfuller(x, autolag="AIC", **_):
    """adfuller(x, maxlag=None, regression='c', autolag='AIC') -> (teststat, pvalue, usedlag)."""
    if autolag is not None:
        autolag = autolag.lower()
    if autolag not in ("aic", "bic", "t-stat", None):
        raise ValueError("autolag can only be None or one of 'AIC', 'BIC', 't-stat'")
    x = np.asarray(x, dtype=np.float64)
    if not np.isfinite(x).all():
        raise MissingDataError("exog contains inf or nans")
    if x.max() == x.min():
        raise ValueError("Invalid input, x is constant")
    n = x.shape[0]
    maxlag = int(np.ceil(12.0 * np.power(n / 100.0, 1 / 4.0)))
    maxlag = min(n // 2 - 1 - 1, maxlag)
    if maxlag < 0:
        raise ValueError("sample size is too short to use selected regression component")
    d = np.diff(x)
    xdall = _lagmat_both_in(d, maxlag)
    nobs = xdall.shape[0]
    xdall[:, 0] = x[-nobs - 1: -1]
    y = d[-nobs:]
    full = np.column_stack([np.ones(nobs), xdall])          # [const, level, dlag1..]
    if autolag is None:
        usedlag = maxlag
    elif autolag == "t-stat":
        usedlag = maxlag
        for ncol in range(2 + maxlag, 1, -1):
            beta, ssr, xtxi = _ols(y, full[:, :ncol])
            tlast = beta[-1] / np.sqrt(ssr / (nobs - ncol) * xtxi[-1, -1])
            usedlag = ncol - 2
            if abs(tlast) >= 1.6448536269514722:
                break
    else:
        best = None
        for ncol in range(2, 2 + maxlag + 1):
            _, ssr, _ = _ols(y, full[:, :ncol])
            llf = -nobs / 2.0 * np.log(2 * np.pi) - nobs / 2.0 * np.log(ssr / nobs) - nobs / 2.0
            pen = 2.0 * ncol if autolag == "aic" else np.log(nobs) * ncol
            ic = -2.0 * llf + pen
            if best is None or (ic, ncol) < best:
                best = (ic, ncol)
        usedlag = best[1] - 2
    xdall = _lagmat_both_in(d, usedlag)
    nobs = xdall.shape[0]
    xdall[:, 0] = x[-nobs - 1: -1]
    y = d[-nobs:]
    X = np.column_stack([xdall[:, : usedlag + 1], np.ones(nobs)])
    beta, ssr, xtxi = _ols(y, X)
    scale = ssr / (nobs - X.shape[1])
    stat = beta[0] / np.sqrt(scale * xtxi[0, 0])
    return float(stat), mackinnonp_c(float(stat)), usedlag


# ----------------------------------------------------------------------------- AutoReg
class _Fit:
    def __init__(self, params):
        self.params = params


class AutoReg:
    """AutoReg(x, lags=k, trend='c').fit().params == OLS of x[t] on [1, x[t-1..t-k]] (conditional MLE)."""

    def __init__(self, endog, lags, trend="c"):
        if trend != "c":
            raise NotImplementedError
        x = np.asarray(endog, dtype=np.float64)
        k = int(lags)
        n = x.shape[0]
        if k >= n:
            raise ValueError("maxlag should be < nobs")
        rows = n - k
        if rows < k + 1:
            raise ValueError("The model specification cannot be estimated: more regressors than data points")
        self._y = x[k:]
        self._X = np.column_stack([np.ones(rows)] + [x[k - j: n - j] for j in range(1, k + 1)])

    def fit(self):
        return _Fit(np.linalg.pinv(self._X) @ self._y)


# ----------------------------------------------------------------------------- pywt.cwt (mexh)
def mexh_int_psi():
    """pywt.integrate_wavelet(ContinuousWavelet('mexh'), precision=10)."""
    t = np.linspace(-8.0, 8.0, 1024)
    psi = (1.0 - t ** 2) * np.exp(-(t ** 2) / 2.0) * 2.0 / (math.sqrt(3.0) * math.sqrt(math.sqrt(math.pi)))
    step = t[1] - t[0]
    return np.cumsum(psi) * step, t


def mexh_scaled_kernel(scale):
    """The reversed, resampled integrated wavelet pywt.cwt convolves the data with at `scale`."""
    int_psi, t = mexh_int_psi()
    step = t[1] - t[0]
    j = np.arange(scale * (t[-1] - t[0]) + 1) / (scale * step)
    j = j.astype(int)
    if j[-1] >= int_psi.size:
        j = np.extract(j < int_psi.size, j)
    return int_psi[j][::-1]


def cwt(data, scales, wavelet="mexh", **_):
    if wavelet != "mexh":
        raise NotImplementedError
    data = np.asarray(data, dtype=np.float64)
    out = np.empty((len(scales), data.shape[0]))
    for i, scale in enumerate(scales):
        ker = mexh_scaled_kernel(scale)
        conv = np.convolve(data, ker)
        coef = -np.sqrt(scale) * np.diff(conv)
        dd = (coef.shape[-1] - data.shape[-1]) / 2.0
        if dd > 0:
            coef = coef[int(np.floor(dd)): -int(np.ceil(dd))]
        elif dd < 0:
            raise ValueError("Selected scale of {} too small.".format(scale))
        out[i] = coef
    return out, None


# ----------------------------------------------------------------------------- statsmodels.stats.multitest.multipletests
def multipletests(pvals, alpha=0.05, method="fdr_bh", **_):
    """multipletests(pvals, alpha, "fdr_bh" | "fdr_by") -> (reject, pvals_corrected, None, None): the Benjamini-Hochberg
    step-up procedure (fdrcorrection, method "indep") and its Benjamini-Yekutieli variant ("negcorr": thresholds divided by
    sum_{i<=m} 1/i).  Call site: tsfresh/feature_selection/relevance.py:347-351 (only element [0] is used)."""
    if method not in ("fdr_bh", "fdr_by"):
        raise NotImplementedError(method)
    p = np.asarray(pvals, dtype=np.float64)
    m = len(p)
    order = np.argsort(p)
    ps = p[order]
    ecdf = np.arange(1, m + 1) / float(m)
    if method == "fdr_by":
        ecdf = ecdf / np.sum(1.0 / np.arange(1, m + 1))
    reject = ps <= ecdf * alpha
    if reject.any():
        reject[:np.max(np.nonzero(reject)[0]) + 1] = True
    corrected_raw = ps / ecdf
    corrected = np.minimum.accumulate(corrected_raw[::-1])[::-1]
    corrected[corrected > 1] = 1
    out_r, out_c = np.empty(m, dtype=bool), np.empty(m)
    out_r[order] = reject
    out_c[order] = corrected
    return out_r, out_c, None, None
