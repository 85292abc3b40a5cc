// tsfx_math.cuh -- scalar float64 routines used by single lanes inside the warp kernels.
// Plain C++ (compiles with g++ for the host-side unit test tests/test_host_math.py and with nvcc for
// the device).  Nothing here touches memory except through the pointers it is given.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TSFX_HD __host__ __device__ __forceinline__
#define TSFX_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define TSFX_HD inline
#define TSFX_HD_NOINLINE inline
#endif

namespace tsfx {

TSFX_HD double m_nan() {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
}

// ------------------------------------------------------------------ regularised incomplete beta
// Continued fraction (modified Lentz) for I_x(a,b); relative accuracy ~1e-14.
TSFX_HD double m_betacf(double a, double b, double x) {
    const double TINY = 1e-300, EPS = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < TINY) d = TINY;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 400; ++m) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < TINY) d = TINY;
        c = 1.0 + aa / c;
        if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < TINY) d = TINY;
        c = 1.0 + aa / c;
        if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < EPS) break;
    }
    return h;
}

TSFX_HD double m_incbeta(double a, double b, double x) {
    if (!(x >= 0.0) || !(x <= 1.0)) return m_nan();
    if (x == 0.0) return 0.0;
    if (x == 1.0) return 1.0;
    double lbt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x);
    double bt = exp(lbt);
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * m_betacf(a, b, x) / a;
    return 1.0 - bt * m_betacf(b, a, 1.0 - x) / b;
}

// 2 * stdtr(df, -|t|): the two-sided p-value scipy.stats.linregress reports (_stats_py.py linregress).
TSFX_HD_NOINLINE double m_student_two_sided(double t, double df) {
    if (!(df > 0.0) || t != t) return m_nan();
    double t2 = t * t;
    if (isinf(t2)) return 0.0;
    return m_incbeta(0.5 * df, 0.5, df / (df + t2));
}

TSFX_HD double m_norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }

// MacKinnon (1994) approximate p-value for the ADF statistic, regression "c", N=1
// (statsmodels.tsa.adfvalues.mackinnonp as called by adfuller; oracle/thirdparty.py mackinnonp_c).
TSFX_HD double m_mackinnon_p_c(double stat) {
    if (stat != stat) return m_nan();
    if (stat > 2.74) return 1.0;
    if (stat < -18.83) return 0.0;
    double z;
    if (stat <= -1.61) z = 2.1659 + stat * (1.4412 + stat * 3.8269e-2);
    else z = 1.7339 + stat * (9.3202e-1 + stat * (-1.2745e-1 + stat * (-1.0368e-2)));
    return m_norm_cdf(z);
}

// ------------------------------------------------------------------ linregress finishing step
struct LinReg { double rvalue, intercept, slope, stderr_, tstat, df; };

// scipy.stats.linregress from the averaged centred sums (ssxm, ssym, ssxym), the means and n.
TSFX_HD LinReg m_linregress(double n, double xmean, double ymean, double ssxm, double ssym, double ssxym) {
    LinReg R;
    if (n < 2.0) { R.rvalue = R.intercept = R.slope = R.stderr_ = R.tstat = m_nan(); R.df = n - 2.0; return R; }
    double r;
    if (ssxm == 0.0 || ssym == 0.0) r = (ssxym == 0.0) ? m_nan() : 0.0;
    else {
        r = ssxym / sqrt(ssxm * ssym);
        if (r > 1.0) r = 1.0;
        if (r < -1.0) r = -1.0;
    }
    R.rvalue = r;
    R.slope = ssxym / ssxm;
    R.intercept = ymean - R.slope * xmean;
    double df = n - 2.0;
    double t = r * sqrt(df / ((1.0 - r + 1e-20) * (1.0 + r + 1e-20)));
    R.tstat = t;          // the p-value (an incomplete-beta evaluation) is formed only when asked for
    R.df = df;
    // df == 0 (two points): r is +-1 up to rounding, so scipy's (1 - r^2) * ssym / ssxm / 0 is 0/0 = NaN
    // or tiny/0 = inf depending on the last bit of r; the mathematically exact value is returned here.
    R.stderr_ = (df == 0.0) ? m_nan() : sqrt((1.0 - r * r) * ssym / ssxm / df);
    return R;
}

TSFX_HD double m_linreg_pick(const LinReg& R, int attr) {
    switch (attr) {
        case 0: return m_student_two_sided(R.tstat, R.df);
        case 1: return R.rvalue;
        case 2: return R.intercept;
        case 3: return R.slope;
        default: return R.stderr_;
    }
}

// ------------------------------------------------------------------ dense SPD solves (row-major, lda)
// In-place lower Cholesky of the leading n x n block.  Returns false on a non-positive pivot.
TSFX_HD bool m_cholesky(double* A, int n, int lda) {
    for (int j = 0; j < n; ++j) {
        double d = A[j * lda + j];
        for (int k = 0; k < j; ++k) d -= A[j * lda + k] * A[j * lda + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        A[j * lda + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * lda + j];
            for (int k = 0; k < j; ++k) s -= A[i * lda + k] * A[j * lda + k];
            A[i * lda + j] = s / d;
        }
    }
    return true;
}
// forward substitution L z = b (in place in b)
TSFX_HD void m_forward(const double* L, int n, int lda, double* b) {
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * lda + k] * b[k];
        b[i] = s / L[i * lda + i];
    }
}
// back substitution L^T x = z (in place)
TSFX_HD void m_backward(const double* L, int n, int lda, double* b) {
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * lda + i] * b[k];
        b[i] = s / L[i * lda + i];
    }
}

// ------------------------------------------------------------------ Levinson-Durbin partial autocorrelation
// acv[0..nlags] autocovariances; out[0..nlags] = pacf (out[0] = 1).  work: 2*(nlags+1) doubles.
// (statsmodels levinson_durbin(..., isacov=True)[2]; oracle/thirdparty.py levinson_durbin_pacf)
TSFX_HD void m_levinson_pacf(const double* acv, int nlags, double* out, double* work) {
    double* prev = work;
    double* cur = work + (nlags + 1);
    out[0] = 1.0;
    if (nlags < 1) return;
    double phi = acv[1] / acv[0];
    double sig = acv[0] - phi * acv[1];
    prev[1] = phi;
    out[1] = phi;
    for (int k = 2; k <= nlags; ++k) {
        double num = acv[k];
        for (int j = 1; j < k; ++j) num -= prev[j] * acv[k - j];
        double pk = num / sig;
        for (int j = 1; j < k; ++j) cur[j] = prev[j] - pk * prev[k - j];
        cur[k] = pk;
        sig = sig * (1.0 - pk * pk);
        out[k] = pk;
        for (int j = 1; j <= k; ++j) prev[j] = cur[j];
    }
}

// ------------------------------------------------------------------ cubic: largest real part of the roots
// np.max(np.real(np.roots([c0,c1,c2,c3]))) (feature_calculators.py:2163): leading zeros are stripped,
// trailing zeros contribute roots at 0.
TSFX_HD double m_quadratic_max_real(double a, double b, double c) {
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return -b / (2.0 * a);
    double q = -0.5 * (b + (b >= 0.0 ? 1.0 : -1.0) * sqrt(disc));
    double r1 = q / a, r2 = (q != 0.0) ? c / q : -b / a - r1;
    return r1 > r2 ? r1 : r2;
}

TSFX_HD double m_poly3_max_real_root(double c0, double c1, double c2, double c3) {
    if (c0 != c0 || c1 != c1 || c2 != c2 || c3 != c3) return m_nan();
    if (isinf(c0) || isinf(c1) || isinf(c2) || isinf(c3)) return m_nan();
    // strip leading zeros
    if (c0 == 0.0) {
        if (c1 == 0.0) {
            if (c2 == 0.0) return m_nan();          // np.roots -> [] -> np.max raises ValueError -> NaN
            return -c3 / c2;
        }
        if (c3 == 0.0) {                             // roots: 0 and -c2/c1
            double r = -c2 / c1;
            return r > 0.0 ? r : 0.0;
        }
        return m_quadratic_max_real(c1, c2, c3);
    }
    double a = c1 / c0, b = c2 / c0, c = c3 / c0;
    // depressed cubic t^3 + p t + q, x = t - a/3
    double a3 = a / 3.0;
    double p = b - a * a3;
    double q = 2.0 * a3 * a3 * a3 - a3 * b + c;
    double disc = 0.25 * q * q + p * p * p / 27.0;
    double best;
    if (disc > 0.0) {               // one real root r, complex pair with real part -(r_t)/2
        double sq = sqrt(disc);
        double u = cbrt(-0.5 * q + sq), v = cbrt(-0.5 * q - sq);
        double t = u + v;
        // polish the real root with Newton on the original cubic
        double x = t - a3;
        for (int it = 0; it < 3; ++it) {
            double f = ((x + a) * x + b) * x + c;
            double fp = (3.0 * x + 2.0 * a) * x + b;
            if (fp != 0.0) x -= f / fp;
        }
        double pair_re = -0.5 * (x + a);            // sum of roots = -a
        best = x > pair_re ? x : pair_re;
    } else {                        // three real roots
        double m = 2.0 * sqrt(-p / 3.0);
        double arg = (p != 0.0) ? (3.0 * q / (p * m)) : 0.0;
        if (arg > 1.0) arg = 1.0;
        if (arg < -1.0) arg = -1.0;
        double th = acos(arg) / 3.0;
        double x = m * cos(th) - a3;                 // k = 0 gives the largest root
        for (int it = 0; it < 3; ++it) {
            double f = ((x + a) * x + b) * x + c;
            double fp = (3.0 * x + 2.0 * a) * x + b;
            if (fp != 0.0) x -= f / fp;
        }
        best = x;
    }
    return best;
}

// ------------------------------------------------------------------ cubic least squares (np.polyfit deg 3)
// Fits y ~ c0 x^3 + c1 x^2 + c2 x + c3 to k points the way np.polyfit does: Vandermonde columns scaled
// to unit 2-norm, least squares, minimum-norm solution when k < 4.  Householder-free: modified
// Gram-Schmidt on the scaled columns (k >= 4) or on the rows (k < 4).  x,y are not modified.
// Returns false when the problem is numerically rank deficient beyond what polyfit would also flag.
TSFX_HD bool m_polyfit3(const double* x, const double* y, int k, double* coef /*4*/) {
    if (k <= 0) return false;
    double scale[4] = {0, 0, 0, 0};
    for (int i = 0; i < k; ++i) {
        double v = x[i], v2 = v * v, v3 = v2 * v;
        scale[0] += v3 * v3; scale[1] += v2 * v2; scale[2] += v * v; scale[3] += 1.0;
    }
    for (int j = 0; j < 4; ++j) scale[j] = sqrt(scale[j]);
    if (k >= 4) {
        // normal equations on the scaled columns solved by Cholesky; G is 4x4, well conditioned after
        // scaling for the binned-mean abscissae this is used on.  One step of iterative refinement.
        double G[16], rhs[4];
        for (int a = 0; a < 16; ++a) G[a] = 0.0;
        for (int a = 0; a < 4; ++a) rhs[a] = 0.0;
        for (int i = 0; i < k; ++i) {
            double v = x[i];
            double col[4] = {v * v * v / scale[0], v * v / scale[1], v / scale[2], 1.0 / scale[3]};
            for (int a = 0; a < 4; ++a) {
                rhs[a] += col[a] * y[i];
                for (int b = 0; b <= a; ++b) G[a * 4 + b] += col[a] * col[b];
            }
        }
        double L[16];
        for (int a = 0; a < 16; ++a) L[a] = G[a];
        if (!m_cholesky(L, 4, 4)) return false;
        double sol[4] = {rhs[0], rhs[1], rhs[2], rhs[3]};
        m_forward(L, 4, 4, sol);
        m_backward(L, 4, 4, sol);
        for (int it = 0; it < 2; ++it) {            // refinement with residual formed from the data
            double r[4] = {0, 0, 0, 0};
            for (int i = 0; i < k; ++i) {
                double v = x[i];
                double col[4] = {v * v * v / scale[0], v * v / scale[1], v / scale[2], 1.0 / scale[3]};
                double e = y[i] - (col[0] * sol[0] + col[1] * sol[1] + col[2] * sol[2] + col[3] * sol[3]);
                for (int a = 0; a < 4; ++a) r[a] += col[a] * e;
            }
            m_forward(L, 4, 4, r);
            m_backward(L, 4, 4, r);
            for (int a = 0; a < 4; ++a) sol[a] += r[a];
        }
        for (int a = 0; a < 4; ++a) coef[a] = sol[a] / scale[a];
        return true;
    }
    // k < 4: minimum-norm solution c = A^T (A A^T)^-1 y on the scaled k x 4 matrix
    double A[12], H[9], z[3];
    for (int i = 0; i < k; ++i) {
        double v = x[i];
        A[i * 4 + 0] = v * v * v / scale[0]; A[i * 4 + 1] = v * v / scale[1];
        A[i * 4 + 2] = v / scale[2];         A[i * 4 + 3] = 1.0 / scale[3];
        z[i] = y[i];
    }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            for (int c = 0; c < 4; ++c) s += A[i * 4 + c] * A[j * 4 + c];
            H[i * 3 + j] = s;
        }
    if (!m_cholesky(H, k, 3)) return false;
    m_forward(H, k, 3, z);
    m_backward(H, k, 3, z);
    for (int c = 0; c < 4; ++c) {
        double s = 0.0;
        for (int i = 0; i < k; ++i) s += A[i * 4 + c] * z[i];
        coef[c] = s / scale[c];
    }
    return true;
}

// ------------------------------------------------------------------ numpy linear quantile on a sorted array
// np.quantile(method="linear") incl. numpy's two-sided lerp (a + (b-a)*t for t < 0.5, b - (b-a)*(1-t) else).
template <typename T>
TSFX_HD double m_quantile_sorted(const T* s, int n, double q) {
    double pos = q * (double)(n - 1);
    double fl = floor(pos);
    int lo = (int)fl;
    if (lo < 0) lo = 0;
    if (lo > n - 1) lo = n - 1;
    int hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
    double t = pos - fl;
    double a = (double)s[lo], b = (double)s[hi];
    double d = b - a;
    if (t >= 0.5) return b - d * (1.0 - t);
    return a + d * t;
}

}  // namespace tsfx
