// k_la.cu -- kernel group LA: small dense least-squares problems per series.
//   ar_coefficient          (feature_calculators.py:1459-1507; statsmodels AutoReg(lags=k, trend="c") OLS)
//   augmented_dickey_fuller (feature_calculators.py:499-544; statsmodels adfuller(regression="c"),
//                            autolag AIC / BIC / None) -- restated in oracle/thirdparty.py
//
// One warp per series.  Normal equations are formed cooperatively (one Gram entry per lane, looping over
// the rows) on mean-centred regressors, then solved by lane 0 with a float64 Cholesky.  The nested ADF
// lag search needs ONE factorisation: with columns ordered [const, level, dlag1, dlag2, ...] the
// residual sum of squares of the model using the first q columns is y'y - sum_{i<q} z_i^2, z = L^-1 X'y.
#include <algorithm>

#include "tsfx_common.cuh"
#include "tsfx_kernels.h"
#include "tsfx_math.cuh"

namespace tsfx {

__host__ __device__ inline int adf_maxlag(int n) {
    // ceil(12 * (n/100)^(1/4)); sqrt(sqrt()) is correctly rounded, so the perfect-fourth-power lengths
    // (100, 1600, 8100, ...) land exactly on the integer like a correctly rounded pow() does
    int m = (int)ceil(12.0 * sqrt(sqrt((double)n / 100.0)));
    int cap = n / 2 - 2;
    return m < cap ? m : cap;
}

// regressor c of row t for the ADF design: 0 const, 1 level (centred), c>=2: d[t-(c-1)]
__device__ __forceinline__ double adf_reg(int c, int t, const double* lev, const double* dx) {
    if (c == 0) return 1.0;
    if (c == 1) return lev[t];
    return dx[t - (c - 1)];
}

// Gram of the columns `cols[0..q)` (indices into the ADF regressor set) over rows t0..t1-1, plus X'y and y'y.
// Layout: G row-major q x q (lower triangle filled), b[q], yy.  One entry per lane and round.
__device__ __forceinline__ void adf_gram(const int* colmap, int q, int t0, int t1, const double* lev, const double* dx,
                                         double* G, double* b, double* yy, int lane) {
    const int ntri = q * (q + 1) / 2;
    const int total = ntri + q + 1;
    for (int e = lane; e < total; e += 32) {
        int a = 0, c = 0, kind;                 // kind 0: G[a][c], 1: b[a], 2: yy
        if (e < ntri) {
            // invert e = a(a+1)/2 + c, c <= a
            a = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
            while (a * (a + 1) / 2 > e) --a;
            while ((a + 1) * (a + 2) / 2 <= e) ++a;
            c = e - a * (a + 1) / 2;
            kind = 0;
        } else if (e < ntri + q) { a = e - ntri; kind = 1; }
        else kind = 2;
        const int ca = colmap ? colmap[a] : a, cc = colmap ? colmap[c] : c;
        double acc = 0.0;
        for (int t = t0; t < t1; ++t) {
            double y = dx[t];
            double va = (kind == 2) ? y : adf_reg(ca, t, lev, dx);
            double vb = (kind == 0) ? adf_reg(cc, t, lev, dx) : y;
            acc = fma(va, vb, acc);
        }
        if (kind == 0) G[a * q + c] = acc;
        else if (kind == 1) b[a] = acc;
        else *yy = acc;
    }
    __syncwarp();
}

// ---- warp-cooperative dense SPD kernels (row-major lower triangle in shared memory) -----------------
// In-place Cholesky; returns the number of columns factorised before a non-positive pivot (q if none).
__device__ __forceinline__ int warp_cholesky(double* G, int q, int lda, int lane) {
    for (int j = 0; j < q; ++j) {
        double s = 0.0;
        for (int k = lane; k < j; k += 32) { double v = G[j * lda + k]; s = fma(v, v, s); }
        double d = G[j * lda + j] - wsum(s);
        if (!(d > 0.0)) return j;
        d = sqrt(d);
        __syncwarp();
        if (lane == 0) G[j * lda + j] = d;
        for (int i = j + 1 + lane; i < q; i += 32) {
            double t = G[i * lda + j];
            for (int k = 0; k < j; ++k) t = fma(-G[i * lda + k], G[j * lda + k], t);
            G[i * lda + j] = t / d;
        }
        __syncwarp();
    }
    return q;
}
// L z = b in place (column oriented: after z_k is known every lane retires it from its own row)
__device__ __forceinline__ void warp_forward(const double* L, int q, int lda, double* b, int lane) {
    for (int k = 0; k < q; ++k) {
        double zk = b[k] / L[k * lda + k];
        __syncwarp();
        if (lane == 0) b[k] = zk;
        for (int i = k + 1 + lane; i < q; i += 32) b[i] = fma(-L[i * lda + k], zk, b[i]);
        __syncwarp();
    }
}
// L^T x = z in place
__device__ __forceinline__ void warp_backward(const double* L, int q, int lda, double* b, int lane) {
    for (int i = q - 1; i >= 0; --i) {
        double xi = b[i] / L[i * lda + i];
        __syncwarp();
        if (lane == 0) b[i] = xi;
        for (int k = lane; k < i; k += 32) b[k] = fma(-L[i * lda + k], xi, b[k]);
        __syncwarp();
    }
}

template <int WPC>
__global__ void __launch_bounds__(WPC * 32) k_la(LaArgs A, int pmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* base = smem_raw + (size_t)warp * A.bytes_per_warp;
    double* xc = reinterpret_cast<double*>(base);          // npad : centred series (level)
    double* dx = xc + A.npad;                              // npad : first differences
    double* G = dx + A.npad;                               // pmax*pmax
    double* bvec = G + pmax * pmax;                        // pmax
    double* res = bvec + pmax;                             // 8 : staged results
    int* colmap = reinterpret_cast<int*>(res + 8);         // pmax ints (pmax is even)
    float* xs = reinterpret_cast<float*>(colmap + ((pmax + 3) & ~3));
    const int64_t warps_total = (int64_t)gridDim.x * WPC;

    for (int64_t s = (int64_t)blockIdx.x * WPC + warp; s < A.R.n_series; s += warps_total) {
        const int n = load_series(A.R, s, xs, lane);
        const Moments M = moments(xs, n, xc, lane);
        for (int i = lane; i + 1 < n; i += 32) dx[i] = (double)xs[i + 1] - (double)xs[i];
        __syncwarp();
        double* orow = A.out + (size_t)s * A.ncols;
        int ar_k = -1; bool ar_ok = false;
        int adf_mode = -1;

        for (int j = 0; j < A.nd; ++j) {
            const Desc d = A.descs[j];
            double r = dnan();
            if (d.calc == TSFX_AR_COEFFICIENT) {
                const int k = d.i1, p = d.i0;
                if (k != ar_k) {
                    ar_k = k;
                    const int rows = n - k;
                    ar_ok = (k < n) && (rows >= k + 1);
                    if (ar_ok) {
                        const int q = k + 1;
                        // columns: 0 const, j>=1: xc[t-j]; target xc[t]; rows t = k..n-1
                        const int ntri = q * (q + 1) / 2, total = ntri + q;
                        for (int e = lane; e < total; e += 32) {
                            int a, c;
                            bool rhs = e >= ntri;
                            if (!rhs) {
                                a = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
                                while (a * (a + 1) / 2 > e) --a;
                                while ((a + 1) * (a + 2) / 2 <= e) ++a;
                                c = e - a * (a + 1) / 2;
                            } else { a = e - ntri; c = 0; }
                            double acc = 0.0;
                            for (int t = k; t < n; ++t) {
                                double va = a == 0 ? 1.0 : xc[t - a];
                                double vb = rhs ? xc[t] : (c == 0 ? 1.0 : xc[t - c]);
                                acc = fma(va, vb, acc);
                            }
                            if (rhs) bvec[a] = acc; else G[a * q + c] = acc;
                        }
                        __syncwarp();
                        int good = 0;
                        if (M.vmax == M.vmin) {
                            // rank-one design (constant series): numpy pinv's minimum-norm solution
                            double cst = M.vmin, sc = cst / (1.0 + (double)k * cst * cst);
                            for (int a = lane; a < q; a += 32) bvec[a] = (a == 0) ? sc : sc * cst;
                            good = 2;
                        } else if (warp_cholesky(G, q, q, lane) == q) {
                            warp_forward(G, q, q, bvec, lane);
                            warp_backward(G, q, q, bvec, lane);
                            // undo the centring: const = c~ + mean * (1 - sum phi)
                            double sphi = 0.0;
                            for (int a = 1 + lane; a < q; a += 32) sphi += bvec[a];
                            sphi = wsum(sphi);
                            if (lane == 0) bvec[0] = bvec[0] + M.mean * (1.0 - sphi);
                            good = 1;
                        }
                        __syncwarp();
                        if (!good) { ar_ok = true; for (int a = lane; a <= k; a += 32) bvec[a] = dnan(); __syncwarp(); }
                    }
                }
                if (p > k) r = dnan();
                else if (!ar_ok) r = (p < k) ? dnan() : 0.0;       // params = [nan]*k ; index k -> IndexError -> 0
                else r = bvec[p];
            } else if (d.calc == TSFX_AUGMENTED_DICKEY_FULLER) {
                if (adf_mode != d.i0) {
                    adf_mode = d.i0;
                    __syncwarp();
                    double stat = dnan(), pval = dnan(), ulag = dnan();
                    const int M0 = adf_maxlag(n);
                    if (M.vmax != M.vmin && M0 >= 0) {
                        const int nd_ = n - 1;
                        int used = M0;
                        if (adf_mode != TSFX_AUTOLAG_NONE) {
                            const int p = M0 + 2, t0 = M0, nobs = nd_ - M0;
                            double* yy = res + 7;
                            adf_gram(nullptr, p, t0, nd_, xc, dx, G, bvec, yy, lane);
                            int best_q = 2;
                            {
                                // factorise as far as the pivots stay positive (the models are nested)
                                const int okq = warp_cholesky(G, p, p, lane);
                                double ssr = *yy, best_ic = 0.0;
                                bool have = false;
                                const double dnobs = (double)nobs;
                                for (int i = 0; i < okq; ++i) {
                                    double z = bvec[i] / G[i * p + i];
                                    __syncwarp();
                                    for (int r2 = i + 1 + lane; r2 < okq; r2 += 32) bvec[r2] = fma(-G[r2 * p + i], z, bvec[r2]);
                                    __syncwarp();
                                    ssr -= z * z;
                                    const int q = i + 1;
                                    if (q >= 2) {
                                        double llf = -dnobs / 2.0 * log(2.0 * 3.14159265358979323846) -
                                                     dnobs / 2.0 * log(ssr / dnobs) - dnobs / 2.0;
                                        double pen = (adf_mode == TSFX_AUTOLAG_AIC) ? 2.0 * (double)q : log(dnobs) * (double)q;
                                        double ic = -2.0 * llf + pen;
                                        if (!have || ic < best_ic) { have = true; best_ic = ic; best_q = q; }
                                    }
                                }
                            }
                            best_q = __shfl_sync(FULL, best_q, 0);
                            used = best_q - 2;
                            __syncwarp();
                        }
                        // final regression on the longer sample, columns [const, dlag1..dlagU, level]
                        const int q = used + 2, t0 = used, nobs = nd_ - used;
                        for (int c = lane; c < q; c += 32) colmap[c] = (c == 0) ? 0 : (c == q - 1 ? 1 : c + 1);
                        __syncwarp();
                        double* yy = res + 7;
                        adf_gram(colmap, q, t0, nd_, xc, dx, G, bvec, yy, lane);
                        if (warp_cholesky(G, q, q, lane) == q) {
                            warp_forward(G, q, q, bvec, lane);
                            double ssr = 0.0;
                            for (int i = lane; i < q; i += 32) ssr = fma(bvec[i], bvec[i], ssr);
                            ssr = *yy - wsum(ssr);
                            double s2 = ssr / (double)(nobs - q);
                            stat = bvec[q - 1] / sqrt(s2);
                            pval = m_mackinnon_p_c(stat);
                            ulag = (double)used;
                        }
                        __syncwarp();
                        if (lane == 0) { res[0] = stat; res[1] = pval; res[2] = ulag; }
                    } else if (lane == 0) { res[0] = stat; res[1] = pval; res[2] = ulag; }
                    __syncwarp();
                }
                r = (d.attr >= 0 && d.attr <= 2) ? res[d.attr] : dnan();
            }
            if (lane == 0) orow[d.col] = r;
        }
        __syncwarp();
    }
}

cudaError_t launch_la(const LaArgs& A0, int max_len, cudaStream_t st, int sm_count) {
    LaArgs A = A0;
    A.npad = (max_len + 3) & ~3;
    int pmax = std::max(adf_maxlag(max_len) + 2, 34);       // 33 covers ar k <= 32
    pmax = (pmax + 1) & ~1;
    size_t per = (size_t)A.npad * 16 + (size_t)pmax * pmax * 8 + (size_t)pmax * 8 + 64 + (size_t)((pmax + 3) & ~3) * 4 + (size_t)A.npad * 4;
    per = (per + 15) & ~(size_t)15;
    A.bytes_per_warp = (int)per;
    if (per > 227 * 1024) return cudaErrorInvalidConfiguration;
    int wpc = (int)std::min<size_t>(8, std::max<size_t>(1, 100 * 1024 / per));
    wpc = wpc >= 8 ? 8 : wpc >= 4 ? 4 : wpc >= 2 ? 2 : 1;
    size_t smem = per * wpc;
    int64_t cap = (int64_t)sm_count * 16;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((A.R.n_series + wpc - 1) / wpc, cap));
#define TSFX_LAUNCH(W)                                                                                 \
    {                                                                                                  \
        cudaError_t e = cudaFuncSetAttribute(k_la<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                \
        k_la<W><<<grid, W * 32, smem, st>>>(A, pmax);                                                  \
    }
    switch (wpc) {
        case 8: TSFX_LAUNCH(8) break;
        case 4: TSFX_LAUNCH(4) break;
        case 2: TSFX_LAUNCH(2) break;
        default: TSFX_LAUNCH(1) break;
    }
#undef TSFX_LAUNCH
    return cudaGetLastError();
}

}  // namespace tsfx
