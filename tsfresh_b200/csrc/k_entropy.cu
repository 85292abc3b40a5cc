// k_entropy.cu -- kernel group ENTROPY: sample_entropy (feature_calculators.py:1701-1754) and
// approximate_entropy with m = 2 (feature_calculators.py:1759-1805) -- the O(n^2) "class Q" rows.
//
// Both need, for a tolerance tau and every template start i, the number of template starts j whose
// Chebyshev distance is <= tau, for templates of length 2 (i, j in [0, n-2]) and of length 3
// (i, j in [0, n-3]).  One warp per series; up to NT = 6 tolerances share one pass so the distances are formed
// once.  Differences are float64 of float32-origin values, i.e. the very same IEEE operations numpy performs,
// so the counts are bit-identical to the reference's.  Two formulations of the counting:
//   * bit tiles (default, entropy_bittile): lane = row i, 32-column bit words per tolerance, counts by popcount
//     of three shifted rows -- ~17 warp instructions per 32 pair tests and 6 tolerances;
//   * pair sweep (TSFX_ENTROPY=pairs, entropy_sweep): lane = row i, sequential sweep over j with the three
//     neighbouring samples in registers -- ~35; kept for A/B measurements and as a cross-check.
#include <algorithm>
#include <cstdlib>

#include "tsfx_common.cuh"
#include "tsfx_kernels.h"

namespace tsfx {

// c += (m <= tau) as DSETP + one predicated IADD (the compiler's own choice is VIADD + predicated MOV)
__device__ __forceinline__ void count_le(int& c, double m, double tau) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.f64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "d"(m), "d"(tau));
}
// max of two non-NaN magnitudes without fmax()'s NaN handling (DSETP + SEL instead of DSETP.MAX/FSEL/SEL/LOP3)
__device__ __forceinline__ double max_nn(double a, double b) { return a > b ? a : b; }
// |x| as one integer AND on the high word (keeps the half-rate FP64 pipe for the subtractions and compares)
__device__ __forceinline__ double abs_bits(double x) {
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}

template <int NT>
__device__ __forceinline__ void entropy_sweep(const double* xd, int n, const double (&tau)[NT], double (&sum_ln2)[NT],
                                              double (&sum_ln3)[NT], double (&sumB)[NT], double (&sumA)[NT], int lane) {
#pragma unroll
    for (int t = 0; t < NT; ++t) { sum_ln2[t] = 0.0; sum_ln3[t] = 0.0; sumB[t] = 0.0; sumA[t] = 0.0; }
    const int n2 = n - 1, n3 = n - 2;          // number of length-2 / length-3 templates
    if (n2 <= 0) return;
    const double inv2 = 1.0 / (double)n2, inv3 = n3 > 0 ? 1.0 / (double)n3 : 0.0;
    for (int r0 = 0; r0 < n2; r0 += 32) {
        const int i = r0 + lane;
        const bool v2 = i < n2, v3 = i < n3;
        const double a0 = v2 ? xd[i] : 0.0, a1 = v2 ? xd[i + 1] : 0.0, a2 = v3 ? xd[i + 2] : 0.0;
        int c2[NT], c3[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) { c2[t] = 0; c3[t] = 0; }
        double b0 = xd[0], b1 = xd[1];
        for (int j = 0; j < n3; ++j) {
            const double b2 = xd[j + 2];
            const double m2 = max_nn(abs_bits(a0 - b0), abs_bits(a1 - b1));
            const double m3 = max_nn(m2, abs_bits(a2 - b2));
#pragma unroll
            for (int t = 0; t < NT; ++t) { count_le(c2[t], m2, tau[t]); count_le(c3[t], m3, tau[t]); }
            b0 = b1; b1 = b2;
        }
        {   // last length-2 template j = n2 - 1
            const double m2 = max_nn(abs_bits(a0 - b0), abs_bits(a1 - b1));
#pragma unroll
            for (int t = 0; t < NT; ++t) count_le(c2[t], m2, tau[t]);
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (v2) { sum_ln2[t] += log((double)c2[t] * inv2); sumB[t] += (double)(c2[t] - 1); }
            if (v3) { sum_ln3[t] += log((double)c3[t] * inv3); sumA[t] += (double)(c3[t] - 1); }
        }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        sum_ln2[t] = wsum(sum_ln2[t]);
        sum_ln3[t] = wsum(sum_ln3[t]);
        sumB[t] = wsum(sumB[t]);
        sumA[t] = wsum(sumA[t]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Bit-tile formulation (default).  For a tolerance tau let R_i be the bit row R_i[j] = [ |x_i - x_j| <= tau ].
// The template counts are then pure bit operations on three consecutive rows:
//     c2(i) = popc( R_i & (R_{i+1} >> 1) )                      c3(i) = popc( R_i & (R_{i+1} >> 1) & (R_{i+2} >> 2) )
// (samples beyond n are NaN, whose comparisons are false, so the ranges j <= n-2 / j <= n-3 need no masks).
// Lane = row i; a 32-column tile of R_i is built with one float64 subtract per pair plus one DSETP + predicated OR
// per tolerance -- the same IEEE operations numpy performs, so the counts stay bit-identical -- and rows i+1, i+2
// come from the neighbouring lanes by shuffle, which is why a row block advances by 30 rows, not 32.  Per 32 pairs
// and 6 tolerances this issues ~17 warp instructions (7 of them FP64) against ~35 (17 FP64) for the pair sweep.
__device__ __forceinline__ void or_le(unsigned& w, double d, double tau, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.f64 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(w) : "d"(d), "d"(tau), "r"(bit));
}

template <int NT>
__device__ __forceinline__ void entropy_bittile(const double* xd, const double* lnk, int n, const double (&tau)[NT],
                                                double (&sum_ln2)[NT], double (&sum_ln3)[NT], double (&sumB)[NT],
                                                double (&sumA)[NT], int lane) {
#pragma unroll
    for (int q = 0; q < NT; ++q) { sum_ln2[q] = 0.0; sum_ln3[q] = 0.0; sumB[q] = 0.0; sumA[q] = 0.0; }
    const int n2 = n - 1, n3 = n - 2;          // number of length-2 / length-3 templates
    if (n2 <= 0) return;
    const int W = (n + 31) >> 5;
    int iB[NT], iA[NT];
#pragma unroll
    for (int q = 0; q < NT; ++q) { iB[q] = 0; iA[q] = 0; }
    for (int r0 = 0; r0 < n2; r0 += 30) {
        const int i = r0 + lane;
        const double a = xd[i];                 // NaN beyond n: an all-zero row
        unsigned wp[NT], s1p[NT], s2p[NT];      // previous tile of rows i, i+1, i+2
        int c2[NT], c3[NT];
#pragma unroll
        for (int q = 0; q < NT; ++q) { wp[q] = 0u; s1p[q] = 0u; s2p[q] = 0u; c2[q] = 0; c3[q] = 0; }
        for (int t = 0; t <= W; ++t) {
            unsigned wn[NT];
#pragma unroll
            for (int q = 0; q < NT; ++q) wn[q] = 0u;
            if (t < W) {
                const double* xt = xd + t * 32;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const double d = abs_bits(a - xt[jj]);
#pragma unroll
                    for (int q = 0; q < NT; ++q) or_le(wn[q], d, tau[q], 1u << jj);
                }
            }
#pragma unroll
            for (int q = 0; q < NT; ++q) {      // finish tile t-1 now that the bit spilling over from tile t is known
                const unsigned s1n = __shfl_down_sync(FULL, wn[q], 1), s2n = __shfl_down_sync(FULL, wn[q], 2);
                const unsigned m2 = wp[q] & __funnelshift_r(s1p[q], s1n, 1);
                const unsigned m3 = m2 & __funnelshift_r(s2p[q], s2n, 2);
                c2[q] += __popc(m2);
                c3[q] += __popc(m3);
                wp[q] = wn[q]; s1p[q] = s1n; s2p[q] = s2n;
            }
        }
        const bool v2 = lane < 30 && i < n2, v3 = lane < 30 && i < n3;
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            if (v2) { sum_ln2[q] += lnk[c2[q]]; iB[q] += c2[q] - 1; }
            if (v3) { sum_ln3[q] += lnk[c3[q]]; iA[q] += c3[q] - 1; }
        }
    }
    const double ln2 = log((double)n2), ln3 = n3 > 0 ? log((double)n3) : 0.0;
#pragma unroll
    for (int q = 0; q < NT; ++q) {              // sum_i log(c_i / N) = sum_i log(c_i) - N log(N)
        sum_ln2[q] = wsum(sum_ln2[q]) - (double)n2 * ln2;
        sum_ln3[q] = wsum(sum_ln3[q]) - (double)n3 * ln3;
        sumB[q] = wsum((double)iB[q]);
        sumA[q] = wsum((double)iA[q]);
    }
}

template <int NT>
__device__ __forceinline__ void entropy_batch(const Desc* descs, int j0, int cnt, const double* xd, int n, double sd,
                                              double* orow, int lane, const double* lnk, bool bittile) {
    double tau[NT], l2[NT], l3[NT], sB[NT], sA[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        if (t < cnt) {
            const Desc d = descs[j0 + t];
            tau[t] = (d.calc == TSFX_SAMPLE_ENTROPY) ? 0.2 * sd : d.p0 * sd;
        } else tau[t] = -1.0;
    }
    if (bittile) entropy_bittile<NT>(xd, lnk, n, tau, l2, l3, sB, sA, lane);
    else entropy_sweep<NT>(xd, n, tau, l2, l3, sB, sA, lane);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        if (t < cnt) {
            const Desc d = descs[j0 + t];
            double r;
            if (d.calc == TSFX_SAMPLE_ENTROPY) r = -log(sA[t] / sB[t]);
            else if (n <= 3) r = 0.0;                                  // N <= m + 1
            else r = fabs(l2[t] / (double)(n - 1) - l3[t] / (double)(n - 2));
            if (lane == 0) orow[d.col] = r;
        }
    }
}

template <int WPC, bool GS>
__global__ void __launch_bounds__(WPC * 32, (WPC == 4 ? 3 : 1)) k_entropy(EntropyArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* base = warp_region<GS>(smem_raw, A.gscratch, A.bytes_per_warp, WPC, warp);
    double* xd = reinterpret_cast<double*>(base);                         // xpad doubles (NaN beyond n)
    double* lnk = xd + A.xpad;                                             // log(k), k = 0..npad
    float* xs = reinterpret_cast<float*>(lnk + A.npad + 4);
    const int64_t warps_total = (int64_t)gridDim.x * WPC;

    for (int64_t s = (int64_t)blockIdx.x * WPC + warp; s < A.R.n_series; s += warps_total) {
        const int n = load_series(A.R, s, xs, lane);
        const Moments M = moments(xs, n, nullptr, lane);
        const bool bm = A.bittile != 0;
        // padding: NaN for the bit tiles (comparisons false), zeros at n, n+1 for the pair sweep
        for (int i = lane; i < A.xpad; i += 32) xd[i] = i < n ? (double)xs[i] : ((bm || i >= n + 2) ? dnan() : 0.0);
        if (bm) for (int k = lane; k <= n; k += 32) lnk[k] = log((double)k);
        __syncwarp();
        double* orow = A.out + (size_t)s * A.ncols;
        int j = 0;
        while (j < A.nd) {
            int left = A.nd - j;
            if (left >= 6) { entropy_batch<6>(A.descs, j, 6, xd, n, M.sd, orow, lane, lnk, bm); j += 6; }
            else if (left > 3) { entropy_batch<6>(A.descs, j, left, xd, n, M.sd, orow, lane, lnk, bm); j += left; }
            else if (left == 3) { entropy_batch<3>(A.descs, j, 3, xd, n, M.sd, orow, lane, lnk, bm); j += 3; }
            else if (left == 2) { entropy_batch<2>(A.descs, j, 2, xd, n, M.sd, orow, lane, lnk, bm); j += 2; }
            else { entropy_batch<1>(A.descs, j, 1, xd, n, M.sd, orow, lane, lnk, bm); j += 1; }
        }
        __syncwarp();
    }
}

cudaError_t launch_entropy(const EntropyArgs& A0, int max_len, cudaStream_t st, int sm_count) {
    EntropyArgs A = A0;
    A.npad = (max_len + 3) & ~3;
    A.xpad = ((A.npad + 2 + 31) / 32) * 32 + 32;          // NaN padding up to a whole 32-sample tile / 32-row block
    {
        static int mode = -1;
        if (mode < 0) { const char* e = getenv("TSFX_ENTROPY"); mode = (e && e[0] == 'p') ? 0 : 1; }   // "pairs" = pair sweep
        A.bittile = mode;
    }
    size_t per = (size_t)A.xpad * 8 + (size_t)(A.npad + 4) * 8 + (size_t)A.npad * 4;
    per = (per + 15) & ~(size_t)15;
    A.bytes_per_warp = (int)per;
    Geometry G;
    if (!plan_geometry(per, 64 * 1024, 4, A.R.n_series, sm_count, A.gscratch, A.gscratch_bytes, &G)) return cudaErrorInvalidConfiguration;
    A.gscratch = G.gscratch;
    TSFX_DISPATCH(k_entropy, G, st, A)
    return cudaGetLastError();
}

}  // namespace tsfx
