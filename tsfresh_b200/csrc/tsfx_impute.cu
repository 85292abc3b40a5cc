// tsfx_impute.cu -- column-wise imputation of the feature matrix on the device.
//
// Restates tsfresh/utilities/dataframe_functions.py: impute :49-78 = get_range_values_per_column :170-212
// (finite max / min / median per column; a column without any finite value -> 0 for all three) followed by
// impute_dataframe_range :104-167 (+inf -> max, -inf -> min, NaN -> median), and impute_dataframe_zero :81-101.
//
// The matrix is the row-major [rows x cols] float64 block the extraction kernels write.  One HBM-bound statistics
// sweep (warps read 32 consecutive columns of a row, 256 B, four rows in flight per thread), one radix sort per
// column that actually needs a median (it contains a NaN; every column when the caller asks for the medians), and a
// replacement sweep that skips every (row slice, column tile) the statistics found clean.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <vector>

#include "tsfx_impute.h"

namespace tsfx {

namespace {

constexpr int TILE_C = 32;      // columns per block (one warp reads one 256-byte row segment)
constexpr int TILE_R = 8;       // row lanes per block
constexpr int UNR = 8;          // rows per thread and trip (memory-level parallelism)

struct ColPartial { double vmin, vmax; long long finite, nan; };

// finite <=> exponent field != all ones (integer test on the high word: cheaper than isfinite() on FP64)
__device__ __forceinline__ bool finite_bits(double v) { return (__double2hiint(v) & 0x7ff00000) != 0x7ff00000; }

__device__ __forceinline__ void stat_update(double v, double& vmin, double& vmax, int& fin, int& nan) {
    if (finite_bits(v)) {
        vmin = v < vmin ? v : vmin;
        vmax = v > vmax ? v : vmax;
        ++fin;
    } else if (v != v) ++nan;
}

__global__ void __launch_bounds__(TILE_C * TILE_R) k_col_stats(const double* __restrict__ m, int64_t rows, int cols,
                                                              int64_t rows_per_slice, ColPartial* __restrict__ part) {
    __shared__ ColPartial sh[TILE_R][TILE_C];
    const int c = blockIdx.x * TILE_C + threadIdx.x;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_slice;
    const int64_t r_hi = r_lo + rows_per_slice < rows ? r_lo + rows_per_slice : rows;
    double vmin = INFINITY, vmax = -INFINITY;
    int fin = 0, nan = 0;                                   // a thread sees rows_per_slice / TILE_R rows: int is plenty
    if (c < cols) {
        const int64_t first = r_lo + threadIdx.y;
        const int64_t mine = first < r_hi ? (r_hi - first + TILE_R - 1) / TILE_R : 0;     // rows of this thread
        const size_t step = (size_t)TILE_R * cols;
        const double* ptr = m + (size_t)first * cols + c;
        int64_t k = 0;
        for (; k + UNR <= mine; k += UNR) {                // UNR independent loads in flight
            double v[UNR];
#pragma unroll
            for (int q = 0; q < UNR; ++q) v[q] = ptr[(size_t)q * step];
            ptr += (size_t)UNR * step;
#pragma unroll
            for (int q = 0; q < UNR; ++q) stat_update(v[q], vmin, vmax, fin, nan);
        }
        for (; k < mine; ++k) { stat_update(*ptr, vmin, vmax, fin, nan); ptr += step; }
    }
    ColPartial p;
    p.vmin = vmin; p.vmax = vmax; p.finite = fin; p.nan = nan;
    sh[threadIdx.y][threadIdx.x] = p;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int y = 1; y < TILE_R; ++y) {
            const ColPartial q = sh[y][threadIdx.x];
            p.vmin = fmin(p.vmin, q.vmin); p.vmax = fmax(p.vmax, q.vmax); p.finite += q.finite; p.nan += q.nan;
        }
        part[(size_t)blockIdx.y * cols + c] = p;
    }
}

// stats[0..cols) = min, [cols..2cols) = max, [2cols..3cols) = median (left untouched here), counts separately
__global__ void k_col_reduce(const ColPartial* __restrict__ part, int slices, int cols, double* __restrict__ stats,
                             long long* __restrict__ counts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    ColPartial p = part[c];
    for (int s = 1; s < slices; ++s) {
        const ColPartial q = part[(size_t)s * cols + c];
        p.vmin = fmin(p.vmin, q.vmin); p.vmax = fmax(p.vmax, q.vmax); p.finite += q.finite; p.nan += q.nan;
    }
    const bool none = p.finite == 0;          // no finite value at all: 0 replaces everything (:194-203)
    stats[c] = none ? 0.0 : p.vmin;
    stats[cols + c] = none ? 0.0 : p.vmax;
    stats[2 * cols + c] = none ? 0.0 : NAN;     // medians are filled in per column afterwards (NaN = not computed)
    counts[c] = p.finite;
    counts[cols + c] = p.nan;
}

// column c with every non-finite entry pushed to +inf so that the finite values sort to the front
__global__ void k_gather_column(const double* __restrict__ m, int64_t rows, int cols, int c, double* __restrict__ dst) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const double v = m[(size_t)r * cols + c];
        dst[r] = isfinite(v) ? v : INFINITY;
    }
}

// np.ma.median: middle element, or the mean of the two middle elements
__global__ void k_pick_median(const double* __restrict__ sorted, long long finite, double* __restrict__ dst) {
    const double lo = sorted[(finite - 1) / 2], hi = sorted[finite / 2];
    *dst = (lo == hi) ? lo : (lo + hi) / 2.0;
}

// zero_mode 0: +inf -> max, -inf -> min, NaN -> median;  zero_mode 1: every non-finite value -> 0.
// With the partials of the statistics sweep a block first asks whether its (row slice, column tile) holds any
// non-finite value at all and leaves otherwise: on a typical feature matrix the replacement sweep touches a few tiles.
__global__ void __launch_bounds__(TILE_C * TILE_R) k_impute_apply(double* __restrict__ m, int64_t rows, int cols,
                                                                 int64_t rows_per_slice, const double* __restrict__ stats,
                                                                 const ColPartial* __restrict__ part, int zero_mode) {
    const int c = blockIdx.x * TILE_C + threadIdx.x;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_slice;
    const int64_t r_hi = r_lo + rows_per_slice < rows ? r_lo + rows_per_slice : rows;
    if (part) {
        const int dirty = (c < cols) && (part[(size_t)blockIdx.y * cols + c].finite != (long long)(r_hi - r_lo));
        if (!__syncthreads_or(dirty)) return;
    }
    if (c >= cols) return;
    double vmin = 0.0, vmax = 0.0, vmed = 0.0;
    if (!zero_mode) { vmin = stats[c]; vmax = stats[cols + c]; vmed = stats[2 * cols + c]; }
    for (int64_t r = r_lo + threadIdx.y; r < r_hi; r += UNR * TILE_R) {
        double v[UNR];
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
            const int64_t rr = r + (int64_t)q * TILE_R;
            v[q] = rr < r_hi ? m[(size_t)rr * cols + c] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < UNR; ++q)
            if (!isfinite(v[q])) m[(size_t)(r + (int64_t)q * TILE_R) * cols + c] = (v[q] != v[q]) ? vmed : (v[q] > 0.0 ? vmax : vmin);
    }
}

cudaError_t reserve(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) *cap = bytes;
    return e;
}

}  // namespace

void ImputeWorkspace::release() {
    for (int i = 0; i < 5; ++i) { if (bufs[i]) cudaFree(bufs[i]); bufs[i] = nullptr; caps[i] = 0; }
}

#define ICK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return e__; } while (0)

cudaError_t impute_device(ImputeWorkspace& W, double* d_m, int64_t rows, int cols, int mode, bool all_medians,
                          double* h_stats, int sm_count, cudaStream_t st, int* launches) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    if (rows > 0x7fffffffLL) return cudaErrorInvalidValue;      // cub::DeviceRadixSort item count
    int n_launch = 0;
    const int ctiles = (cols + TILE_C - 1) / TILE_C;
    int slices = std::max(1, (sm_count * 8 + ctiles - 1) / ctiles);
    int64_t rows_per_slice = (rows + slices - 1) / slices;
    rows_per_slice = std::max<int64_t>(rows_per_slice, TILE_R);
    slices = (int)((rows + rows_per_slice - 1) / rows_per_slice);
    const dim3 grid(ctiles, slices), block(TILE_C, TILE_R);
    ICK(reserve(&W.bufs[1], &W.caps[1], (size_t)3 * cols * sizeof(double)));
    double* d_stats = (double*)W.bufs[1];

    if (mode == TSFX_IMPUTE_ZERO) {
        k_impute_apply<<<grid, block, 0, st>>>(d_m, rows, cols, rows_per_slice, d_stats, nullptr, 1);
        if (launches) *launches = 1;
        return cudaGetLastError();
    }
    if (mode == TSFX_IMPUTE_GIVEN) {           // caller-provided replacement values (impute_dataframe_range)
        ICK(cudaMemcpyAsync(d_stats, h_stats, (size_t)3 * cols * sizeof(double), cudaMemcpyHostToDevice, st));
        k_impute_apply<<<grid, block, 0, st>>>(d_m, rows, cols, rows_per_slice, d_stats, nullptr, 0);
        if (launches) *launches = 1;
        return cudaGetLastError();
    }

    ICK(reserve(&W.bufs[0], &W.caps[0], (size_t)slices * cols * sizeof(ColPartial)));
    ICK(reserve(&W.bufs[2], &W.caps[2], (size_t)2 * cols * sizeof(long long)));
    long long* d_counts = (long long*)W.bufs[2];
    k_col_stats<<<grid, block, 0, st>>>(d_m, rows, cols, rows_per_slice, (ColPartial*)W.bufs[0]);
    k_col_reduce<<<(cols + 127) / 128, 128, 0, st>>>((const ColPartial*)W.bufs[0], slices, cols, d_stats, d_counts);
    n_launch += 2;
    std::vector<long long> counts((size_t)2 * cols);
    ICK(cudaMemcpyAsync(counts.data(), d_counts, counts.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    ICK(cudaStreamSynchronize(st));

    // medians: only the columns that will use one (a NaN to replace) unless the caller wants them all
    size_t temp_bytes = 0;
    bool sized = false;
    for (int c = 0; c < cols; ++c) {
        const long long finite = counts[c], nan = counts[cols + c];
        if (finite == 0) continue;                                  // median already 0
        if (!all_medians && nan == 0) continue;
        if (!sized) {
            ICK(reserve(&W.bufs[3], &W.caps[3], (size_t)2 * rows * sizeof(double)));
            ICK(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const double*)W.bufs[3], (double*)W.bufs[3] + rows,
                                               (int)rows, 0, 64, st));
            ICK(reserve(&W.bufs[4], &W.caps[4], std::max<size_t>(temp_bytes, 16)));
            sized = true;
        }
        double* col = (double*)W.bufs[3];
        const int gblocks = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)sm_count * 8);
        k_gather_column<<<gblocks, 256, 0, st>>>(d_m, rows, cols, c, col);
        size_t tb = W.caps[4];
        ICK(cub::DeviceRadixSort::SortKeys(W.bufs[4], tb, (const double*)col, col + rows, (int)rows, 0, 64, st));
        k_pick_median<<<1, 1, 0, st>>>(col + rows, finite, d_stats + 2 * cols + c);
        n_launch += 4;                                              // gather + radix sort passes (counted as 2) + pick
    }
    if (mode == TSFX_IMPUTE_RANGE) {
        k_impute_apply<<<grid, block, 0, st>>>(d_m, rows, cols, rows_per_slice, d_stats, (const ColPartial*)W.bufs[0], 0);
        ++n_launch;
    }
    if (h_stats) {
        ICK(cudaMemcpyAsync(h_stats, d_stats, (size_t)3 * cols * sizeof(double), cudaMemcpyDeviceToHost, st));
        ICK(cudaStreamSynchronize(st));
    }
    if (launches) *launches = n_launch;
    return cudaGetLastError();
}

}  // namespace tsfx
