// tsfx_select.cu -- feature-selection statistics on the resident feature matrix.
// Restates, column-parallel on the device, what tsfresh/feature_selection/relevance.py:31-322 does per feature with a
// Python loop (get_feature_type :325-345, target_binary_feature_real_test / _binary_test, significance_tests.py:43-132):
// for a classification target every feature column needs its type (constant / binary / real) and, per class label
// (one-vs-rest), the Mann-Whitney U statistic with its tie correction, the two-sample Kolmogorov-Smirnov statistic and
// -- for binary features -- the 2 x 2 contingency table.  All of them are functions of the SORTED column:
//   gather column c of the row-major matrix -> radix sort (value, row) -> one pass over the runs of equal values.
// For a run [a, b) with c1 rows of the class and z rows of the rest BEFORE it:
//   U1 += c1 * (z + (b - a - c1) / 2)      tie += (b - a)^3 - (b - a)      D = max |ones(b) / n1 - zeros(b) / n0|
// (U terms are multiples of 1/2 and stay exact in float64 in any summation order).  The p-values are finished on the host
// from these sufficient statistics with scipy's own distribution functions (tsfresh_b200/feature_selection.py).
#include <cub/cub.cuh>
#include <algorithm>
#include <string>

#include "../../include/tsfx.h"
#include "tsfx_select.h"

namespace tsfx {

cudaError_t SelectWorkspace::reserve(int slot, size_t bytes) {
    if (bytes <= caps[slot]) return cudaSuccess;
    if (bufs[slot]) cudaFree(bufs[slot]);
    bufs[slot] = nullptr;
    caps[slot] = 0;
    cudaError_t e = cudaMalloc(&bufs[slot], bytes);
    if (e == cudaSuccess) caps[slot] = bytes;
    return e;
}
void SelectWorkspace::release() {
    for (int i = 0; i < 8; ++i) { if (bufs[i]) cudaFree(bufs[i]); bufs[i] = nullptr; caps[i] = 0; }
}

// column c of the row-major matrix -> contiguous keys, identity row index; NaN -> flag
__global__ void k_sel_gather(const double* __restrict__ X, int64_t n, int ncols, int c, double* __restrict__ keys,
                             uint32_t* __restrict__ rows, int* __restrict__ nan_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i < n; i += stride) {
        double v = X[(size_t)i * ncols + c];
        bad |= v != v;
        if (v == 0.0) v = 0.0;                 // -0.0 and +0.0 are one value (np.unique / comparisons), one radix key
        keys[i] = v;
        rows[i] = (uint32_t)i;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(nan_flag, 1);
}

struct SelAcc {                 // per (column, class) accumulators, zeroed before the pass
    double u1;                  // Mann-Whitney U of the class sample
    unsigned long long tie;     // sum t^3 - t
    unsigned long long runs;    // distinct values
    unsigned long long d_bits;  // KS statistic as float64 bits (non-negative doubles order like their bit patterns)
    double first_c1, first_t, last_c1, last_t;   // class count / length of the first and of the last run (binary features)
};

#define SEL_CHUNK 16            // sorted elements per thread

// ones[b] = rows of the class in block b's slice of the sorted column
__global__ void k_sel_count(const uint32_t* __restrict__ rows, const int32_t* __restrict__ y, int cls, int64_t n,
                            unsigned int* __restrict__ block_ones) {
    __shared__ unsigned int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * SEL_CHUNK;
    unsigned int c = 0;
    for (int k = 0; k < SEL_CHUNK; ++k) { const int64_t i = base + k; if (i < n) c += y[rows[i]] == cls; }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) block_ones[blockIdx.x] = s;
}

// every thread owns the runs that START inside its chunk of SEL_CHUNK sorted elements
__global__ void k_sel_runs(const double* __restrict__ keys, const uint32_t* __restrict__ rows, const int32_t* __restrict__ y, int cls,
                           int64_t n, const unsigned int* __restrict__ block_ones_excl, double n1, double n0, SelAcc* acc) {
    typedef cub::BlockScan<unsigned int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * SEL_CHUNK;
    unsigned int mine = 0;
    for (int k = 0; k < SEL_CHUNK; ++k) { const int64_t i = base + k; if (i < n) mine += y[rows[i]] == cls; }
    unsigned int before;
    Scan(tmp).ExclusiveSum(mine, before);
    unsigned long long ones = (unsigned long long)block_ones_excl[blockIdx.x] + before;     // class rows in [0, base)
    double u1 = 0.0, dmax = 0.0;
    unsigned long long tie = 0, runs = 0;
    int64_t i = base;
    const int64_t end = base + SEL_CHUNK < n ? base + SEL_CHUNK : n;
    if (i < n && i > 0) {                         // skip the tail of a run that started in an earlier chunk
        const double prev = keys[i - 1];
        while (i < end && keys[i] == prev) { ones += y[rows[i]] == cls; ++i; }
    }
    while (i < end) {                             // a run starts at i (it may extend past `end`: this thread finishes it)
        const double v = keys[i];
        const int64_t a = i;
        const unsigned long long ones_a = ones;
        do { ones += y[rows[i]] == cls; ++i; } while (i < n && keys[i] == v);      // always advances (a NaN key equals nothing)
        const double t = (double)(i - a), c1 = (double)(ones - ones_a);
        const double z_before = (double)a - (double)ones_a;
        u1 += c1 * (z_before + 0.5 * (t - c1));
        const unsigned long long tt = (unsigned long long)(i - a);
        tie += tt * tt * tt - tt;
        runs += 1;
        const double cdf1 = n1 > 0.0 ? (double)ones / n1 : 0.0, cdf0 = n0 > 0.0 ? ((double)i - (double)ones) / n0 : 0.0;
        dmax = fmax(dmax, fabs(cdf1 - cdf0));
        if (a == 0) { acc->first_c1 = c1; acc->first_t = t; }
        if (i == n) { acc->last_c1 = c1; acc->last_t = t; }
    }
    if (runs) {
        atomicAdd(&acc->u1, u1);
        atomicAdd(&acc->tie, tie);
        atomicAdd(&acc->runs, runs);
        atomicMax(&acc->d_bits, (unsigned long long)__double_as_longlong(dmax));
    }
}

__global__ void k_sel_finish(const SelAcc* acc, int n_items, double n1, double n0, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const SelAcc a = acc[i];
    double* o = out + (size_t)i * TSFX_SEL_NSTAT;
    const double type = a.runs <= 1 ? 0.0 : (a.runs == 2 ? 1.0 : 2.0);
    o[0] = type; o[1] = n1; o[2] = n0;
    if (type == 1.0) {       // binary feature: first run = smaller value x0, last run = larger value x1
        o[3] = a.last_c1;                      // y = class, x = x1
        o[4] = a.first_c1;                     // y = class, x = x0
        o[5] = a.last_t - a.last_c1;           // y = rest,  x = x1
        o[6] = a.first_t - a.first_c1;         // y = rest,  x = x0
        o[7] = a.u1;
    } else {
        o[3] = a.u1; o[4] = (double)a.tie; o[5] = __longlong_as_double((long long)a.d_bits); o[6] = (double)a.runs; o[7] = 0.0;
    }
}

#define CKS(call, what)                                                                    \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            if (msg) *msg = std::string(what) + ": " + cudaGetErrorString(e__);           \
            return TSFX_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)

int select_class_stats(SelectWorkspace& W, const double* d_X, int64_t n, int ncols, const int32_t* d_y, int n_classes,
                       const int64_t* class_counts, double* d_out, int* h_nan, cudaStream_t st, std::string* msg) {
    if (n >= ((int64_t)1 << 31)) { if (msg) *msg = "more than 2^31-1 rows"; return TSFX_E_UNSUPPORTED; }
    enum { B_KEYS_A = 0, B_KEYS_B, B_ROWS_A, B_ROWS_B, B_TEMP, B_BLOCKS, B_ACC, B_FLAG };
    const int threads = 256;
    const int64_t per_block = (int64_t)threads * SEL_CHUNK;
    const int nblocks = (int)std::max<int64_t>(1, (n + per_block - 1) / per_block);
    CKS(W.reserve(B_KEYS_A, n * 8), "alloc"); CKS(W.reserve(B_KEYS_B, n * 8), "alloc");
    CKS(W.reserve(B_ROWS_A, n * 4), "alloc"); CKS(W.reserve(B_ROWS_B, n * 4), "alloc");
    CKS(W.reserve(B_BLOCKS, (size_t)(nblocks + 1) * 4 * 2), "alloc");
    CKS(W.reserve(B_ACC, (size_t)ncols * n_classes * sizeof(SelAcc)), "alloc");
    CKS(W.reserve(B_FLAG, 64), "alloc");
    double *ka = (double*)W.bufs[B_KEYS_A], *kb = (double*)W.bufs[B_KEYS_B];
    uint32_t *ra = (uint32_t*)W.bufs[B_ROWS_A], *rb = (uint32_t*)W.bufs[B_ROWS_B];
    unsigned int* blk = (unsigned int*)W.bufs[B_BLOCKS];
    unsigned int* blk_ex = blk + nblocks + 1;
    SelAcc* acc = (SelAcc*)W.bufs[B_ACC];
    int* d_flag = (int*)W.bufs[B_FLAG];
    size_t tb_sort = 0, tb_scan = 0;
    CKS(cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, ka, kb, ra, rb, (int)n, 0, 64, st), "sort size");
    CKS(cub::DeviceScan::ExclusiveSum(nullptr, tb_scan, blk, blk_ex, nblocks, st), "scan size");
    CKS(W.reserve(B_TEMP, std::max(tb_sort, tb_scan)), "alloc temp");
    CKS(cudaMemsetAsync(acc, 0, (size_t)ncols * n_classes * sizeof(SelAcc), st), "memset");
    CKS(cudaMemsetAsync(d_flag, 0, 64, st), "memset");
    const int gg = (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, 148 * 8));
    for (int c = 0; c < ncols; ++c) {
        k_sel_gather<<<gg, threads, 0, st>>>(d_X, n, ncols, c, ka, ra, d_flag);
        CKS(cub::DeviceRadixSort::SortPairs(W.bufs[B_TEMP], tb_sort, ka, kb, ra, rb, (int)n, 0, 64, st), "sort");
        for (int k = 0; k < n_classes; ++k) {
            const double n1 = (double)class_counts[k], n0 = (double)(n - class_counts[k]);
            k_sel_count<<<nblocks, threads, 0, st>>>(rb, d_y, k, n, blk);
            CKS(cub::DeviceScan::ExclusiveSum(W.bufs[B_TEMP], tb_scan, blk, blk_ex, nblocks, st), "scan");
            k_sel_runs<<<nblocks, threads, 0, st>>>(kb, rb, d_y, k, n, blk_ex, n1, n0, acc + (size_t)k * ncols + c);
        }
    }
    for (int k = 0; k < n_classes; ++k) {
        const double n1 = (double)class_counts[k], n0 = (double)(n - class_counts[k]);
        k_sel_finish<<<(ncols + 127) / 128, 128, 0, st>>>(acc + (size_t)k * ncols, ncols, n1, n0, d_out + (size_t)k * ncols * TSFX_SEL_NSTAT);
    }
    CKS(cudaGetLastError(), "launch");
    CKS(cudaMemcpyAsync(h_nan, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H");
    return TSFX_OK;
}

}  // namespace tsfx
