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
    for (int i = 0; i < 20; ++i) { if (bufs[i]) cudaFree(bufs[i]); bufs[i] = nullptr; caps[i] = 0; }
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

// ------------------------------------------------------------------------------------------ regression targets
// Kendall's tau-b with the asymptotic p-value (scipy.stats.kendalltau(x, y, method="asymptotic"), called at
// significance_tests.py:187) needs, per feature column: the discordant pairs dis = #{ x_a < x_b, y_a > y_b }, the tie
// sums of x ( sum t(t-1)/2, sum t(t-1)(t-2), sum t(t-1)(2t+5) ), the joint ties of (x, y) and the same tie sums of y.
// scipy's recipe, restated for the device: order the rows by y and replace y by its dense rank; stable-sort them by x
// (ties in x stay in rank order); dis = strict inversions of the rank sequence (inside a tie group of x the ranks ascend,
// so only pairs with different x count).  The inversions come from a bottom-up merge: at run width w every element of a
// right run adds the number of larger elements of its left sibling (binary search) and every element writes itself to
// its merged position (rank in the sibling + own index) -- one kernel per level, log2(n) levels.
struct RegAcc {
    unsigned long long runs, tie2, tie3, tie5;     // distinct x values; sum t(t-1)/2, sum t(t-1)(t-2), sum t(t-1)(2t+5)
    unsigned long long joint2;                     // sum c(c-1)/2 over runs of equal (x, rank(y))
    unsigned long long dis;                        // strict inversions
    double first_t, last_t, last_key;              // binary features: counts of the smaller / larger value, the larger value
    unsigned long long d_bits;                     // KS statistic (binary feature, real target), float64 bits
    double ones;
};

__global__ void k_reg_gather(const double* __restrict__ X, int64_t n, int ncols, int c, const uint32_t* __restrict__ order,
                             double* __restrict__ keys, int* __restrict__ nan_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i < n; i += stride) {
        double v = X[(size_t)order[i] * ncols + c];
        bad |= v != v;
        if (v == 0.0) v = 0.0;
        keys[i] = v;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(nan_flag, 1);
}

__global__ void k_iota_keys(const double* __restrict__ y, int64_t n, double* __restrict__ keys, uint32_t* __restrict__ rows, int* nan_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i < n; i += stride) {
        double v = y[i];
        bad |= v != v;
        if (v == 0.0) v = 0.0;
        keys[i] = v;
        rows[i] = (uint32_t)i;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(nan_flag, 1);
}

// head flag of every position of a sorted array (1 where a new value starts); its inclusive sum - 1 is the dense rank
__global__ void k_head_flags(const double* __restrict__ keys, int64_t n, int32_t* __restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) flags[i] = (i > 0 && keys[i] != keys[i - 1]) ? 1 : 0;
}

// tie sums over the runs of equal keys (JOINT: equal key AND equal rank); every thread owns the runs that start in its chunk
template <bool JOINT>
__global__ void k_reg_runs(const double* __restrict__ keys, const int32_t* __restrict__ ranks, int64_t n, RegAcc* acc) {
    const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * SEL_CHUNK;
    const int64_t end = base + SEL_CHUNK < n ? base + SEL_CHUNK : n;
    int64_t i = base;
    auto same = [&](int64_t a, int64_t b) { return keys[a] == keys[b] && (!JOINT || ranks[a] == ranks[b]); };
    if (i < n && i > 0) while (i < end && same(i, i - 1)) ++i;
    unsigned long long runs = 0, t2 = 0, t3 = 0, t5 = 0;
    while (i < end) {
        const int64_t a = i;
        do { ++i; } while (i < n && same(i, a));
        const unsigned long long t = (unsigned long long)(i - a);
        runs += 1;
        t2 += t * (t - 1) / 2;
        if (!JOINT) {
            t3 += t * (t - 1) * (t >= 2 ? t - 2 : 0);
            t5 += t * (t - 1) * (2 * t + 5);
            if (a == 0) acc->first_t = (double)t;
            if (i == n) { acc->last_t = (double)t; acc->last_key = keys[a]; }
        }
    }
    if (runs) {
        if (JOINT) atomicAdd(&acc->joint2, t2);
        else { atomicAdd(&acc->runs, runs); atomicAdd(&acc->tie2, t2); atomicAdd(&acc->tie3, t3); atomicAdd(&acc->tie5, t5); }
    }
}

// one level of the bottom-up merge: runs of width w in `in` are sorted; out gets runs of width 2w; *dis += strict inversions
__global__ void k_merge_count(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int64_t w, unsigned long long* dis) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (; i < n; i += stride) {
        const int64_t pair = i / (2 * w) * (2 * w);
        const int64_t mid = pair + w < n ? pair + w : n, hi = pair + 2 * w < n ? pair + 2 * w : n;
        const int32_t v = in[i];
        if (i < mid) {                                  // left run: position = own index + #(right < v)
            int64_t lo_ = mid, hi_ = hi;
            while (lo_ < hi_) { const int64_t m = (lo_ + hi_) >> 1; if (in[m] < v) lo_ = m + 1; else hi_ = m; }
            out[i + (lo_ - mid)] = v;
        } else {                                        // right run: position = own index - (mid - pair) + #(left <= v)
            int64_t lo_ = pair, hi_ = mid;
            while (lo_ < hi_) { const int64_t m = (lo_ + hi_) >> 1; if (in[m] <= v) lo_ = m + 1; else hi_ = m; }
            out[pair + (i - mid) + (lo_ - pair)] = v;
            mine += (unsigned long long)(mid - lo_);    // left elements strictly greater than v
        }
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(dis, mine);
}

__global__ void k_reg_labels(const double* __restrict__ keys_in_y_order, int64_t n, const RegAcc* acc, int32_t* __restrict__ lab) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double x1 = acc->last_key;
    for (; i < n; i += stride) lab[i] = keys_in_y_order[i] == x1 ? 1 : 0;
}

__global__ void k_reg_finish(const RegAcc* acc, const SelAcc* ks, int ncols, double n, const RegAcc* yacc, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        double* t = out + (size_t)ncols * TSFX_SEL_NSTAT;
        t[0] = (double)yacc->tie2; t[1] = (double)yacc->tie3; t[2] = (double)yacc->tie5; t[3] = n;
    }
    if (i >= ncols) return;
    const RegAcc a = acc[i];
    double* o = out + (size_t)i * TSFX_SEL_NSTAT;
    const double type = a.runs <= 1 ? 0.0 : (a.runs == 2 ? 1.0 : 2.0);
    o[0] = type; o[1] = n;
    if (type == 1.0) {            // binary feature: KS of the target between the rows with the larger / the smaller value
        o[2] = __longlong_as_double((long long)ks[i].d_bits); o[3] = a.last_t; o[4] = a.first_t; o[5] = 0.0; o[6] = 0.0; o[7] = 0.0;
    } else {
        o[2] = (double)a.dis; o[3] = (double)a.tie2; o[4] = (double)a.tie3; o[5] = (double)a.tie5; o[6] = (double)a.joint2; o[7] = 0.0;
    }
}

int select_regression_stats(SelectWorkspace& W, const double* d_X, int64_t n, int ncols, const double* d_y, double* d_out,
                            int* h_nan, cudaStream_t st, std::string* msg) {
    if (n >= ((int64_t)1 << 31)) { if (msg) *msg = "more than 2^31-1 rows"; return TSFX_E_UNSUPPORTED; }
    enum { B_KEYS_A = 0, B_KEYS_B, B_ROWS_A, B_ROWS_B, B_TEMP, B_BLOCKS, B_ACC, B_FLAG,
           R_YS = 8, R_PY, R_YR, R_IOTA, R_SEQ_A, R_SEQ_B, R_LAB, R_RACC, R_HOST };
    const int threads = 256;
    const int64_t per_block = (int64_t)threads * SEL_CHUNK;
    const int nblocks = (int)std::max<int64_t>(1, (n + per_block - 1) / per_block);
    const int gg = (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, 148 * 8));
    CKS(W.reserve(B_KEYS_A, n * 8), "alloc"); CKS(W.reserve(B_KEYS_B, n * 8), "alloc");
    CKS(W.reserve(B_ROWS_A, n * 4), "alloc"); CKS(W.reserve(B_ROWS_B, n * 4), "alloc");
    CKS(W.reserve(B_BLOCKS, (size_t)(nblocks + 1) * 4 * 2), "alloc");
    CKS(W.reserve(B_ACC, (size_t)ncols * sizeof(SelAcc)), "alloc");
    CKS(W.reserve(B_FLAG, 64), "alloc");
    CKS(W.reserve(R_YS, n * 8), "alloc"); CKS(W.reserve(R_PY, n * 4), "alloc"); CKS(W.reserve(R_YR, n * 4), "alloc");
    CKS(W.reserve(R_IOTA, n * 4), "alloc"); CKS(W.reserve(R_SEQ_A, n * 4), "alloc"); CKS(W.reserve(R_SEQ_B, n * 4), "alloc");
    CKS(W.reserve(R_LAB, n * 4), "alloc"); CKS(W.reserve(R_RACC, (size_t)(ncols + 1) * sizeof(RegAcc)), "alloc");
    double *ka = (double*)W.bufs[B_KEYS_A], *kb = (double*)W.bufs[B_KEYS_B], *ys = (double*)W.bufs[R_YS];
    uint32_t *ra = (uint32_t*)W.bufs[B_ROWS_A], *py = (uint32_t*)W.bufs[R_PY], *iota = (uint32_t*)W.bufs[R_IOTA];
    int32_t *yr = (int32_t*)W.bufs[R_YR], *sa = (int32_t*)W.bufs[R_SEQ_A], *sb = (int32_t*)W.bufs[R_SEQ_B], *lab = (int32_t*)W.bufs[R_LAB];
    unsigned int* blk = (unsigned int*)W.bufs[B_BLOCKS];
    unsigned int* blk_ex = blk + nblocks + 1;
    SelAcc* ksacc = (SelAcc*)W.bufs[B_ACC];
    RegAcc* racc = (RegAcc*)W.bufs[R_RACC];
    RegAcc* yacc = racc + ncols;
    int* d_flag = (int*)W.bufs[B_FLAG];
    size_t tb1 = 0, tb2 = 0, tb3 = 0, tb4 = 0;
    CKS(cub::DeviceRadixSort::SortPairs(nullptr, tb1, ka, kb, ra, py, (int)n, 0, 64, st), "sort size");
    CKS(cub::DeviceRadixSort::SortPairs(nullptr, tb2, ka, kb, yr, sa, (int)n, 0, 64, st), "sort size");
    CKS(cub::DeviceScan::InclusiveSum(nullptr, tb3, yr, yr, (int)n, st), "scan size");
    CKS(cub::DeviceScan::ExclusiveSum(nullptr, tb4, blk, blk_ex, nblocks, st), "scan size");
    CKS(W.reserve(B_TEMP, std::max(std::max(tb1, tb2), std::max(tb3, tb4))), "alloc temp");
    CKS(cudaMemsetAsync(racc, 0, (size_t)(ncols + 1) * sizeof(RegAcc), st), "memset");
    CKS(cudaMemsetAsync(ksacc, 0, (size_t)ncols * sizeof(SelAcc), st), "memset");
    CKS(cudaMemsetAsync(d_flag, 0, 64, st), "memset");
    // ---- the target, once: order by y, dense ranks, tie sums
    k_iota_keys<<<gg, threads, 0, st>>>(d_y, n, ka, ra, d_flag);
    CKS(cub::DeviceRadixSort::SortPairs(W.bufs[B_TEMP], tb1, ka, ys, ra, py, (int)n, 0, 64, st), "sort y");
    k_head_flags<<<gg, threads, 0, st>>>(ys, n, yr);
    CKS(cub::DeviceScan::InclusiveSum(W.bufs[B_TEMP], tb3, yr, yr, (int)n, st), "rank scan");
    k_reg_runs<false><<<nblocks, threads, 0, st>>>(ys, yr, n, yacc);
    CKS(cudaMemcpyAsync(iota, ra, n * 4, cudaMemcpyDeviceToDevice, st), "iota");      // ra still holds 0..n-1
    RegAcc host_acc;
    for (int c = 0; c < ncols; ++c) {
        RegAcc* acc = racc + c;
        k_reg_gather<<<gg, threads, 0, st>>>(d_X, n, ncols, c, py, ka, d_flag);                   // x in y order
        CKS(cub::DeviceRadixSort::SortPairs(W.bufs[B_TEMP], tb2, ka, kb, yr, sa, (int)n, 0, 64, st), "sort x");   // stable: ties keep rank order
        k_reg_runs<false><<<nblocks, threads, 0, st>>>(kb, sa, n, acc);
        CKS(cudaMemcpyAsync(&host_acc, acc, sizeof(RegAcc), cudaMemcpyDeviceToHost, st), "D2H");
        CKS(cudaStreamSynchronize(st), "sync");
        if (host_acc.runs <= 1) continue;                                                          // constant feature
        if (host_acc.runs == 2) {
            // binary feature: two-sample KS of the target between the two groups (significance_tests.py:135-167)
            k_reg_labels<<<gg, threads, 0, st>>>(ka, n, acc, lab);
            k_sel_count<<<nblocks, threads, 0, st>>>(iota, lab, 1, n, blk);
            CKS(cub::DeviceScan::ExclusiveSum(W.bufs[B_TEMP], tb4, blk, blk_ex, nblocks, st), "scan");
            k_sel_runs<<<nblocks, threads, 0, st>>>(ys, iota, lab, 1, n, blk_ex, host_acc.last_t, host_acc.first_t, ksacc + c);
            continue;
        }
        k_reg_runs<true><<<nblocks, threads, 0, st>>>(kb, sa, n, acc);
        int32_t *src = sa, *dst = sb;
        for (int64_t w = 1; w < n; w <<= 1) {
            k_merge_count<<<gg, threads, 0, st>>>(src, dst, n, w, &acc->dis);
            std::swap(src, dst);
        }
    }
    k_reg_finish<<<(ncols + 127) / 128, 128, 0, st>>>(racc, ksacc, ncols, (double)n, yacc, d_out);
    CKS(cudaGetLastError(), "launch");
    CKS(cudaMemcpyAsync(h_nan, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H");
    return TSFX_OK;
}

}  // namespace tsfx
