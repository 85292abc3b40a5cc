// k_assemble.cu -- scatters the per-group staging matrices into the final [n_series x ncols] matrix
// (the dense float64 result PartitionedTsData.pivot would build, tsfresh/feature_extraction/data.py:86-121).
// One warp per row: coalesced reads of the staging rows into shared memory at their final column,
// then one coalesced write of the whole row.  Pure HBM traffic: 16 bytes per feature value.
// Multi-GPU: the same row is also stored into the peers' result matrices (NVLink P2P stores or one multicast store),
// which replaces the all-gather of SURVEY.md section 8e -- no collective kernel, no extra pass over the matrix.
#include <algorithm>

#include "tsfx_common.cuh"
#include "tsfx_kernels.h"

namespace tsfx {

template <int WPC>
__global__ void __launch_bounds__(WPC * 32) k_assemble(AssembleArgs A, int row_pad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* row = reinterpret_cast<double*>(smem_raw) + (size_t)warp * row_pad;
    const int64_t warps_total = (int64_t)gridDim.x * WPC;
    for (int64_t s = (int64_t)blockIdx.x * WPC + warp; s < A.n_series; s += warps_total) {
        for (int j = lane; j < A.ncols; j += 32) row[j] = dnan();      // columns no descriptor covers
        __syncwarp();
        for (int g = 0; g < A.n_groups; ++g) {
            const int c0 = A.cum[g], w = A.cum[g + 1] - c0;
            if (w == 0) continue;
            const double* src = A.stage + (size_t)A.n_series * c0 + (size_t)s * w;
            for (int j = lane; j < w; j += 32) row[__ldg(A.final_col + c0 + j)] = __ldcs(src + j);
        }
        __syncwarp();
        const size_t roff = (size_t)s * A.ld;
        if (A.out_mc) {
            double* dst = A.out_mc + roff;
            for (int j = lane; j < A.ncols; j += 32)
                asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" :: "l"(dst + j), "d"(row[j]) : "memory");
        } else {
            double* dst = A.out + roff;
            for (int j = lane; j < A.ncols; j += 32) __stcs(dst + j, row[j]);
        }
        for (int p = 0; p < A.n_extra; ++p) {
            double* dst = A.extra[p] + roff;
            for (int j = lane; j < A.ncols; j += 32) __stcs(dst + j, row[j]);
        }
        __syncwarp();
    }
}

template <int WPC>
static cudaError_t launch_assemble_w(const AssembleArgs& A, cudaStream_t st, int sm_count, int row_pad) {
    size_t smem = (size_t)WPC * row_pad * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(k_assemble<WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t ctas = (A.n_series + WPC - 1) / WPC;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)sm_count * 8));
    k_assemble<WPC><<<grid, WPC * 32, smem, st>>>(A, row_pad);
    return cudaGetLastError();
}

cudaError_t launch_assemble(const AssembleArgs& A, cudaStream_t st, int sm_count) {
    const int row_pad = (A.ncols + 1) & ~1;
    const size_t row_bytes = (size_t)row_pad * sizeof(double);
    if (row_bytes * 8 <= 200 * 1024) return launch_assemble_w<8>(A, st, sm_count, row_pad);
    if (row_bytes * 2 <= 200 * 1024) return launch_assemble_w<2>(A, st, sm_count, row_pad);
    if (row_bytes <= 227 * 1024) return launch_assemble_w<1>(A, st, sm_count, row_pad);
    return cudaErrorInvalidConfiguration;          // more than ~29 000 columns in one plan
}

}  // namespace tsfx
