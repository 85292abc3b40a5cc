// tsfx_csr.cu -- device CSR build (stage (a) of the hot path); see tsfx_csr.h.
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <algorithm>

#include "../../include/tsfx.h"
#include "tsfx_csr.h"

namespace tsfx {

enum { S_IDS = 0, S_KEYS, S_VALS, S_PERM_A, S_PERM_B, S_KEY_A, S_KEY_B, S_TEMP, S_FLAG, S_OUT_UID, S_OUT_BEGIN, S_OUT_LEN };
// S_VALS doubles as gather destination via a 13th implicit buffer: we keep sorted values in S_KEY_B's
// storage when a sort happened (see below), otherwise S_VALS itself is the CSR value array.

cudaError_t CsrWorkspace::reserve(int slot, size_t bytes) {
    if (bytes <= caps[slot]) return cudaSuccess;
    if (bufs[slot]) cudaFree(bufs[slot]);
    bufs[slot] = nullptr;
    caps[slot] = 0;
    cudaError_t e = cudaMalloc(&bufs[slot], bytes);
    if (e == cudaSuccess) caps[slot] = bytes;
    return e;
}
void CsrWorkspace::release() {
    for (int i = 0; i < 12; ++i) { if (bufs[i]) cudaFree(bufs[i]); bufs[i] = nullptr; caps[i] = 0; }
    d_uid = nullptr; d_begin = nullptr; d_len = nullptr; d_values = nullptr;
}

__device__ __forceinline__ uint64_t key_i64(int64_t v) { return (uint64_t)v ^ 0x8000000000000000ull; }
__device__ __forceinline__ uint64_t key_f64(uint64_t b) { return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull); }

__global__ void k_check_sorted(const int64_t* ids, const uint64_t* keys, int is_f64, int64_t n, int* flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i + 1 < n; i += stride) {
        int64_t a = ids[i], b = ids[i + 1];
        if (a > b) bad = 1;
        else if (a == b && keys) {
            uint64_t ka = keys[i], kb = keys[i + 1];
            if (is_f64) { ka = key_f64(ka); kb = key_f64(kb); } else { ka = key_i64((int64_t)ka); kb = key_i64((int64_t)kb); }
            if (ka > kb) bad = 1;
        }
    }
    if (bad) atomicOr(flag, 1);
}

__global__ void k_make_keys(const uint64_t* raw, int is_f64, int64_t n, uint64_t* key, uint32_t* perm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        if (key) key[i] = is_f64 ? key_f64(raw[i]) : key_i64((int64_t)raw[i]);
        perm[i] = (uint32_t)i;
    }
}
__global__ void k_gather_id_keys(const int64_t* ids, const uint32_t* perm, int64_t n, uint64_t* key) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) key[i] = key_i64(ids[perm[i]]);
}
__global__ void k_gather_final(const uint64_t* idkey_sorted, const float* values, const uint32_t* perm, int64_t n,
                               int64_t* ids_out, float* vals_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        ids_out[i] = (int64_t)(idkey_sorted[i] ^ 0x8000000000000000ull);
        vals_out[i] = values[perm[i]];
    }
}

struct ToI64 {
    __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};

#define CKE(call, what)                                                                    \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            if (msg) *msg = std::string(what) + ": " + cudaGetErrorString(e__);           \
            return TSFX_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)

int csr_build_from_host(CsrWorkspace& W, const int64_t* ids, const void* sort_keys, int sort_key_is_f64,
                        const float* values, int64_t n, cudaStream_t st, int64_t* n_series, std::string* msg) {
    if (n >= (int64_t)1 << 31) { if (msg) *msg = "more than 2^31-1 rows in one call"; return TSFX_E_UNSUPPORTED; }
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 148 * 16);
    CKE(W.reserve(S_IDS, n * 8), "alloc ids");
    CKE(W.reserve(S_VALS, n * 4), "alloc values");
    CKE(W.reserve(S_FLAG, 64), "alloc flag");
    CKE(cudaMemcpyAsync(W.bufs[S_IDS], ids, n * 8, cudaMemcpyHostToDevice, st), "H2D ids");
    CKE(cudaMemcpyAsync(W.bufs[S_VALS], values, n * 4, cudaMemcpyHostToDevice, st), "H2D values");
    if (sort_keys) {
        CKE(W.reserve(S_KEYS, n * 8), "alloc keys");
        CKE(cudaMemcpyAsync(W.bufs[S_KEYS], sort_keys, n * 8, cudaMemcpyHostToDevice, st), "H2D sort keys");
    }
    int* d_flag = (int*)W.bufs[S_FLAG];
    CKE(cudaMemsetAsync(d_flag, 0, 64, st), "memset");
    k_check_sorted<<<blocks, threads, 0, st>>>((const int64_t*)W.bufs[S_IDS], sort_keys ? (const uint64_t*)W.bufs[S_KEYS] : nullptr,
                                               sort_key_is_f64, n, d_flag);
    int unsorted = 0;
    CKE(cudaMemcpyAsync(&unsorted, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H flag");
    CKE(cudaStreamSynchronize(st), "sync");

    const int64_t* d_ids_sorted = (const int64_t*)W.bufs[S_IDS];
    W.d_values = (float*)W.bufs[S_VALS];
    if (unsorted) {
        CKE(W.reserve(S_PERM_A, n * 4), "alloc perm");
        CKE(W.reserve(S_PERM_B, n * 4), "alloc perm");
        CKE(W.reserve(S_KEY_A, n * 8), "alloc key");
        CKE(W.reserve(S_KEY_B, n * 8), "alloc key");
        uint32_t *pa = (uint32_t*)W.bufs[S_PERM_A], *pb = (uint32_t*)W.bufs[S_PERM_B];
        uint64_t *ka = (uint64_t*)W.bufs[S_KEY_A], *kb = (uint64_t*)W.bufs[S_KEY_B];
        size_t tb = 0;
        CKE(cub::DeviceRadixSort::SortPairs(nullptr, tb, ka, kb, pa, pb, (int)n, 0, 64, st), "radix size");
        CKE(W.reserve(S_TEMP, tb), "alloc temp");
        const uint32_t* perm = pa;
        k_make_keys<<<blocks, threads, 0, st>>>(sort_keys ? (const uint64_t*)W.bufs[S_KEYS] : nullptr, sort_key_is_f64, n,
                                                sort_keys ? ka : nullptr, pa);
        if (sort_keys) {      // pass A: stable sort of row indices by sort key
            CKE(cub::DeviceRadixSort::SortPairs(W.bufs[S_TEMP], tb, ka, kb, pa, pb, (int)n, 0, 64, st), "radix sort A");
            perm = pb;
        }
        // pass B: stable sort by id, carrying the pass-A order
        k_gather_id_keys<<<blocks, threads, 0, st>>>((const int64_t*)W.bufs[S_IDS], perm, n, ka);
        uint32_t* perm_out = (perm == pa) ? pb : pa;
        CKE(cub::DeviceRadixSort::SortPairs(W.bufs[S_TEMP], tb, ka, kb, perm, perm_out, (int)n, 0, 64, st), "radix sort B");
        // gather: sorted ids into S_KEYS storage (reuse), sorted values into S_KEY_A storage (reuse)
        CKE(W.reserve(S_KEYS, n * 8), "alloc ids sorted");
        k_gather_final<<<blocks, threads, 0, st>>>(kb, (const float*)W.bufs[S_VALS], perm_out, n,
                                                   (int64_t*)W.bufs[S_KEYS], (float*)ka);
        d_ids_sorted = (const int64_t*)W.bufs[S_KEYS];
        W.d_values = (float*)ka;
    }
    // run-length encode ids -> unique ids + counts ; n_series <= n
    CKE(W.reserve(S_OUT_UID, n * 8), "alloc uid");
    CKE(W.reserve(S_OUT_LEN, n * 4 + 64), "alloc len");
    W.d_uid = (int64_t*)W.bufs[S_OUT_UID];
    W.d_len = (int32_t*)W.bufs[S_OUT_LEN];
    int* d_nruns = d_flag + 4;
    size_t tb = 0;
    CKE(cub::DeviceRunLengthEncode::Encode(nullptr, tb, d_ids_sorted, W.d_uid, W.d_len, d_nruns, (int)n, st), "rle size");
    CKE(W.reserve(S_TEMP, tb), "alloc temp");
    CKE(cub::DeviceRunLengthEncode::Encode(W.bufs[S_TEMP], tb, d_ids_sorted, W.d_uid, W.d_len, d_nruns, (int)n, st), "rle");
    int nruns = 0;
    CKE(cudaMemcpyAsync(&nruns, d_nruns, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H nruns");
    CKE(cudaStreamSynchronize(st), "sync");
    CKE(W.reserve(S_OUT_BEGIN, (size_t)nruns * 8 + 8), "alloc begin");
    W.d_begin = (int64_t*)W.bufs[S_OUT_BEGIN];
    auto it = thrust::make_transform_iterator((const int32_t*)W.d_len, ToI64());
    tb = 0;
    CKE(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, W.d_begin, nruns, st), "scan size");
    CKE(W.reserve(S_TEMP, tb), "alloc temp");
    CKE(cub::DeviceScan::ExclusiveSum(W.bufs[S_TEMP], tb, it, W.d_begin, nruns, st), "scan");
    *n_series = nruns;
    return TSFX_OK;
}

int csr_max_len(CsrWorkspace& W, const int32_t* d_len, int64_t n, cudaStream_t st, int* out) {
    std::string* msg = nullptr;
    CKE(W.reserve(S_FLAG, 64), "alloc");
    int* d_out = (int*)W.bufs[S_FLAG] + 8;
    size_t tb = 0;
    CKE(cub::DeviceReduce::Max(nullptr, tb, d_len, d_out, (int)n, st), "max size");
    // S_TEMP may be in use by nobody at this point (stream ordered)
    CKE(W.reserve(S_TEMP, tb), "alloc temp");
    CKE(cub::DeviceReduce::Max(W.bufs[S_TEMP], tb, d_len, d_out, (int)n, st), "max");
    CKE(cudaMemcpyAsync(out, d_out, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H");
    CKE(cudaStreamSynchronize(st), "sync");
    return TSFX_OK;
}

}  // namespace tsfx
