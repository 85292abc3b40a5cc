// tsfx_csr.h -- stage (a): long (id, sort key, value) frame -> CSR [ids, begin, len, values] on the device.
// Restates what the reference's adapters do per series on the host (tsfresh/feature_extraction/data.py:
// groupby(id) :217/:280, per-group sort_values(sort) :226/:289, value column slice :230/:291) as device passes.
//
// Fast path (rows already ordered by (id, sort key), the layout the reference's own benchmarks produce): ONE pass
// over the id column (sortedness check + unique-by-key with the row index as payload -> unique ids and begin
// offsets, lengths, longest series, row-block boundaries) and ONE host synchronisation to learn the sizes; the
// sort-key / value columns are then streamed in row blocks while the kernels of earlier blocks run.
// Slow path (any other order): two stable 64-bit radix sorts of the row index (by sort key, then by id), gather.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace tsfx {

#define TSFX_CSR_MAX_BLOCKS 32

struct CsrInfo {                  // written by the device, mirrored in pinned host memory
    int32_t unsorted_ids;         // some id is smaller than its predecessor
    int32_t unsorted_keys;        // inside one id, some sort key is smaller than its predecessor
    int32_t has_nan;              // a value is NaN (data.py:148-167 raises ValueError)
    int32_t max_len;
    int64_t n_series;
    int32_t n_blocks;
    int32_t pad;
    int64_t series_lo[TSFX_CSR_MAX_BLOCKS + 1];   // block b covers series [series_lo[b], series_lo[b+1])
    int64_t row_lo[TSFX_CSR_MAX_BLOCKS + 1];      //             and rows [row_lo[b], row_lo[b+1])
};

struct CsrWorkspace {
    // outputs (device)
    int64_t* d_uid = nullptr;     // unique ids, ascending
    int64_t* d_begin = nullptr;
    int32_t* d_len = nullptr;
    float* d_values = nullptr;    // values in (id, sort key) order
    const uint32_t* d_perm = nullptr;   // after csr_sort_pass: input row of every CSR row (nullptr: rows were in order)
    // internals
    void* bufs[16] = {nullptr};
    size_t caps[16] = {0};
    CsrInfo* h_info = nullptr;    // pinned
    CsrInfo* d_info = nullptr;
    void release();
    cudaError_t reserve(int slot, size_t bytes);
    cudaError_t init_info();
    // device copies of the input columns (device-pointer callers alias their own buffers instead)
    int64_t* ids() const { return (int64_t*)bufs[0]; }
    uint64_t* keys() const { return (uint64_t*)bufs[1]; }
    float* vals() const { return (float*)bufs[2]; }
};

// Pass over the id column (already on the device): fills d_uid / d_begin / d_len for the case the ids are
// non-decreasing, and d_info (copied to h_info asynchronously; the caller synchronises).  min_block = smallest number
// of series per row block, max_blocks <= TSFX_CSR_MAX_BLOCKS.
int csr_ids_pass(CsrWorkspace& W, const int64_t* d_ids, int64_t n_rows, int64_t min_block, int max_blocks,
                 cudaStream_t st, std::string* msg);

// Rows [row_lo, row_hi): flags sort keys that decrease inside one id (pairs (i-1, i)) and NaN values.
void csr_check_rows(CsrWorkspace& W, const int64_t* d_ids, const uint64_t* d_keys, int is_f64, const float* d_values,
                    int64_t row_lo, int64_t row_hi, bool check_nan, cudaStream_t st);

// Slow path: ids / keys / values are on the device in arbitrary row order; sorts and rebuilds the CSR (d_values then
// points at the gathered copy), refreshes d_info / h_info (caller synchronises).
int csr_sort_pass(CsrWorkspace& W, const int64_t* d_ids, const uint64_t* d_keys, int is_f64, const float* d_values,
                  int64_t n_rows, int64_t min_block, int max_blocks, bool check_nan, cudaStream_t st, std::string* msg);

// dst[i] = src[d_perm[i]] (row timestamps follow the sort)
void csr_gather_i64(CsrWorkspace& W, const int64_t* src, int64_t* dst, int64_t n, cudaStream_t st);

void csr_gather_f32(CsrWorkspace& W, const float* src, float* dst, int64_t n, cudaStream_t st);

// longest series of a device CSR (synchronises the stream)
int csr_max_len(CsrWorkspace& W, const int32_t* d_len, int64_t n, cudaStream_t st, int* out);

// NaN scan of a device float array; result lands in d_info->has_nan (async)
void csr_check_nan(CsrWorkspace& W, const float* d_values, int64_t n, cudaStream_t st);

}  // namespace tsfx
