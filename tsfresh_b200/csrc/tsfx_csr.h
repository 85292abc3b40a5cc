// tsfx_csr.h -- stage (a): long (id, sort key, value) frame -> CSR [ids, begin, len, values] on the device.
// Restates what the reference's adapters do per series on the host (tsfresh/feature_extraction/data.py:
// groupby(id) :217/:280, per-group sort_values(sort) :226/:289, value column slice :230/:291) as one
// device pass: sortedness check -> (if needed) two stable radix sorts -> run-length encode -> scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace tsfx {

struct CsrWorkspace {
    // outputs (device)
    int64_t* d_uid = nullptr;     // unique ids, ascending
    int64_t* d_begin = nullptr;
    int32_t* d_len = nullptr;
    float* d_values = nullptr;    // values in (id, sort key) order
    // internals
    void* bufs[12] = {nullptr};
    size_t caps[12] = {0};
    void release();
    cudaError_t reserve(int slot, size_t bytes);
};

// Host inputs -> device CSR.  Returns 0 or a TSFX_E_* code with a message.
int csr_build_from_host(CsrWorkspace& W, const int64_t* ids, const void* sort_keys, int sort_key_is_f64,
                        const float* values, int64_t n_rows, cudaStream_t st, int64_t* n_series,
                        std::string* msg);

// max over a device int32 array (synchronises the stream)
int csr_max_len(CsrWorkspace& W, const int32_t* d_len, int64_t n, cudaStream_t st, int* out);

}  // namespace tsfx
