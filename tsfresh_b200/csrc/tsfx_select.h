// tsfx_select.h -- feature-selection statistics on the device (see tsfx_select.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/tsfx.h"

namespace tsfx {

struct SelectWorkspace {
    void* bufs[8] = {nullptr};
    size_t caps[8] = {0};
    cudaError_t reserve(int slot, size_t bytes);
    void release();
};

// d_X: [n x ncols] row-major float64 on the device; d_y: class code (0 .. n_classes-1) of every row; class_counts (host).
// d_out: [n_classes x ncols x TSFX_SEL_NSTAT] (layout: include/tsfx.h tsfx_select_classification).  *h_nan (pinned or
// pageable host int) receives 1 when the matrix holds a NaN; the caller synchronises the stream.
int select_class_stats(SelectWorkspace& W, const double* d_X, int64_t n, int ncols, const int32_t* d_y, int n_classes,
                       const int64_t* class_counts, double* d_out, int* h_nan, cudaStream_t st, std::string* msg);

}  // namespace tsfx
