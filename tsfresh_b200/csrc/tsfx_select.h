// tsfx_select.h -- feature-selection statistics on the device (see tsfx_select.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/tsfx.h"

namespace tsfx {

struct SelectWorkspace {
    void* bufs[20] = {nullptr};
    size_t caps[20] = {0};
    cudaError_t reserve(int slot, size_t bytes);
    void release();
};

// d_X: [n x ncols] row-major float64 on the device; d_y: class code (0 .. n_classes-1) of every row; class_counts (host).
// d_out: [n_classes x ncols x TSFX_SEL_NSTAT] (layout: include/tsfx.h tsfx_select_classification).  *h_nan (pinned or
// pageable host int) receives 1 when the matrix holds a NaN; the caller synchronises the stream.
int select_class_stats(SelectWorkspace& W, const double* d_X, int64_t n, int ncols, const int32_t* d_y, int n_classes,
                       const int64_t* class_counts, double* d_out, int* h_nan, cudaStream_t st, std::string* msg);

// Regression targets (relevance.py:282-296; significance_tests.py:135-188): d_y is the float64 target of every row.
// d_out: [ncols x TSFX_SEL_NSTAT] + 4 trailing doubles (ytie, y0, y1, n) -- layout in include/tsfx.h
// (tsfx_select_regression).  Kendall's tau needs the discordant pairs: the column is brought into (x, rank(y)) order by
// two stable sorts and the strict inversions of the rank sequence are counted by a bottom-up merge (one kernel per level).
int select_regression_stats(SelectWorkspace& W, const double* d_X, int64_t n, int ncols, const double* d_y, double* d_out,
                            int* h_nan, cudaStream_t st, std::string* msg);

}  // namespace tsfx
