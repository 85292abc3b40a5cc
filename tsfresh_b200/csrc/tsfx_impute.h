// tsfx_impute.h -- device-side impute of the feature matrix (tsfx_impute.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tsfx.h"

namespace tsfx {

struct ImputeWorkspace {
    void* bufs[5] = {nullptr};     // column partials, column statistics, counts, column sort buffers, cub temp
    size_t caps[5] = {0};
    void release();
};

// In-place imputation of the row-major device matrix d_m[rows x cols] (mode = TSFX_IMPUTE_*).
// h_stats (host, 3*cols doubles: min | max | median) is an output for RANGE / STATS and the input for GIVEN; may be
// NULL for RANGE / ZERO.  Medians are computed for the columns that contain a NaN, or for all when all_medians.
cudaError_t impute_device(ImputeWorkspace& W, double* d_m, int64_t rows, int cols, int mode, bool all_medians,
                          double* h_stats, int sm_count, cudaStream_t st, int* launches);

}  // namespace tsfx
