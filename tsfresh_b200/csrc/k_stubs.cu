// temporary stubs for kernel groups not yet implemented
#include "tsfx_kernels.h"
namespace tsfx {
cudaError_t launch_sorted(const SortedArgs&, int, cudaStream_t, int) { return cudaErrorNotSupported; }
cudaError_t launch_spectral(const SpectralArgs&, int, cudaStream_t, int) { return cudaErrorNotSupported; }
cudaError_t launch_la(const LaArgs&, int, cudaStream_t, int) { return cudaErrorNotSupported; }
cudaError_t launch_entropy(const EntropyArgs&, int, cudaStream_t, int) { return cudaErrorNotSupported; }
cudaError_t launch_seq(const SeqArgs&, int, cudaStream_t, int) { return cudaErrorNotSupported; }
cudaError_t launch_fill_twiddle(double2*, int, cudaStream_t) { return cudaErrorNotSupported; }
}
