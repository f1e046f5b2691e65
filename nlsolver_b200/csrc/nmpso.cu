// NelderMeadPSO batches (nmpso_impl.cuh), both element types
#include "nmpso_impl.cuh"
namespace nls {
cudaError_t nmpso_launch_f64(const NMPSOState &s, cudaStream_t st) { return nmpso_launch<double>(s, st); }
cudaError_t nmpso_launch_f32(const NMPSOState &s, cudaStream_t st) { return nmpso_launch<float>(s, st); }
}  // namespace nls
