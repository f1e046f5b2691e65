// the one-block solver for tiny DE problems (de_tiny.cuh), both element types
#include "de_tiny.cuh"
#include "launch.h"
namespace nls {
cudaError_t de_tiny_launch_f64(int objective, const DETinyArgs &a, cudaStream_t st) { return de_tiny_launch<double>(objective, a, st); }
cudaError_t de_tiny_launch_f32(int objective, const DETinyArgs &a, cudaStream_t st) { return de_tiny_launch<float>(objective, a, st); }
}  // namespace nls
