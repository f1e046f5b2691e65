// pso_f32.cu — fp32 instantiation of the PSO kernels.
#include "pso_impl.cuh"
namespace nls { NLS_DEFINE_PSO_OPS(float, pso_ops_f32) }
