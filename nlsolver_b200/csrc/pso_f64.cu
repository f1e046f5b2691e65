// pso_f64.cu — fp64 instantiation of the PSO kernels.
#include "pso_impl.cuh"
namespace nls { NLS_DEFINE_PSO_OPS(double, pso_ops_f64) }
