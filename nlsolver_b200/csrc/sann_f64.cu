// sann_f64.cu — fp64 instantiation of the annealing-chain kernels.
#include "sann_impl.cuh"
namespace nls { NLS_DEFINE_SANN_OPS(double, sann_ops_f64) }
