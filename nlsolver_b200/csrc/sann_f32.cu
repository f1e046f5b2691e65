// sann_f32.cu — fp32 instantiation of the annealing-chain kernels.
#include "sann_impl.cuh"
namespace nls { NLS_DEFINE_SANN_OPS(float, sann_ops_f32) }
