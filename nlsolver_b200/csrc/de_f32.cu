// de_f32.cu — fp32 instantiation of the DE kernels (scalar_t = float).
#include "de_impl.cuh"
namespace nls { NLS_DEFINE_DE_OPS(float, de_ops_f32) }
