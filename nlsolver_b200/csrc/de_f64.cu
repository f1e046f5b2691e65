// de_f64.cu — fp64 instantiation of the DE kernels (scalar_t = double, the reference default).
#include "de_impl.cuh"
namespace nls { NLS_DEFINE_DE_OPS(double, de_ops_f64) }
