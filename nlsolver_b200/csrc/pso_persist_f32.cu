#include "pso_persist.cuh"
namespace nls {
NLS_DEFINE_PSO_PERSISTENT(float)
}
