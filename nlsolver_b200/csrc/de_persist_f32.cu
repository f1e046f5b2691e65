#include "de_persist.cuh"
namespace nls {
NLS_DEFINE_DE_PERSISTENT(float)
}
