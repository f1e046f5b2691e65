// pso_persist.cuh — the one-launch path of the PSO loop for small swarms (its own translation units).
#pragma once
#include "pso_impl.cuh"

namespace nls {

// ------------------------------------------------------------------------------------------------ one-launch path
// Small swarms are latency-bound (a generation is a few microseconds of work): all the generations of a step run in
// one launch on ONE thread-block cluster (1 .. 16 CTAs), the three phases separated by cluster barriers; the record
// of the shard goes through the swarm's own record buffer exactly as in the separate kernels, so results are identical.
template <class T, int OBJ, int TYPE, int W, int U, int S>
__global__ void __launch_bounds__(kBlock, 1) pso_persistent_kernel(PSOState s, void *record, u64 record_bytes,
                                                                   unsigned long long n_generations) {
  namespace cg = cooperative_groups;
  for (u64 g = 0; g < n_generations; g++) {
    if (*reinterpret_cast<volatile int *>(&s.ctrl->stop)) break;   // uniform: written before the last barrier
    pso_move_pass<T, OBJ, TYPE, W, U, S>(s);
    cg::this_cluster().sync();
    pso_candidate_pass<T>(s, record);
    cg::this_cluster().sync();
    if (blockIdx.x == 0) pso_apply_pass<T>(s, record, 1, record_bytes, 0);
    cg::this_cluster().sync();
  }
}

// ---- the one-launch path: n generations on one thread-block cluster
template <class T, int O, int TYPE, int W, int U, int S>
cudaError_t pso_launch_persistent_w(const PSOState &s, void *record, u64 record_bytes, unsigned long long n, cudaStream_t st) {
  auto kernel = pso_persistent_kernel<T, O, TYPE, W, U, S>;
  static signed char wide_of[64] = {};           // per DEVICE: 0 unknown, 1 allowed, -1 refused
  int dev = 0;
  cudaGetDevice(&dev);
  signed char &wide = wide_of[dev % 64];
  if (wide == 0) wide = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 1 : -1;
  const bool wide_ok = wide > 0;
  const u64 per_block = u64(kWarpsPerBlock) * (32 / W) * 2;   // two sweeps per warp and generation
  int blocks = int(std::min<u64>((s.P + per_block - 1) / per_block, wide_ok ? 16 : 8));
  if (blocks < 1) blocks = 1;
  for (;;) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kBlock);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = blocks; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, s, record, record_bytes, n);
    if (e == cudaSuccess || blocks == 1) return e;
    cudaGetLastError();
    blocks = blocks > 8 ? 8 : blocks / 2;
  }
}
template <class T, int O, int TYPE>
cudaError_t pso_launch_persistent_t(const PSOState &s, void *record, u64 rb, unsigned long long n, cudaStream_t st) {
  const u64 vecs = (s.d + Vec<T>::V - 1) / Vec<T>::V;
  if constexpr (closed_form_dim(O) > 0) return pso_launch_persistent_w<T, O, TYPE, 4, 1, 1>(s, record, rb, n, st);
  if (vecs <= 4) return pso_launch_persistent_w<T, O, TYPE, 4, 1, 1>(s, record, rb, n, st);
  if (vecs <= 8) return pso_launch_persistent_w<T, O, TYPE, 8, 1, 1>(s, record, rb, n, st);
  if (vecs <= 16) return pso_launch_persistent_w<T, O, TYPE, 16, 1, 1>(s, record, rb, n, st);
  return pso_launch_persistent_w<T, O, TYPE, 32, 1, 1>(s, record, rb, n, st);
}
template <class T>
cudaError_t pso_launch_persistent(const PSOState &s, void *record, unsigned long long record_bytes, unsigned long long n,
                                  cudaStream_t st) {
  cudaError_t e = cudaSuccess;
#define NLS_CALL(O)                                                                           \
  e = s.pso_type == 0 ? pso_launch_persistent_t<T, O, 0>(s, record, record_bytes, n, st)      \
                      : pso_launch_persistent_t<T, O, 1>(s, record, record_bytes, n, st)
  NLS_PSO_OBJ_SWITCH(s.objective, NLS_CALL)
#undef NLS_CALL
  return e;
}
#define NLS_DEFINE_PSO_PERSISTENT(T)                                                                          \
  template cudaError_t pso_launch_persistent<T>(const PSOState &s, void *record, unsigned long long record_bytes, \
                                                unsigned long long n, cudaStream_t st);

}  // namespace nls
