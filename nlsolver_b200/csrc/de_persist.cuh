// de_persist.cuh — the one-launch path of the DE loop for small populations (its own translation units: the kernels
// inline all three passes of a generation, which is a lot of code per objective and lane-group shape).
#pragma once
#include "de_impl.cuh"

namespace nls {

// ------------------------------------------------------------------------------------------------ one-launch path
// Small populations (P * d <= 2^16 elements) are latency-bound: a generation is a few microseconds of work, and three
// launches (or three graph nodes) plus the grid barriers of the repair cost more than the work.  This kernel runs ALL
// the generations of a step in one launch on ONE thread-block cluster (1 .. 16 CTAs, sized to the population): the same
// three passes, separated by cluster barriers (hardware barrier, no global-memory round trip), with the loop state
// (iter, best_id, stop) read back from the control block after each barrier.  Generations after a stop rule fired are
// skipped, exactly as the separate kernels return early.
template <class T, int OBJ, int W, int U, int S, bool SKIP_BASE>
__global__ void __launch_bounds__(kBlock, 1) de_persistent_kernel(DEState s, unsigned long long n_generations) {
  __shared__ DETileEntry tile_mem[kWarpsPerBlock][32];
  DETileEntry *tile_entries = tile_mem[threadIdx.x >> 5];
  ClusterSync sync;
  const u64 n_warps = (u64(gridDim.x) * kBlock) >> 5;
  int shift = 5;                                           // every warp should get a tile
  while (shift > 0 && ((s.P + (1ull << shift) - 1) >> shift) < n_warps) shift--;
  for (u64 g = 0; g < n_generations; g++) {
    if (*reinterpret_cast<volatile int *>(&s.ctrl->stop)) break;   // uniform: written before the last barrier
    de_generation_pass<T, OBJ, W, U, S, SKIP_BASE>(s, tile_entries, shift);
    sync();
    de_repair_pass<T, OBJ, W, U, S, SKIP_BASE>(s, tile_entries, sync);
    sync();
    de_commit_pass<T>(s, 0);
    sync();
  }
}

// ---- the one-launch path: n generations on one thread-block cluster
constexpr int kMaxClusterBlocks = 16;                      // > 8 needs the non-portable opt-in (B200 allows 16)
template <class T, int O, int W, int U, int S, bool SKIP_BASE>
cudaError_t de_launch_persistent_w(const DEState &s, unsigned long long n_generations, cudaStream_t st) {
  auto kernel = de_persistent_kernel<T, O, W, U, S, SKIP_BASE>;
  static signed char wide_of[kMaxDevices] = {};           // per DEVICE: 0 unknown, 1 allowed, -1 refused
  int dev = 0;
  cudaGetDevice(&dev);
  signed char &wide = wide_of[dev % kMaxDevices];
  if (wide == 0) wide = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 1 : -1;
  const bool wide_ok = wide > 0;
  // one warp streams 32 / W agents at a time; aim at two such sweeps per warp and generation
  const u64 per_block = u64(kWarpsPerBlock) * (32 / W) * 2;
  int blocks = int(std::min<u64>((s.P + per_block - 1) / per_block, wide_ok ? kMaxClusterBlocks : 8));
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, s, n_generations);
    if (e == cudaSuccess || blocks == 1) return e;
    cudaGetLastError();                                  // a cluster this wide cannot be placed: halve it
    blocks = blocks > 8 ? 8 : blocks / 2;
  }
}
template <class T, int O>
cudaError_t de_launch_persistent_o(const DEState &s, unsigned long long n, cudaStream_t st) {
  const u64 vecs = (s.d + Vec<T>::V - 1) / Vec<T>::V;
  if constexpr (closed_form_dim(O) > 0) return de_launch_persistent_w<T, O, 4, 1, 1, false>(s, n, st);
  if (vecs <= 4) return de_launch_persistent_w<T, O, 4, 1, 1, false>(s, n, st);
  if (vecs <= 8) return de_launch_persistent_w<T, O, 8, 1, 1, false>(s, n, st);
  if (vecs <= 16) return de_launch_persistent_w<T, O, 16, 1, 1, false>(s, n, st);
  return de_launch_persistent_w<T, O, 32, 1, 1, true>(s, n, st);
}
template <class T>
cudaError_t de_launch_persistent(const DEState &s, unsigned long long n_generations, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
#define NLS_CALL(O) e = de_launch_persistent_o<T, O>(s, n_generations, st)
  NLS_OBJ_SWITCH(s.objective, NLS_CALL)
#undef NLS_CALL
  return e;
}

#define NLS_DEFINE_DE_PERSISTENT(T) \
  template cudaError_t de_launch_persistent<T>(const DEState &s, unsigned long long n_generations, cudaStream_t st);

}  // namespace nls
