// launch.h — host-callable launchers implemented by the per-dtype translation units (de_f64.cu, de_f32.cu,
// pso_f64.cu, pso_f32.cu); api.cu picks the table that matches cfg.dtype.
#pragma once
#include "state.h"
namespace nls {

struct DEOps {
  cudaError_t (*init)(const DEState &s, const void *x0_dev, const LaunchGeom &g, cudaStream_t st);
  // ev: NULL, or 4 events recorded before K2, between K2/K2r, between K2r/K3 and after K3
  cudaError_t (*generation)(const DEState &s, const LaunchGeom &g, cudaStream_t st, cudaEvent_t *ev);
  cudaError_t (*export_best)(const DEState &s, void *record, cudaStream_t st);
  cudaError_t (*migrate)(const DEState &s, int sign, unsigned long long k, void *rows, void *scores,
                         const LaunchGeom &g, cudaStream_t st);
  cudaError_t (*gather_rows)(const DEState &s, unsigned long long first, unsigned long long count, void *out,
                             cudaStream_t st);
  // small populations: n generations (K2 + repair + K3 each) in ONE launch on one thread-block cluster
  cudaError_t (*persistent)(const DEState &s, unsigned long long n_generations, cudaStream_t st);
  // best re-scan without a generation (after migration / when an exchange window is attached)
  cudaError_t (*rescan)(const DEState &s, const LaunchGeom &g, cudaStream_t st);
};
struct PSOOps {
  cudaError_t (*init)(const PSOState &s, const LaunchGeom &g, cudaStream_t st);
  cudaError_t (*move)(const PSOState &s, const LaunchGeom &g, cudaStream_t st);
  cudaError_t (*candidate)(const PSOState &s, void *record, const LaunchGeom &g, cudaStream_t st);
  cudaError_t (*apply)(const PSOState &s, const void *records, unsigned long long n, unsigned long long record_bytes,
                       int initial, cudaStream_t st);
  // fused exchange over peer memory: candidate + publish to all peers, then wait for all peers + apply
  cudaError_t (*candidate_publish)(const PSOState &s, const XchgWindow &w, int initial, const LaunchGeom &g,
                                   cudaStream_t st);
  cudaError_t (*gather_apply)(const PSOState &s, const XchgWindow &w, int initial, cudaStream_t st);
  // small single-GPU swarms: n generations (move + candidate + apply each) in ONE launch on one thread-block cluster
  cudaError_t (*persistent)(const PSOState &s, void *record, unsigned long long record_bytes,
                            unsigned long long n_generations, cudaStream_t st);
  // unsharded swarms: candidate reduction and apply in one launch (the last block applies its own record)
  cudaError_t (*candidate_apply)(const PSOState &s, void *record, unsigned long long record_bytes, const LaunchGeom &g,
                                 cudaStream_t st);
};
struct SANNOps {
  // x0: device, x0_count rows of d elements (1 = shared start)
  cudaError_t (*init)(const SANNState &s, const void *x0_dev, unsigned long long x0_count, const LaunchGeom &g,
                      cudaStream_t st);
  // candidates step_begin + 1 .. step_begin + n_steps of every chain
  cudaError_t (*steps)(const SANNState &s, unsigned long long step_begin, unsigned long long n_steps,
                       const LaunchGeom &g, cudaStream_t st);
  // which: 0 best points x, 1 current points p -> dense [C][d]
  cudaError_t (*gather)(const SANNState &s, int which, void *out, const LaunchGeom &g, cudaStream_t st);
  cudaError_t (*best)(const SANNState &s, cudaStream_t st);
};
// tiny DE problems (pop_size <= 1024, dim <= 8, built-in objectives): the whole solve in one launch of one block
struct DETinyArgs;
cudaError_t de_tiny_launch_f64(int objective, const DETinyArgs &a, cudaStream_t st);
cudaError_t de_tiny_launch_f32(int objective, const DETinyArgs &a, cudaStream_t st);
struct NMPSOState;
cudaError_t nmpso_launch_f64(const NMPSOState &s, cudaStream_t st);
cudaError_t nmpso_launch_f32(const NMPSOState &s, cudaStream_t st);
const SANNOps *sann_ops_f64();
const SANNOps *sann_ops_f32();
const DEOps *de_ops_f64();
const DEOps *de_ops_f32();
const PSOOps *pso_ops_f64();
const PSOOps *pso_ops_f32();

}  // namespace nls
