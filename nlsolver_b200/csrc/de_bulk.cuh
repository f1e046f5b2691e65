// de_bulk.cuh — K2 for LONG rows with the four rows of every agent staged through shared memory by the TMA unit
// (cp.async.bulk, 1-D bulk copies completing on an mbarrier; SASS: UBLKCP / SYNCS).
//
// Why: the LDG version of the generation pass keeps (resident warps) x 32 lanes x 16 B x (3 or 4 rows) in flight per
// SM — 49 KB with three DRAM streams per warp (best recombination), which at ~0.7 us of loaded DRAM latency is just
// short of what 6.5 TB/s needs, and every step pays the address arithmetic of four 128-bit loads per lane.  Here one
// elected lane per warp issues four bulk copies per CHUNK of the row (kSteps sweep steps = kSteps x 512 bytes per
// row) into a ring of kStages stages owned by the warp; the copy engine keeps (kStages - 1) chunks x 4 rows per warp in
// flight independently of registers and occupancy, the lanes read the staged rows back with conflict-free 128-bit
// shared-memory loads, and the ring runs ACROSS the agents of the warp's tile, so it only drains at tile boundaries.
// Same arithmetic, same draw tape, same summation order as de_sweep: results are bit-identical.
#pragma once
#include "common.cuh"

namespace nls {

__device__ __forceinline__ u32 smem_addr(const void *p) { return static_cast<u32>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, u32 arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on `bar` (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, u32 bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, u32 parity) {
  u32 done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  } while (!done);
}
// generic-proxy reads of a stage (by every lane, ordered by __syncwarp) before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void lds_row(const void *p, double (&x)[2]) {
  const double2 v = *reinterpret_cast<const double2 *>(p);
  x[0] = v.x; x[1] = v.y;
}
__device__ __forceinline__ void lds_row(const void *p, float (&x)[4]) {
  const float4 v = *reinterpret_cast<const float4 *>(p);
  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}

}  // namespace nls
