// api.cu — host side of the C ABI declared in include/nls_b200.h.
//
// Owns every device allocation behind opaque handles, validates arguments (the reference validates nothing:
// pop_size < 4 loops forever in generate_indices, nlsolver.h:2344-2354), turns CUDA errors into return codes and
// enqueues the kernels of de_impl.cuh / pso_impl.cuh / sann_impl.cuh on the context stream.  No CPU compute path exists here.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/nls_b200.h"
#include "common.cuh"
#include "de_tiny.cuh"
#include "launch.h"
#include "nmpso_impl.cuh"
#include "reduce.cuh"
#include "state.h"

using namespace nls;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define NLS_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return fail(e__ == cudaErrorMemoryAllocation ? NLS_ERR_NOMEM : NLS_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                         \
  } while (0)

size_t elem_size(int dtype) { return dtype == NLS_F64 ? 8 : 4; }

// objective plugins (objective_plugin.cuh): same ops tables as the built-in dtypes, instantiated for a user functor
struct ObjectivePlugin {
  int abi;
  unsigned full_dim;
  unsigned de_state_bytes, pso_state_bytes, sann_state_bytes, _pad;
  const DEOps *de_f64, *de_f32;
  const PSOOps *pso_f64, *pso_f32;
  const SANNOps *sann_f64, *sann_f32;
};
constexpr int kFirstPluginId = 100;
constexpr int kPluginAbi = 5;
constexpr int kMaxPlugins = 256;
// fixed-capacity registry: entries are written once under the mutex and published by the count, so lookups from solver
// threads never see a reallocating container
std::mutex g_plugin_mutex;
const ObjectivePlugin *g_plugins[kMaxPlugins];
std::atomic<int> g_plugin_count{0};
const ObjectivePlugin *plugin_for(int objective) {
  const int k = objective - kFirstPluginId;
  return (k >= 0 && k < g_plugin_count.load(std::memory_order_acquire)) ? g_plugins[k] : nullptr;
}
bool objective_known(int objective) {
  return (objective >= 0 && objective < NLS_OBJECTIVE_COUNT) || plugin_for(objective) != nullptr;
}
// dimension a closed-form objective is defined for (0: any)
unsigned objective_fixed_dim(int objective) {
  if (const ObjectivePlugin *pl = plugin_for(objective)) return pl->full_dim;
  if (objective == NLS_SHEKEL) return 4;
  if (objective >= NLS_BEALE && objective <= NLS_LEVI_N13 && objective != NLS_STYBLINSKI_TANG) return 2;
  return 0;
}
u64 round_up(u64 v, u64 m) { return (v + m - 1) / m * m; }

// host mirror of unit<T>() for the crossover threshold search
template <class T> T host_unit(u64 u) { return static_cast<T>(u / static_cast<T>(18446744073709551615U)); }

// `unit(raw) < CR` is monotone in raw, so it equals `raw <= cr_le` for the largest raw that still satisfies it.
template <class T>
void crossover_threshold(double cr, u64 *cr_le, int *cr_none) {
  const T c = static_cast<T>(cr);
  if (!(host_unit<T>(0) < c)) { *cr_none = 1; *cr_le = 0; return; }
  *cr_none = 0;
  u64 lo = 0, hi = ~0ull;                 // invariant: unit(lo) < c
  if (host_unit<T>(hi) < c) { *cr_le = hi; return; }
  while (hi - lo > 1) {                   // unit(hi) >= c
    const u64 mid = lo + (hi - lo) / 2;
    if (host_unit<T>(mid) < c) lo = mid; else hi = mid;
  }
  *cr_le = lo;
}

}  // namespace

struct nls_ctx {
  int device;
  cudaStream_t stream;
  bool own_stream;
  int sm_count;
  // Device buffers released by destroyed solver handles, kept for the next solve: cudaMalloc / cudaFree of multi-GB
  // populations cost hundreds of milliseconds, which would dominate repeated minimize() calls.  The pool is bounded:
  // it never holds more than `pool_limit` bytes (nls_ctx_set_pool_limit), by default no more than the largest single
  // solver handle released so far, and the oldest entries go first — a sequence of solves of different shapes cannot
  // pin every population it ever used.
  std::vector<std::pair<void *, size_t>> pool;   // oldest first
  size_t pool_bytes = 0;
  void *tiny_host = nullptr;    // mapped host memory the one-block tiny solver writes its result to (de_tiny.cuh)
  void *tiny_dev = nullptr;
  size_t pool_limit = 0;        // 0: automatic (the largest single handle released so far)
  size_t largest_handle = 0;
  void evict_to(size_t cap) {
    size_t k = 0;
    while (pool_bytes > cap && k < pool.size()) { cudaFree(pool[k].first); pool_bytes -= pool[k].second; k++; }
    pool.erase(pool.begin(), pool.begin() + k);
  }
};

namespace {

// Debug aid (NLS_B200_GUARD=1 in the environment): every device buffer gets a 256-byte guard zone on either side, filled
// with a pattern when the buffer is created and verified when its solver handle is destroyed; a kernel that stores
// outside its buffers shows up in nls_debug_guard_violations(), one that reads outside them computes NaNs and fails the
// parity tests.  (compute-sanitizer is not available everywhere.)
constexpr size_t kGuardBytes = 256;
constexpr unsigned char kGuardPattern = 0xFF;   // NaN as fp32 / fp64: a read that strays into a guard zone poisons the result
bool guard_mode() {
  static const bool on = [] { const char *e = std::getenv("NLS_B200_GUARD"); return e && e[0] == '1'; }();
  return on;
}
unsigned long long g_guard_violations = 0;

struct DeviceBuffers {
  nls_ctx *ctx = nullptr;
  std::vector<std::pair<void *, size_t>> held;
  int alloc(void **out, size_t bytes) {
    *out = nullptr;
    bytes = bytes ? bytes : 1;
    if (guard_mode()) {                        // never pooled: base = user pointer - kGuardBytes
      char *base = nullptr;
      NLS_CUDA(cudaMalloc(reinterpret_cast<void **>(&base), bytes + 2 * kGuardBytes));
      NLS_CUDA(cudaMemset(base, kGuardPattern, kGuardBytes));
      NLS_CUDA(cudaMemset(base + kGuardBytes + bytes, kGuardPattern, kGuardBytes));
      *out = base + kGuardBytes;
      held.push_back({*out, bytes});
      return NLS_OK;
    }
    if (ctx) {
      // best fit: the smallest cached buffer that holds the request without wasting more than half of itself
      size_t best = ctx->pool.size();
      for (size_t k = 0; k < ctx->pool.size(); k++) {
        const size_t have = ctx->pool[k].second;
        if (have >= bytes && have - bytes <= std::max<size_t>(bytes, size_t(1) << 20) &&
            (best == ctx->pool.size() || have < ctx->pool[best].second))
          best = k;
      }
      if (best < ctx->pool.size()) {
        *out = ctx->pool[best].first;
        held.push_back(ctx->pool[best]);          // remembered with its true size
        ctx->pool_bytes -= ctx->pool[best].second;
        ctx->pool.erase(ctx->pool.begin() + best);
        return NLS_OK;
      }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation && ctx && !ctx->pool.empty()) {   // make room: drop the cached buffers, retry
      cudaGetLastError();
      for (auto &b : ctx->pool) cudaFree(b.first);
      ctx->pool.clear();
      ctx->pool_bytes = 0;
      e = cudaMalloc(out, bytes);
    }
    NLS_CUDA(e);
    held.push_back({*out, bytes});
    return NLS_OK;
  }
  void release() {
    size_t total = 0;
    for (auto &b : held) total += b.second;
    if (ctx && !guard_mode()) ctx->largest_handle = std::max(ctx->largest_handle, total);
    for (auto &b : held) {
      if (guard_mode()) {
        char *base = static_cast<char *>(b.first) - kGuardBytes;
        unsigned char zone[2 * kGuardBytes];
        bool ok = cudaMemcpy(zone, base, kGuardBytes, cudaMemcpyDeviceToHost) == cudaSuccess &&
                  cudaMemcpy(zone + kGuardBytes, base + kGuardBytes + b.second, kGuardBytes, cudaMemcpyDeviceToHost) == cudaSuccess;
        for (size_t k = 0; ok && k < 2 * kGuardBytes; k++) ok = zone[k] == kGuardPattern;
        if (!ok) {
          g_guard_violations++;
          std::fprintf(stderr, "nls_b200: guard zone of a %zu-byte device buffer was overwritten\n", b.second);
        }
        cudaFree(base);
        continue;
      }
      if (ctx) { ctx->pool.push_back(b); ctx->pool_bytes += b.second; }
      else cudaFree(b.first);
    }
    held.clear();
    if (ctx) ctx->evict_to(ctx->pool_limit ? ctx->pool_limit : ctx->largest_handle);
  }
};

LaunchGeom make_geom(const nls_ctx *ctx, u64 n) {
  LaunchGeom g;
  g.sm_count = ctx->sm_count;
  const u64 want = (n + kBlock - 1) / kBlock, cap = u64(ctx->sm_count) * 4;
  g.reduce_blocks = int(want < cap ? (want < 1 ? 1 : want) : cap);
  return g;
}

}  // namespace

// Small populations are launch-bound (a generation is three kernels of a few microseconds each): their generations are
// replayed from a CUDA graph of kGraphGens generations instead of being launched one kernel at a time.
constexpr int kGraphGens = 8;
constexpr unsigned long long kGraphMaxElems = 1ull << 24;   // P * d above which launch overhead no longer matters
struct GraphCache {
  cudaGraphExec_t exec = nullptr;
  bool failed = false;
  void reset() { if (exec) cudaGraphExecDestroy(exec); exec = nullptr; }
};
// capture `body` (which enqueues kGraphGens generations on `st`) once, then replay it
template <class Body>
static bool graph_replay(GraphCache &gc, cudaStream_t st, Body body) {
  if (gc.failed) return false;
  if (!gc.exec) {
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); gc.failed = true; return false; }
    const bool ok = body();
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (!ok || e != cudaSuccess || !graph || cudaGraphInstantiate(&gc.exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      gc.exec = nullptr;
      gc.failed = true;
      return false;
    }
    cudaGraphDestroy(graph);
  }
  return cudaGraphLaunch(gc.exec, st) == cudaSuccess;
}

struct nls_de {
  nls_ctx *ctx;
  nls_de_cfg cfg;
  DEState s;
  LaunchGeom g;
  const DEOps *ops;
  DeviceBuffers mem;
  void *record;       // internal exchange record
  void *staging;      // read-back staging
  size_t staging_bytes;
  size_t elem;
  size_t topk_capacity;                 // bytes behind s.topk_scratch (migration top-k candidates)
  bool timing;                          // measurement hook (nls_de_enable_kernel_timing)
  std::vector<cudaEvent_t> events;      // 4 per timed generation
  double timed_ms[3];
  u64 timed_generations;
  GraphCache graph;
  bool persistent_failed = false;
  struct nls_xchg *xchg = nullptr;      // islands: window the commit kernel publishes into
};

struct nls_xchg {
  nls_ctx *ctx;
  XchgWindow w;
  void *base;                       // this rank's window: records [2][world][record_bytes], then flags [2][world]
  size_t bytes, flags_offset;
  void *peer_base[kMaxPeers];       // IPC mappings of the peers' windows (NULL for self / not opened)
  bool opened;
  bool attached;                    // sequence flags only ever grow: a window serves exactly one swarm
};

struct nls_pso {
  nls_ctx *ctx;
  nls_pso_cfg cfg;
  PSOState s;
  LaunchGeom g;
  const PSOOps *ops;
  DeviceBuffers mem;
  void *record;
  u64 record_bytes;
  u64 enqueued;          // generations enqueued so far (the iter value the next move kernel will see)
  bool first_apply_pending;
  nls_xchg *xchg;        // fused peer exchange, or NULL
  GraphCache graph;
  size_t elem;
  bool persistent_failed = false;
};

extern "C" {

const char *nls_last_error(void) { return g_err.c_str(); }
int nls_version(void) { return NLS_B200_VERSION; }

int nls_ctx_create(int device, void *stream, nls_ctx **out) {
  if (!out) return fail(NLS_ERR_INVALID, "nls_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(NLS_ERR_CUDA, "nls_ctx_create: no CUDA device (%s); this library has no CPU path",
                cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(NLS_ERR_INVALID, "nls_ctx_create: device %d out of range [0,%d)", device, n);
  NLS_CUDA(cudaSetDevice(device));
  nls_ctx *c = new nls_ctx();
  c->device = device;
  c->own_stream = stream == nullptr;
  c->stream = static_cast<cudaStream_t>(stream);
  if (c->own_stream) {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(NLS_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  }
  int coop = 0;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  if (!coop) { delete c; return fail(NLS_ERR_CUDA, "device %d lacks cooperative launch", device); }
  *out = c;
  return NLS_OK;
}

int nls_ctx_destroy(nls_ctx *ctx) {
  if (!ctx) return NLS_OK;
  cudaSetDevice(ctx->device);
  for (auto &b : ctx->pool) cudaFree(b.first);
  if (ctx->tiny_host) cudaFreeHost(ctx->tiny_host);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return NLS_OK;
}
int nls_ctx_trim(nls_ctx *ctx) {
  if (!ctx) return fail(NLS_ERR_INVALID, "nls_ctx_trim: NULL context");
  NLS_CUDA(cudaSetDevice(ctx->device));
  NLS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &b : ctx->pool) cudaFree(b.first);
  ctx->pool.clear();
  ctx->pool_bytes = 0;
  return NLS_OK;
}
int nls_ctx_set_pool_limit(nls_ctx *ctx, uint64_t bytes) {
  if (!ctx) return fail(NLS_ERR_INVALID, "nls_ctx_set_pool_limit: NULL context");
  NLS_CUDA(cudaSetDevice(ctx->device));
  ctx->pool_limit = bytes;
  if (bytes) ctx->evict_to(bytes);
  return NLS_OK;
}
uint64_t nls_ctx_pool_bytes(const nls_ctx *ctx) { return ctx ? ctx->pool_bytes : 0; }
unsigned long long nls_debug_guard_violations(void) { return g_guard_violations; }
int nls_ctx_device(const nls_ctx *ctx) { return ctx ? ctx->device : -1; }
int nls_ctx_sm_count(const nls_ctx *ctx) { return ctx ? ctx->sm_count : -1; }

uint64_t nls_record_bytes(int32_t dtype, uint64_t dim) {
  return sizeof(RecordHeader) + round_up(dim * elem_size(dtype), 8);
}

/* ================================================================ DE ========================================= */

static int de_validate(const nls_de_cfg *c) {
  if (c->dtype != NLS_F32 && c->dtype != NLS_F64) return fail(NLS_ERR_INVALID, "DE: unknown dtype %d", c->dtype);
  if (!objective_known(c->objective)) return fail(NLS_ERR_INVALID, "DE: unknown objective %d", c->objective);
  if (const unsigned fd = objective_fixed_dim(c->objective))
    if (fd != c->dim)
      return fail(NLS_ERR_INVALID, "DE: objective %d is a closed form of dimension %u, dim is %llu", c->objective, fd,
                  static_cast<unsigned long long>(c->dim));
  if (c->strategy != NLS_DE_BEST && c->strategy != NLS_DE_RANDOM) return fail(NLS_ERR_INVALID, "DE: unknown strategy %d", c->strategy);
  if (c->pop_size < 4) return fail(NLS_ERR_INVALID, "DE: pop_size must be >= 4 (three distinct donors besides the fixed agent)");
  if (c->dim < 1) return fail(NLS_ERR_INVALID, "DE: dim must be >= 1");
  if (c->pop_size >= 0xffffffffull || c->dim >= 0xffffffffull) return fail(NLS_ERR_INVALID, "DE: pop_size and dim must fit 32 bits");
  return NLS_OK;
}

int nls_de_destroy(nls_de *de) {
  if (!de) return NLS_OK;
  cudaSetDevice(de->ctx->device);
  cudaStreamSynchronize(de->ctx->stream);
  for (cudaEvent_t e : de->events) cudaEventDestroy(e);
  de->graph.reset();
  de->mem.release();
  delete de;
  return NLS_OK;
}

static int de_build(nls_ctx *ctx, const nls_de_cfg *cfg, const void *x0_host, nls_de *de) {
  const u64 P = cfg->pop_size, d = cfg->dim;
  de->ctx = ctx;
  de->mem.ctx = ctx;
  de->cfg = *cfg;
  de->timing = false;
  de->topk_capacity = 0;
  de->timed_ms[0] = de->timed_ms[1] = de->timed_ms[2] = 0.0;
  de->timed_generations = 0;
  de->elem = elem_size(cfg->dtype);
  de->ops = cfg->dtype == NLS_F64 ? de_ops_f64() : de_ops_f32();
  if (const ObjectivePlugin *pl = plugin_for(cfg->objective)) de->ops = cfg->dtype == NLS_F64 ? pl->de_f64 : pl->de_f32;
  de->g = make_geom(ctx, P);
  DEState &s = de->s;
  std::memset(&s, 0, sizeof(s));
  s.P = P; s.d = d;
  s.stride = round_up(d, 32 / de->elem);            // rows start on 32-byte sector boundaries
  s.seed = cfg->seed; s.offset = cfg->agent_offset;
  s.strategy = cfg->strategy;
  s.objective = plugin_for(cfg->objective) ? 100 /* OBJ_CUSTOM */ : cfg->objective;
  s.F = cfg->differential_weight; s.fm = cfg->minimize ? 1.0 : -1.0; s.eps = cfg->eps;
  s.max_iter = cfg->max_iter; s.vnc_limit = cfg->best_val_no_change;
  u64 cr_le; int cr_none;
  if (cfg->dtype == NLS_F64) crossover_threshold<double>(cfg->crossover_prob, &cr_le, &cr_none);
  else crossover_threshold<float>(cfg->crossover_prob, &cr_le, &cr_none);
  s.cr_le = cr_le; s.cr_none = cr_none;

  const size_t row_bytes = size_t(P) * s.stride * de->elem;
  int rc;
  void *x0_dev = nullptr;
#define NLS_ALLOC(ptr, bytes) if ((rc = de->mem.alloc(reinterpret_cast<void **>(&(ptr)), (bytes))) != NLS_OK) return rc
  NLS_ALLOC(s.buf[0], row_bytes);
  NLS_ALLOC(s.buf[1], row_bytes);
  NLS_ALLOC(s.where, P);
  NLS_ALLOC(s.score, P * de->elem);
  NLS_ALLOC(s.tscore, P * de->elem);
  NLS_ALLOC(s.acc, P);
  NLS_ALLOC(s.fin, P * sizeof(uint16_t));
  NLS_ALLOC(s.dec, P * sizeof(uint4));
  NLS_ALLOC(s.rej, P * sizeof(uint32_t));
  NLS_ALLOC(s.list, P * sizeof(uint32_t));
  s.coarse_words = (P + 1023) / 1024;
  NLS_ALLOC(s.coarse, 3 * s.coarse_words * sizeof(uint32_t));
  NLS_ALLOC(s.ctrl, sizeof(DECtrl));
  // (the one-launch path reduces with up to 16 blocks whatever the population)
  const size_t n_part = std::max(de->g.reduce_blocks, 16);
  NLS_ALLOC(s.part_min, n_part * sizeof(double));
  NLS_ALLOC(s.part_idx, n_part * sizeof(unsigned long long));
  NLS_ALLOC(s.part_mom, n_part * sizeof(Moments));
  if (cfg->flags & NLS_FLAG_RECORD_MASKS) NLS_ALLOC(s.masks, P * d);
  NLS_ALLOC(de->record, nls_record_bytes(cfg->dtype, d));
  // read-back staging: up to 256 MB, but never less than one row (rows are gathered whole)
  de->staging_bytes = std::max<size_t>(std::min<size_t>(size_t(P) * d * de->elem, size_t(256) << 20), size_t(d) * de->elem);
  NLS_ALLOC(de->staging, de->staging_bytes);
  NLS_ALLOC(x0_dev, d * de->elem);
#undef NLS_ALLOC
  cudaStream_t st = ctx->stream;
  NLS_CUDA(cudaMemsetAsync(s.ctrl, 0, sizeof(DECtrl), st));
  NLS_CUDA(cudaMemsetAsync(s.fin, 0, P * sizeof(uint16_t), st));
  NLS_CUDA(cudaMemsetAsync(s.dec, 0, P * sizeof(uint4), st));
  NLS_CUDA(cudaMemsetAsync(s.rej, 0, P * sizeof(uint32_t), st));
  NLS_CUDA(cudaMemsetAsync(s.coarse, 0, 3 * s.coarse_words * sizeof(uint32_t), st));
  if (s.masks) NLS_CUDA(cudaMemsetAsync(s.masks, 0, P * d, st));
  NLS_CUDA(cudaMemcpyAsync(x0_dev, x0_host, d * de->elem, cudaMemcpyHostToDevice, st));
  NLS_CUDA(de->ops->init(s, x0_dev, de->g, st));
  NLS_CUDA(cudaStreamSynchronize(st));
  return NLS_OK;
}

int nls_de_create(nls_ctx *ctx, const nls_de_cfg *cfg, const void *x0_host, nls_de **out) {
  if (!ctx || !cfg || !x0_host || !out) return fail(NLS_ERR_INVALID, "nls_de_create: NULL argument");
  *out = nullptr;
  int rc = de_validate(cfg);
  if (rc != NLS_OK) return rc;
  NLS_CUDA(cudaSetDevice(ctx->device));
  nls_de *de = new nls_de();
  rc = de_build(ctx, cfg, x0_host, de);
  if (rc != NLS_OK) { de->mem.release(); delete de; return rc; }
  *out = de;
  return NLS_OK;
}

// Small populations can take the one-launch path: all the generations of a step in one kernel on one thread-block
// cluster (de_persistent_kernel / pso_persistent_kernel).  Measured on B200 (tools/bench_small.py): 27 against 36 us per
// generation at pop 50 x 2, but no gain over the graph replay at 1024 x 64 and beyond (16 us either way: the latency
// chain of the three passes is the same, only the launches go), and slower for the accelerated move.  So by default
// only populations of up to kPersistDefaultAgents agents take it; NLS_DE_ONE_LAUNCH=1 in the environment sends every
// population of up to kPersistMaxElems elements there, NLS_DE_ONE_LAUNCH=0 none.
constexpr unsigned long long kPersistMaxElems = 1ull << 16;
constexpr unsigned long long kPersistDefaultAgents = 256;
static bool de_one_launch_enabled(unsigned long long agents) {
  const char *e = std::getenv("NLS_DE_ONE_LAUNCH");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;
  return agents <= kPersistDefaultAgents;
}

int nls_de_step(nls_de *de, uint64_t n_generations) {
  if (!de) return fail(NLS_ERR_INVALID, "nls_de_step: NULL handle");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  uint64_t left = n_generations;
  if (!de->timing && left > 0 && de->s.P * de->s.d <= kPersistMaxElems && de->ops->persistent && !de->persistent_failed &&
      de_one_launch_enabled(de->s.P)) {
    if (de->ops->persistent(de->s, left, de->ctx->stream) == cudaSuccess) return NLS_OK;
    cudaGetLastError();
    de->persistent_failed = true;                 // e.g. no cluster launch on this device: the graph path still works
  }
  if (!de->timing && de->s.P * de->s.d <= kGraphMaxElems) {
    while (left >= kGraphGens) {
      const bool ok = graph_replay(de->graph, de->ctx->stream, [&] {
        for (int g = 0; g < kGraphGens; g++)
          if (de->ops->generation(de->s, de->g, de->ctx->stream, nullptr) != cudaSuccess) return false;
        return true;
      });
      if (!ok) break;
      left -= kGraphGens;
    }
  }
  for (uint64_t g = 0; g < left; g++) {
    cudaEvent_t *ev = nullptr;
    if (de->timing) {
      const size_t base = de->events.size();
      de->events.resize(base + 4);
      for (int k = 0; k < 4; k++) NLS_CUDA(cudaEventCreate(&de->events[base + k]));
      ev = &de->events[base];
    }
    NLS_CUDA(de->ops->generation(de->s, de->g, de->ctx->stream, ev));
  }
  return NLS_OK;
}

static int de_drain_events(nls_de *de) {
  for (size_t b = 0; b + 4 <= de->events.size(); b += 4) {
    NLS_CUDA(cudaEventSynchronize(de->events[b + 3]));
    for (int k = 0; k < 3; k++) {
      float ms = 0.f;
      NLS_CUDA(cudaEventElapsedTime(&ms, de->events[b + k], de->events[b + k + 1]));
      de->timed_ms[k] += ms;
    }
    de->timed_generations++;
    for (int k = 0; k < 4; k++) cudaEventDestroy(de->events[b + k]);
  }
  de->events.clear();
  return NLS_OK;
}

int nls_de_enable_kernel_timing(nls_de *de, int enable) {
  if (!de) return fail(NLS_ERR_INVALID, "nls_de_enable_kernel_timing: NULL handle");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  int rc = de_drain_events(de);
  de->timing = enable != 0;
  return rc;
}

int nls_de_kernel_times(nls_de *de, double ms[3], uint64_t *generations) {
  if (!de || !ms || !generations) return fail(NLS_ERR_INVALID, "nls_de_kernel_times: NULL argument");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  int rc = de_drain_events(de);
  if (rc != NLS_OK) return rc;
  for (int k = 0; k < 3; k++) { ms[k] = de->timed_ms[k]; de->timed_ms[k] = 0.0; }
  *generations = de->timed_generations;
  de->timed_generations = 0;
  return NLS_OK;
}

static int de_status(nls_de *de, nls_status *status) {
  DECtrl c;
  NLS_CUDA(cudaMemcpyAsync(&c, de->s.ctrl, sizeof(c), cudaMemcpyDeviceToHost, de->ctx->stream));
  NLS_CUDA(cudaStreamSynchronize(de->ctx->stream));
  if (c.error) return fail(NLS_ERR_INTERNAL, "DE in-place repair did not converge");
  if (status) {
    std::memset(status, 0, sizeof(*status));
    status->f_value = c.best_value;
    status->iterations = c.iter;
    status->function_calls = de->s.P * (c.iter + 1);
    status->best_index = c.best_id;
    status->val_no_change = c.vnc;
    status->stopped = c.stop;
    status->stop_reason = c.stop_reason;
    status->best_valid = 1;
    status->std_err = c.std_err;
    status->repair_reruns = c.reruns;
    status->repair_rounds = c.rounds;
    status->accepted_total = c.accepted;
  }
  return NLS_OK;
}

// re-scan the best and (with an exchange window attached) publish the island's record again
static int de_republish(nls_de *de) {
  NLS_CUDA(de->ops->rescan(de->s, de->g, de->ctx->stream));
  return NLS_OK;
}

int nls_de_sync(nls_de *de, nls_status *status) {
  if (!de) return fail(NLS_ERR_INVALID, "nls_de_sync: NULL handle");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  return de_status(de, status);
}

static int de_read_rows(nls_de *de, u64 first, u64 count, void *host) {
  const size_t row = de->s.d * de->elem;
  const u64 per_chunk = std::max<u64>(1, de->staging_bytes / row);
  char *dst = static_cast<char *>(host);
  for (u64 off = 0; off < count; off += per_chunk) {
    const u64 n = std::min(per_chunk, count - off);
    NLS_CUDA(de->ops->gather_rows(de->s, first + off, n, de->staging, de->ctx->stream));
    NLS_CUDA(cudaMemcpyAsync(dst + off * row, de->staging, n * row, cudaMemcpyDeviceToHost, de->ctx->stream));
    NLS_CUDA(cudaStreamSynchronize(de->ctx->stream));
  }
  return NLS_OK;
}

int nls_de_read_best(nls_de *de, void *x_host) {
  if (!de || !x_host) return fail(NLS_ERR_INVALID, "nls_de_read_best: NULL argument");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  nls_status st;
  int rc = de_status(de, &st);
  if (rc != NLS_OK) return rc;
  return de_read_rows(de, st.best_index, 1, x_host);
}

int nls_de_read_population(nls_de *de, void *rows_host) {
  if (!de || !rows_host) return fail(NLS_ERR_INVALID, "nls_de_read_population: NULL argument");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  NLS_CUDA(cudaStreamSynchronize(de->ctx->stream));
  return de_read_rows(de, 0, de->s.P, rows_host);
}

int nls_de_read_rows(nls_de *de, uint64_t first, uint64_t count, void *rows_host) {
  if (!de || !rows_host) return fail(NLS_ERR_INVALID, "nls_de_read_rows: NULL argument");
  if (count > de->s.P || first > de->s.P - count) return fail(NLS_ERR_INVALID, "nls_de_read_rows: range exceeds the population");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  NLS_CUDA(cudaStreamSynchronize(de->ctx->stream));
  return de_read_rows(de, first, count, rows_host);
}

int nls_de_read_scores(nls_de *de, void *scores_host) {
  if (!de || !scores_host) return fail(NLS_ERR_INVALID, "nls_de_read_scores: NULL argument");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  NLS_CUDA(cudaMemcpyAsync(scores_host, de->s.score, de->s.P * de->elem, cudaMemcpyDeviceToHost, de->ctx->stream));
  NLS_CUDA(cudaStreamSynchronize(de->ctx->stream));
  return NLS_OK;
}

int nls_de_read_decisions(nls_de *de, uint32_t *donors, uint32_t *dim_idx, uint32_t *rejects, uint8_t *accepted,
                          void *trial_scores, uint8_t *masks) {
  if (!de) return fail(NLS_ERR_INVALID, "nls_de_read_decisions: NULL handle");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  cudaStream_t st = de->ctx->stream;
  const u64 P = de->s.P;
  if (donors || dim_idx) {
    std::vector<uint4> dec(P);
    NLS_CUDA(cudaMemcpyAsync(dec.data(), de->s.dec, P * sizeof(uint4), cudaMemcpyDeviceToHost, st));
    NLS_CUDA(cudaStreamSynchronize(st));
    for (u64 i = 0; i < P; i++) {
      if (donors) { donors[3 * i] = dec[i].x; donors[3 * i + 1] = dec[i].y; donors[3 * i + 2] = dec[i].z; }
      if (dim_idx) dim_idx[i] = dec[i].w;
    }
  }
  if (rejects) NLS_CUDA(cudaMemcpyAsync(rejects, de->s.rej, P * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (accepted) NLS_CUDA(cudaMemcpyAsync(accepted, de->s.acc, P, cudaMemcpyDeviceToHost, st));
  if (trial_scores) NLS_CUDA(cudaMemcpyAsync(trial_scores, de->s.tscore, P * de->elem, cudaMemcpyDeviceToHost, st));
  if (masks) {
    if (!de->s.masks) return fail(NLS_ERR_STATE, "masks were not recorded: create the handle with NLS_FLAG_RECORD_MASKS");
    NLS_CUDA(cudaMemcpyAsync(masks, de->s.masks, P * de->s.d, cudaMemcpyDeviceToHost, st));
  }
  NLS_CUDA(cudaStreamSynchronize(st));
  return NLS_OK;
}

int nls_de_export_best(nls_de *de, void *record_dev) {
  if (!de || !record_dev) return fail(NLS_ERR_INVALID, "nls_de_export_best: NULL argument");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  NLS_CUDA(de->ops->export_best(de->s, record_dev, de->ctx->stream));
  return NLS_OK;
}
static int de_ensure_topk_scratch(nls_de *de, uint64_t k) {
  const size_t need = std::min<size_t>((de->s.P + 4095) / 4096, 4096) * k * 16;   // one (key, visit) pair per list and rank
  if (need <= de->topk_capacity) return NLS_OK;
  void *p = nullptr;
  int rc = de->mem.alloc(&p, need);
  if (rc != NLS_OK) return rc;
  de->s.topk_scratch = p;
  de->topk_capacity = need;
  return NLS_OK;
}

int nls_de_export_top(nls_de *de, uint64_t k, void *rows_dev, void *scores_dev) {
  if (!de || !rows_dev || !scores_dev) return fail(NLS_ERR_INVALID, "nls_de_export_top: NULL argument");
  if (k < 1 || k > de->s.P) return fail(NLS_ERR_INVALID, "nls_de_export_top: k out of range");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  if (int rc = de_ensure_topk_scratch(de, k)) return rc;
  NLS_CUDA(de->ops->migrate(de->s, +1, k, rows_dev, scores_dev, de->g, de->ctx->stream));
  return NLS_OK;
}
int nls_de_import_migrants(nls_de *de, uint64_t k, const void *rows_dev, const void *scores_dev) {
  if (!de || !rows_dev || !scores_dev) return fail(NLS_ERR_INVALID, "nls_de_import_migrants: NULL argument");
  if (k < 1 || k > de->s.P) return fail(NLS_ERR_INVALID, "nls_de_import_migrants: k out of range");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  if (int rc = de_ensure_topk_scratch(de, k)) return rc;
  NLS_CUDA(de->ops->migrate(de->s, -1, k, const_cast<void *>(rows_dev), const_cast<void *>(scores_dev), de->g,
                            de->ctx->stream));
  return NLS_OK;
}

// Tiny problems (the reference's own shapes: pop_size 50, 2-D): the whole solve in one launch of one block with the
// population in shared memory and the result written to mapped host memory (de_tiny.cuh).  NLS_DE_TINY=0 in the
// environment sends them through the general kernels instead (same results).
static bool de_tiny_eligible(const nls_de_cfg *c) {
  const char *e = std::getenv("NLS_DE_TINY");
  if (e && e[0] == '0') return false;
  return c->pop_size <= kTinyMaxPop && c->dim <= kTinyMaxDim && c->objective >= 0 && c->objective < NLS_OBJECTIVE_COUNT &&
         c->flags == 0 && !guard_mode();
}
static int de_solve_tiny(nls_ctx *ctx, const nls_de_cfg *cfg, const void *x0_host, void *x_best_host, nls_status *status) {
  NLS_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->tiny_host) {
    NLS_CUDA(cudaHostAlloc(&ctx->tiny_host, sizeof(DETinyResult), cudaHostAllocMapped));
    NLS_CUDA(cudaHostGetDevicePointer(&ctx->tiny_dev, ctx->tiny_host, 0));
  }
  DETinyArgs a;
  std::memset(&a, 0, sizeof(a));
  a.P = unsigned(cfg->pop_size); a.d = unsigned(cfg->dim);
  a.seed = cfg->seed; a.offset = cfg->agent_offset;
  a.max_iter = cfg->max_iter; a.vnc_limit = cfg->best_val_no_change;
  a.strategy = cfg->strategy;
  a.F = cfg->differential_weight; a.fm = cfg->minimize ? 1.0 : -1.0; a.eps = cfg->eps;
  u64 cr_le; int cr_none;
  if (cfg->dtype == NLS_F64) crossover_threshold<double>(cfg->crossover_prob, &cr_le, &cr_none);
  else crossover_threshold<float>(cfg->crossover_prob, &cr_le, &cr_none);
  a.cr_le = cr_le; a.cr_none = cr_none;
  for (u64 j = 0; j < cfg->dim; j++)
    a.x0[j] = cfg->dtype == NLS_F64 ? static_cast<const double *>(x0_host)[j]
                                    : static_cast<double>(static_cast<const float *>(x0_host)[j]);
  a.result = ctx->tiny_dev;
  NLS_CUDA(cfg->dtype == NLS_F64 ? de_tiny_launch_f64(cfg->objective, a, ctx->stream)
                                 : de_tiny_launch_f32(cfg->objective, a, ctx->stream));
  NLS_CUDA(cudaStreamSynchronize(ctx->stream));
  const DETinyResult *r = static_cast<const DETinyResult *>(ctx->tiny_host);
  for (u64 j = 0; j < cfg->dim; j++) {
    if (cfg->dtype == NLS_F64) static_cast<double *>(x_best_host)[j] = r->x[j];
    else static_cast<float *>(x_best_host)[j] = static_cast<float>(r->x[j]);
  }
  if (status) {
    std::memset(status, 0, sizeof(*status));
    status->f_value = r->f_value;
    status->iterations = r->iterations;
    status->function_calls = cfg->pop_size * (r->iterations + 1);
    status->best_index = r->best_id;
    status->val_no_change = r->vnc;
    status->stopped = 1;
    status->stop_reason = r->stop_reason;
    status->best_valid = 1;
    status->std_err = r->std_err;
    status->repair_reruns = r->reruns;
    status->repair_rounds = r->rounds;
    status->accepted_total = r->accepted;
  }
  return NLS_OK;
}

int nls_de_solve(nls_ctx *ctx, const nls_de_cfg *cfg, const void *x0_host, void *x_best_host, nls_status *status) {
  if (!x_best_host) return fail(NLS_ERR_INVALID, "nls_de_solve: x_best_host is NULL");
  if (ctx && cfg && x0_host && de_tiny_eligible(cfg)) {
    int rc = de_validate(cfg);
    if (rc != NLS_OK) return rc;
    return de_solve_tiny(ctx, cfg, x0_host, x_best_host, status);
  }
  nls_de *de = nullptr;
  int rc = nls_de_create(ctx, cfg, x0_host, &de);
  if (rc != NLS_OK) return rc;
  nls_status st;
  rc = nls_de_sync(de, &st);
  // generations past the stop rule are no-ops on the device, so batches only bound the host's polling interval; the
  // one-launch path (small populations) leaves its generation loop when a stop rule fires, so it gets the whole budget
  uint64_t batch = (de && de->s.P * de->s.d <= kPersistMaxElems) ? (1ull << 16) : 4;
  while (rc == NLS_OK && !st.stopped) {
    const uint64_t left = cfg->max_iter > st.iterations ? cfg->max_iter - st.iterations : 1;
    rc = nls_de_step(de, std::min<uint64_t>(batch, left));
    if (rc == NLS_OK) rc = nls_de_sync(de, &st);
    if (batch < 64) batch *= 2;
  }
  if (rc == NLS_OK) rc = nls_de_read_best(de, x_best_host);
  if (rc == NLS_OK && status) *status = st;
  nls_de_destroy(de);
  return rc;
}

/* ================================================================ PSO ======================================== */

static int pso_validate(const nls_pso_cfg *c) {
  if (c->dtype != NLS_F32 && c->dtype != NLS_F64) return fail(NLS_ERR_INVALID, "PSO: unknown dtype %d", c->dtype);
  if (!objective_known(c->objective)) return fail(NLS_ERR_INVALID, "PSO: unknown objective %d", c->objective);
  if (const unsigned fd = objective_fixed_dim(c->objective))
    if (fd != c->dim)
      return fail(NLS_ERR_INVALID, "PSO: objective %d is a closed form of dimension %u, dim is %llu", c->objective, fd,
                  static_cast<unsigned long long>(c->dim));
  if (c->pso_type != NLS_PSO_VANILLA && c->pso_type != NLS_PSO_ACCELERATED) return fail(NLS_ERR_INVALID, "PSO: unknown type %d", c->pso_type);
  if (c->n_particles < 1 || c->dim < 1) return fail(NLS_ERR_INVALID, "PSO: n_particles and dim must be >= 1");
  if (c->n_particles >= 0xffffffffull || c->dim >= 0xffffffffull) return fail(NLS_ERR_INVALID, "PSO: n_particles (per shard) and dim must fit 32 bits");
  const u64 pg = c->n_particles_global ? c->n_particles_global : c->n_particles;
  if (c->particle_offset + c->n_particles > pg) return fail(NLS_ERR_INVALID, "PSO: shard exceeds the global swarm");
  if (c->pso_type == NLS_PSO_VANILLA && !(c->flags & NLS_FLAG_SOCIAL_INDEX_J) && pg > c->dim)
    return fail(NLS_ERR_INVALID,
                "vanilla PSO with n_particles > dim reads swarm_best_position out of bounds in the reference "
                "(nlsolver.h:2674); pass NLS_FLAG_SOCIAL_INDEX_J for the corrected social term");
  return NLS_OK;
}

int nls_pso_destroy(nls_pso *pso) {
  if (!pso) return NLS_OK;
  cudaSetDevice(pso->ctx->device);
  cudaStreamSynchronize(pso->ctx->stream);
  pso->graph.reset();
  pso->mem.release();
  delete pso;
  return NLS_OK;
}

// inertia the reference would hold while moving in iteration `iter` (nlsolver.h:2613; vanilla keeps the ctor value)
static double pso_inertia(const nls_pso *p, u64 iter) {
  if (p->cfg.pso_type == NLS_PSO_VANILLA) return p->cfg.inertia;
  if (p->cfg.dtype == NLS_F64) return std::pow(p->cfg.inertia, static_cast<double>(iter));
  return static_cast<double>(static_cast<float>(std::pow(static_cast<double>(static_cast<float>(p->cfg.inertia)),
                                                         static_cast<double>(iter))));
}

static int pso_build(nls_ctx *ctx, const nls_pso_cfg *cfg, const void *lower, const void *upper, nls_pso *p) {
  const u64 P = cfg->n_particles, d = cfg->dim;
  p->ctx = ctx;
  p->mem.ctx = ctx;
  p->cfg = *cfg;
  p->elem = elem_size(cfg->dtype);
  p->ops = cfg->dtype == NLS_F64 ? pso_ops_f64() : pso_ops_f32();
  if (const ObjectivePlugin *pl = plugin_for(cfg->objective)) p->ops = cfg->dtype == NLS_F64 ? pl->pso_f64 : pl->pso_f32;
  p->g = make_geom(ctx, P);
  p->enqueued = 0;
  p->xchg = nullptr;
  PSOState &s = p->s;
  std::memset(&s, 0, sizeof(s));
  s.P = P; s.d = d; s.stride = round_up(d, 32 / p->elem);
  s.P_global = cfg->n_particles_global ? cfg->n_particles_global : P;
  s.seed = cfg->seed; s.offset = cfg->particle_offset;
  s.pso_type = cfg->pso_type; s.constrained = cfg->constrained;
  s.objective = plugin_for(cfg->objective) ? 100 /* OBJ_CUSTOM */ : cfg->objective;
  s.social_j = (cfg->flags & NLS_FLAG_SOCIAL_INDEX_J) ? 1 : 0;
  s.init_inertia = cfg->inertia; s.cog = cfg->cognitive_coef; s.soc = cfg->social_coef;
  s.fm = cfg->minimize ? 1.0 : -1.0; s.eps = cfg->eps;
  s.max_iter = cfg->max_iter; s.vnc_limit = cfg->best_val_no_change;
  p->record_bytes = nls_record_bytes(cfg->dtype, d);
  const size_t row_bytes = size_t(P) * s.stride * p->elem;
  int rc;
#define NLS_ALLOC(ptr, bytes) if ((rc = p->mem.alloc(reinterpret_cast<void **>(&(ptr)), (bytes))) != NLS_OK) return rc
  NLS_ALLOC(s.pos, row_bytes);
  if (cfg->pso_type == NLS_PSO_VANILLA) NLS_ALLOC(s.vel, row_bytes);
  NLS_ALLOC(s.pbest, P * p->elem);
  NLS_ALLOC(s.last, P * p->elem);
  NLS_ALLOC(s.sbest, s.stride * p->elem);
  NLS_ALLOC(s.lower, s.stride * p->elem);   // padded like a row: the move kernel reads bounds with 128-bit loads
  NLS_ALLOC(s.upper, s.stride * p->elem);
  NLS_ALLOC(s.ctrl, sizeof(PSOCtrl));
  const size_t n_part = std::max(p->g.reduce_blocks, 16);   // (the one-launch path reduces with up to 16 blocks)
  NLS_ALLOC(s.part_min, n_part * sizeof(double));
  NLS_ALLOC(s.part_idx, n_part * sizeof(unsigned long long));
  NLS_ALLOC(s.part_mom, n_part * sizeof(Moments));
  NLS_ALLOC(p->record, p->record_bytes);
  std::vector<double> inertia_table;
  if (cfg->pso_type == NLS_PSO_ACCELERATED) {
    // one entry per iteration, host libm like the reference; 2^20 entries (8 MB) cover any run that is not endless —
    // 0.8^k underflows to zero after ~3400 iterations — and only beyond the table does the device evaluate pow itself
    const u64 n = std::min<u64>(cfg->max_iter == ~0ull ? 16384 : cfg->max_iter + 1, cfg->max_iter > (1ull << 20) ? 16384 : (1ull << 20));
    inertia_table.resize(n);
    for (u64 k = 0; k < n; k++) inertia_table[k] = pso_inertia(p, k);
    double *tab = nullptr;
    NLS_ALLOC(tab, n * sizeof(double));
    s.inertia_table = tab;
    s.inertia_n = n;
  }
#undef NLS_ALLOC
  cudaStream_t st = ctx->stream;
  PSOCtrl c0;
  std::memset(&c0, 0, sizeof(c0));
  c0.best_value = 100000.0;                          // swarm_best_value, nlsolver.h:2631
  NLS_CUDA(cudaMemcpyAsync(s.ctrl, &c0, sizeof(c0), cudaMemcpyHostToDevice, st));
  NLS_CUDA(cudaMemsetAsync(s.sbest, 0, s.stride * p->elem, st));
  NLS_CUDA(cudaMemsetAsync(p->record, 0, p->record_bytes, st));
  NLS_CUDA(cudaMemsetAsync(s.lower, 0, s.stride * p->elem, st));
  NLS_CUDA(cudaMemsetAsync(s.upper, 0, s.stride * p->elem, st));
  if (!inertia_table.empty())
    NLS_CUDA(cudaMemcpyAsync(const_cast<double *>(s.inertia_table), inertia_table.data(),
                             inertia_table.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  NLS_CUDA(cudaMemcpyAsync(s.lower, lower, d * p->elem, cudaMemcpyHostToDevice, st));
  NLS_CUDA(cudaMemcpyAsync(s.upper, upper, d * p->elem, cudaMemcpyHostToDevice, st));
  NLS_CUDA(cudaStreamSynchronize(st));               // c0 lives on this stack frame
  NLS_CUDA(p->ops->init(s, p->g, st));
  NLS_CUDA(p->ops->candidate(s, p->record, p->g, st));
  p->first_apply_pending = true;
  if (s.P_global == P) {                              // single-GPU swarm: finish update_best_positions here
    NLS_CUDA(p->ops->apply(s, p->record, 1, p->record_bytes, 1, st));
    p->first_apply_pending = false;
  }
  NLS_CUDA(cudaStreamSynchronize(st));
  return NLS_OK;
}

int nls_pso_create(nls_ctx *ctx, const nls_pso_cfg *cfg, const void *lower_host, const void *upper_host,
                   nls_pso **out) {
  if (!ctx || !cfg || !lower_host || !upper_host || !out) return fail(NLS_ERR_INVALID, "nls_pso_create: NULL argument");
  *out = nullptr;
  int rc = pso_validate(cfg);
  if (rc != NLS_OK) return rc;
  NLS_CUDA(cudaSetDevice(ctx->device));
  nls_pso *p = new nls_pso();
  rc = pso_build(ctx, cfg, lower_host, upper_host, p);
  if (rc != NLS_OK) { p->mem.release(); delete p; return rc; }
  *out = p;
  return NLS_OK;
}

int nls_pso_step_local(nls_pso *p, void *record_dev) {
  if (!p) return fail(NLS_ERR_INVALID, "nls_pso_step_local: NULL handle");
  if (p->first_apply_pending) return fail(NLS_ERR_STATE, "sharded swarm: apply the initial candidates before stepping");
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  NLS_CUDA(p->ops->move(p->s, p->g, st));
  NLS_CUDA(p->ops->candidate(p->s, p->record, p->g, st));
  p->enqueued++;
  if (record_dev) NLS_CUDA(cudaMemcpyAsync(record_dev, p->record, p->record_bytes, cudaMemcpyDeviceToDevice, st));
  return NLS_OK;
}

int nls_pso_export_candidate(nls_pso *p, void *record_dev) {
  if (!p || !record_dev) return fail(NLS_ERR_INVALID, "nls_pso_export_candidate: NULL argument");
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  NLS_CUDA(cudaMemcpyAsync(record_dev, p->record, p->record_bytes, cudaMemcpyDeviceToDevice, p->ctx->stream));
  return NLS_OK;
}

int nls_pso_apply_candidates(nls_pso *p, const void *records_dev, uint64_t n_records) {
  if (!p || !records_dev || n_records < 1) return fail(NLS_ERR_INVALID, "nls_pso_apply_candidates: bad argument");
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  NLS_CUDA(p->ops->apply(p->s, records_dev, n_records, p->record_bytes, p->first_apply_pending ? 1 : 0, p->ctx->stream));
  p->first_apply_pending = false;
  return NLS_OK;
}

int nls_pso_step(nls_pso *p, uint64_t n_generations) {
  if (!p) return fail(NLS_ERR_INVALID, "nls_pso_step: NULL handle");
  if (p->s.P_global != p->s.P) return fail(NLS_ERR_STATE, "nls_pso_step is for a single-GPU swarm; a shard uses step_local / apply_candidates");
  uint64_t left = n_generations;
  if (left > 0 && p->s.P * p->s.d <= kPersistMaxElems && !p->first_apply_pending && p->ops->persistent &&
      !p->persistent_failed && de_one_launch_enabled(p->s.P)) {
    NLS_CUDA(cudaSetDevice(p->ctx->device));
    if (p->ops->persistent(p->s, p->record, p->record_bytes, left, p->ctx->stream) == cudaSuccess) {
      p->enqueued += left;
      return NLS_OK;
    }
    cudaGetLastError();
    p->persistent_failed = true;
  }
  if (p->s.P * p->s.d <= kGraphMaxElems && !p->first_apply_pending) {
    NLS_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t st = p->ctx->stream;
    while (left >= kGraphGens) {
      const bool ok = graph_replay(p->graph, st, [&] {
        for (int g = 0; g < kGraphGens; g++) {
          if (p->ops->move(p->s, p->g, st) != cudaSuccess) return false;
          if (p->ops->candidate_apply(p->s, p->record, p->record_bytes, p->g, st) != cudaSuccess) return false;
        }
        return true;
      });
      if (!ok) break;
      left -= kGraphGens;
      p->enqueued += kGraphGens;
    }
  }
  if (left > 0 && p->first_apply_pending) return fail(NLS_ERR_STATE, "nls_pso_step: apply the initial candidates first");
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  for (uint64_t g = 0; g < left; g++) {
    NLS_CUDA(p->ops->move(p->s, p->g, p->ctx->stream));
    NLS_CUDA(p->ops->candidate_apply(p->s, p->record, p->record_bytes, p->g, p->ctx->stream));
    p->enqueued++;
  }
  return NLS_OK;
}

int nls_pso_sync(nls_pso *p, nls_status *status) {
  if (!p) return fail(NLS_ERR_INVALID, "nls_pso_sync: NULL handle");
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  PSOCtrl c;
  NLS_CUDA(cudaMemcpyAsync(&c, p->s.ctrl, sizeof(c), cudaMemcpyDeviceToHost, p->ctx->stream));
  NLS_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (status) {
    std::memset(status, 0, sizeof(*status));
    status->f_value = c.best_value;
    status->iterations = c.iter;
    status->function_calls = p->s.P_global * (c.iter + 1);
    status->best_index = c.best_index;
    status->val_no_change = c.vnc;
    status->stopped = c.stop;
    status->stop_reason = c.stop_reason;
    status->best_valid = c.best_valid;
    status->std_err = c.std_err;
  }
  return NLS_OK;
}

static int pso_read(nls_pso *p, const void *dev, size_t bytes, void *host, const char *what) {
  if (!p || !host) return fail(NLS_ERR_INVALID, "%s: NULL argument", what);
  if (!dev) return fail(NLS_ERR_STATE, "%s: not available for this swarm type", what);
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  NLS_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, p->ctx->stream));
  NLS_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return NLS_OK;
}
static int pso_read_rows(nls_pso *p, const void *dev, void *host, const char *what) {
  if (!p || !host) return fail(NLS_ERR_INVALID, "%s: NULL argument", what);
  if (!dev) return fail(NLS_ERR_STATE, "%s: not available for this swarm type", what);
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  NLS_CUDA(cudaMemcpy2DAsync(host, p->s.d * p->elem, dev, p->s.stride * p->elem, p->s.d * p->elem, p->s.P,
                             cudaMemcpyDeviceToHost, p->ctx->stream));
  NLS_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return NLS_OK;
}
int nls_pso_read_best(nls_pso *p, void *x_host) { return pso_read(p, p ? p->s.sbest : nullptr, p ? p->s.d * p->elem : 0, x_host, "nls_pso_read_best"); }
int nls_pso_read_positions(nls_pso *p, void *rows_host) { return pso_read_rows(p, p ? p->s.pos : nullptr, rows_host, "nls_pso_read_positions"); }
int nls_pso_read_velocities(nls_pso *p, void *rows_host) { return pso_read_rows(p, p ? p->s.vel : nullptr, rows_host, "nls_pso_read_velocities"); }
int nls_pso_read_pbest_values(nls_pso *p, void *vals_host) { return pso_read(p, p ? p->s.pbest : nullptr, p ? p->s.P * p->elem : 0, vals_host, "nls_pso_read_pbest_values"); }
int nls_pso_read_last_values(nls_pso *p, void *vals_host) { return pso_read(p, p ? p->s.last : nullptr, p ? p->s.P * p->elem : 0, vals_host, "nls_pso_read_last_values"); }

int nls_pso_solve(nls_ctx *ctx, const nls_pso_cfg *cfg, const void *lower_host, const void *upper_host,
                  void *x_best_host, nls_status *status) {
  if (!x_best_host) return fail(NLS_ERR_INVALID, "nls_pso_solve: x_best_host is NULL");
  nls_pso *p = nullptr;
  int rc = nls_pso_create(ctx, cfg, lower_host, upper_host, &p);
  if (rc != NLS_OK) return rc;
  nls_status st;
  rc = nls_pso_sync(p, &st);
  uint64_t batch = (p && p->s.P * p->s.d <= kPersistMaxElems) ? (1ull << 16) : 4;   // one launch runs to the stop rule
  while (rc == NLS_OK && !st.stopped) {
    const uint64_t left = cfg->max_iter > st.iterations ? cfg->max_iter - st.iterations : 1;
    rc = nls_pso_step(p, std::min<uint64_t>(batch, left));
    if (rc == NLS_OK) rc = nls_pso_sync(p, &st);
    if (batch < 64) batch *= 2;
  }
  if (rc == NLS_OK && st.best_valid) rc = nls_pso_read_best(p, x_best_host);
  if (rc == NLS_OK && status) *status = st;
  nls_pso_destroy(p);
  return rc;
}

/* ================================================================ SANN chains ================================ */

struct nls_sann {
  nls_ctx *ctx;
  nls_sann_cfg cfg;
  SANNState s;
  LaunchGeom g;
  const SANNOps *ops;
  DeviceBuffers mem;
  size_t elem;
  u64 steps_done;
  void *dense;      // [C][d] staging for the per-chain read-out, allocated on first use
};

static int sann_validate(const nls_sann_cfg *c, uint64_t x0_count) {
  if (c->dtype != NLS_F32 && c->dtype != NLS_F64) return fail(NLS_ERR_INVALID, "SANN: unknown dtype %d", c->dtype);
  if (!objective_known(c->objective)) return fail(NLS_ERR_INVALID, "SANN: unknown objective %d", c->objective);
  if (const unsigned fd = objective_fixed_dim(c->objective))
    if (fd != c->dim)
      return fail(NLS_ERR_INVALID, "SANN: objective %d is a closed form of dimension %u, dim is %llu", c->objective, fd,
                  static_cast<unsigned long long>(c->dim));
  if (c->n_chains < 1 || c->dim < 1) return fail(NLS_ERR_INVALID, "SANN: n_chains and dim must be >= 1");
  if (c->n_chains >= 0xffffffffull || c->dim >= 0xffffffffull) return fail(NLS_ERR_INVALID, "SANN: n_chains and dim must fit 32 bits");
  if (x0_count != 1 && x0_count != c->n_chains) return fail(NLS_ERR_INVALID, "SANN: x0_count must be 1 or n_chains");
  if (!(c->temperature_max > 0.0)) return fail(NLS_ERR_INVALID, "SANN: temperature_max must be > 0");
  if (c->temperature_iter > 1 && c->max_iter > (1ull << 62) / (c->temperature_iter - 1))
    return fail(NLS_ERR_INVALID, "SANN: max_iter * (temperature_iter - 1) overflows");
  return NLS_OK;
}

int nls_sann_destroy(nls_sann *sa) {
  if (!sa) return NLS_OK;
  cudaSetDevice(sa->ctx->device);
  cudaStreamSynchronize(sa->ctx->stream);
  sa->mem.release();
  delete sa;
  return NLS_OK;
}

// temperature of outer iteration k exactly as the reference evaluates it (nlsolver.h:2792-2793), in scalar_t
static double sann_temperature(int dtype, double tmax, u64 iter) {
  if (dtype == NLS_F64) return tmax / std::log(static_cast<double>(iter) + 1.7182818);
  const float e_minus_1 = static_cast<float>(1.7182818);
  return static_cast<double>(static_cast<float>(tmax) / std::log(static_cast<float>(iter) + e_minus_1));
}

static int sann_build(nls_ctx *ctx, const nls_sann_cfg *cfg, const void *x0_host, u64 x0_count, nls_sann *sa) {
  const u64 C = cfg->n_chains, d = cfg->dim;
  sa->ctx = ctx;
  sa->mem.ctx = ctx;
  sa->cfg = *cfg;
  sa->elem = elem_size(cfg->dtype);
  sa->steps_done = 0;
  sa->ops = cfg->dtype == NLS_F64 ? sann_ops_f64() : sann_ops_f32();
  if (const ObjectivePlugin *pl = plugin_for(cfg->objective)) sa->ops = cfg->dtype == NLS_F64 ? pl->sann_f64 : pl->sann_f32;
  sa->g = make_geom(ctx, C);
  SANNState &s = sa->s;
  std::memset(&s, 0, sizeof(s));
  s.C = C; s.d = d; s.stride = round_up(d, 32 / sa->elem);
  s.seed = cfg->seed; s.offset = cfg->chain_offset;
  s.inner = cfg->temperature_iter > 1 ? cfg->temperature_iter - 1 : 0;
  s.total_steps = cfg->max_iter * s.inner;
  s.tmax = cfg->temperature_max;
  // scale = 1.0 / temperature_max, a double division rounded to scalar_t (nlsolver.h:2782)
  s.scale = cfg->dtype == NLS_F64 ? 1.0 / cfg->temperature_max
                                  : static_cast<double>(static_cast<float>(1.0 / static_cast<double>(static_cast<float>(cfg->temperature_max))));
  s.fm = cfg->minimize ? 1.0 : -1.0;
  s.objective = plugin_for(cfg->objective) ? 100 /* OBJ_CUSTOM */ : cfg->objective;
  const size_t row_bytes = size_t(C) * s.stride * sa->elem;
  int rc;
#define NLS_ALLOC(ptr, bytes) if ((rc = sa->mem.alloc(reinterpret_cast<void **>(&(ptr)), (bytes))) != NLS_OK) return rc
  for (int k = 0; k < 3; k++) NLS_ALLOC(s.buf[k], row_bytes);
  NLS_ALLOC(s.role, C);
  NLS_ALLOC(s.best, C * sa->elem);
  NLS_ALLOC(s.n_acc, C * sizeof(uint32_t));
  NLS_ALLOC(s.n_imp, C * sizeof(uint32_t));
  NLS_ALLOC(s.ctrl, sizeof(SANNCtrl));
  sa->dense = nullptr;
  const u64 tn = std::min<u64>(std::max<u64>(cfg->max_iter, 1), 1u << 20);   // host libm values; device log only beyond
  std::vector<double> table(tn);
  for (u64 k = 0; k < tn; k++)
    table[k] = sann_temperature(cfg->dtype, cfg->temperature_max, k);
  double *tab = nullptr;
  NLS_ALLOC(tab, tn * sizeof(double));
  s.t_table = tab; s.t_n = tn;
  void *x0_dev = nullptr;
  NLS_ALLOC(x0_dev, size_t(x0_count) * d * sa->elem);
#undef NLS_ALLOC
  cudaStream_t st = ctx->stream;
  NLS_CUDA(cudaMemcpyAsync(tab, table.data(), tn * sizeof(double), cudaMemcpyHostToDevice, st));
  NLS_CUDA(cudaMemcpyAsync(x0_dev, x0_host, size_t(x0_count) * d * sa->elem, cudaMemcpyHostToDevice, st));
  NLS_CUDA(cudaMemsetAsync(s.ctrl, 0, sizeof(SANNCtrl), st));
  NLS_CUDA(sa->ops->init(s, x0_dev, x0_count, sa->g, st));
  NLS_CUDA(cudaStreamSynchronize(st));               // table / x0 are host temporaries
  return NLS_OK;
}

int nls_sann_create(nls_ctx *ctx, const nls_sann_cfg *cfg, const void *x0_host, uint64_t x0_count, nls_sann **out) {
  if (!ctx || !cfg || !x0_host || !out) return fail(NLS_ERR_INVALID, "nls_sann_create: NULL argument");
  *out = nullptr;
  int rc = sann_validate(cfg, x0_count);
  if (rc != NLS_OK) return rc;
  NLS_CUDA(cudaSetDevice(ctx->device));
  nls_sann *sa = new nls_sann();
  rc = sann_build(ctx, cfg, x0_host, x0_count, sa);
  if (rc != NLS_OK) { sa->mem.release(); delete sa; return rc; }
  *out = sa;
  return NLS_OK;
}

int nls_sann_step(nls_sann *sa, uint64_t n_candidates) {
  if (!sa) return fail(NLS_ERR_INVALID, "nls_sann_step: NULL handle");
  NLS_CUDA(cudaSetDevice(sa->ctx->device));
  const u64 left = sa->s.total_steps - sa->steps_done;
  u64 n = std::min<u64>(n_candidates, left);
  // one launch carries every chain through its candidates; bound a launch to ~2^31 coordinate updates (a few tens of
  // milliseconds of device time) so that large batches stay responsive to sync / destroy
  const u64 per_step = std::max<u64>(sa->s.C * sa->s.d, 1);
  const u64 chunk = std::min<u64>(std::max<u64>((1ull << 31) / per_step, 1), 1ull << 20);   // tiny batches: <= ~1 s
  while (n > 0) {
    const u64 k = std::min(n, chunk);
    NLS_CUDA(sa->ops->steps(sa->s, sa->steps_done, k, sa->g, sa->ctx->stream));
    sa->steps_done += k;
    n -= k;
  }
  return NLS_OK;
}

int nls_sann_sync(nls_sann *sa, nls_status *status) {
  if (!sa) return fail(NLS_ERR_INVALID, "nls_sann_sync: NULL handle");
  NLS_CUDA(cudaSetDevice(sa->ctx->device));
  cudaStream_t st = sa->ctx->stream;
  SANNCtrl c;
  if (status) {
    NLS_CUDA(sa->ops->best(sa->s, st));
    NLS_CUDA(cudaMemcpyAsync(&c, sa->s.ctrl, sizeof(c), cudaMemcpyDeviceToHost, st));
  }
  NLS_CUDA(cudaStreamSynchronize(st));
  if (status) {
    std::memset(status, 0, sizeof(*status));
    const bool done = sa->steps_done >= sa->s.total_steps;
    status->f_value = c.best_value;
    status->iterations = done ? sa->cfg.max_iter : sa->steps_done / sa->s.inner;
    status->function_calls = sa->s.C * (1 + sa->steps_done);
    status->best_index = sa->s.offset + c.best_chain;
    status->stopped = done ? 1 : 0;
    status->stop_reason = done ? 1 : 0;
    status->best_valid = c.best_valid;
  }
  return NLS_OK;
}

static int sann_read_rows(nls_sann *sa, int which, u64 first, u64 count, void *host) {
  cudaStream_t st = sa->ctx->stream;
  if (!sa->dense) {
    int rc = sa->mem.alloc(&sa->dense, size_t(sa->s.C) * sa->s.d * sa->elem);
    if (rc != NLS_OK) return rc;
  }
  NLS_CUDA(sa->ops->gather(sa->s, which, sa->dense, sa->g, st));
  const size_t row = sa->s.d * sa->elem;
  NLS_CUDA(cudaMemcpyAsync(host, static_cast<const char *>(sa->dense) + first * row, count * row, cudaMemcpyDeviceToHost, st));
  NLS_CUDA(cudaStreamSynchronize(st));
  return NLS_OK;
}

int nls_sann_read_best(nls_sann *sa, void *x_host) {
  if (!sa || !x_host) return fail(NLS_ERR_INVALID, "nls_sann_read_best: NULL argument");
  NLS_CUDA(cudaSetDevice(sa->ctx->device));
  cudaStream_t st = sa->ctx->stream;
  SANNCtrl c;
  NLS_CUDA(sa->ops->best(sa->s, st));
  NLS_CUDA(cudaMemcpyAsync(&c, sa->s.ctrl, sizeof(c), cudaMemcpyDeviceToHost, st));
  NLS_CUDA(cudaStreamSynchronize(st));
  // the best chain's x is one contiguous row of the buffer its role byte names: no gather of the whole batch
  const char *row = static_cast<const char *>(sa->s.buf[c.best_buf % 3]) + c.best_chain * sa->s.stride * sa->elem;
  NLS_CUDA(cudaMemcpyAsync(x_host, row, sa->s.d * sa->elem, cudaMemcpyDeviceToHost, st));
  NLS_CUDA(cudaStreamSynchronize(st));
  return NLS_OK;
}

int nls_sann_read_chains(nls_sann *sa, void *x_best_host, void *f_best_host, void *p_cur_host, uint32_t *n_accepted,
                         uint32_t *n_improved) {
  if (!sa) return fail(NLS_ERR_INVALID, "nls_sann_read_chains: NULL handle");
  NLS_CUDA(cudaSetDevice(sa->ctx->device));
  cudaStream_t st = sa->ctx->stream;
  const u64 C = sa->s.C;
  int rc;
  if (x_best_host && (rc = sann_read_rows(sa, 0, 0, C, x_best_host)) != NLS_OK) return rc;
  if (p_cur_host && (rc = sann_read_rows(sa, 1, 0, C, p_cur_host)) != NLS_OK) return rc;
  if (f_best_host) NLS_CUDA(cudaMemcpyAsync(f_best_host, sa->s.best, C * sa->elem, cudaMemcpyDeviceToHost, st));
  if (n_accepted) NLS_CUDA(cudaMemcpyAsync(n_accepted, sa->s.n_acc, C * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (n_improved) NLS_CUDA(cudaMemcpyAsync(n_improved, sa->s.n_imp, C * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  NLS_CUDA(cudaStreamSynchronize(st));
  return NLS_OK;
}

int nls_sann_solve(nls_ctx *ctx, const nls_sann_cfg *cfg, const void *x0_host, uint64_t x0_count, void *x_best_host,
                   nls_status *status) {
  if (!x_best_host) return fail(NLS_ERR_INVALID, "nls_sann_solve: x_best_host is NULL");
  nls_sann *sa = nullptr;
  int rc = nls_sann_create(ctx, cfg, x0_host, x0_count, &sa);
  if (rc != NLS_OK) return rc;
  rc = nls_sann_step(sa, ~0ull);
  nls_status st;
  if (rc == NLS_OK) rc = nls_sann_sync(sa, &st);
  if (rc == NLS_OK) rc = nls_sann_read_best(sa, x_best_host);
  if (rc == NLS_OK && status) *status = st;
  nls_sann_destroy(sa);
  return rc;
}

/* ================================================================ NelderMeadPSO batches ======================= */

int nls_nmpso_solve(nls_ctx *ctx, const nls_nmpso_cfg *cfg, const void *x0_host, uint64_t x0_count, void *x_best_host,
                    void *f_best_host, uint64_t *iterations_host, uint64_t *function_calls_host, nls_status *status) {
  if (!ctx || !cfg || !x0_host) return fail(NLS_ERR_INVALID, "nls_nmpso_solve: NULL argument");
  if (cfg->dtype != NLS_F32 && cfg->dtype != NLS_F64) return fail(NLS_ERR_INVALID, "NelderMeadPSO: unknown dtype %d", cfg->dtype);
  if (cfg->objective < 0 || cfg->objective >= NLS_OBJECTIVE_COUNT)
    return fail(NLS_ERR_INVALID, "NelderMeadPSO: objective %d is not a built-in objective", cfg->objective);
  if (const unsigned fd = objective_fixed_dim(cfg->objective))
    if (fd != cfg->dim) return fail(NLS_ERR_INVALID, "NelderMeadPSO: objective %d is a closed form of dimension %u", cfg->objective, fd);
  // dim < 2: the reference prints a notice and returns solver_status(999999, 0, 0) (nlsolver.h:3619-3629)
  if (cfg->dim < 2 || cfg->dim > kNMPSOMaxDim) return fail(NLS_ERR_INVALID, "NelderMeadPSO: 2 <= dim <= %u", kNMPSOMaxDim);
  if (cfg->n_solvers < 1 || cfg->n_solvers >= 0xffffffffull) return fail(NLS_ERR_INVALID, "NelderMeadPSO: n_solvers out of range");
  if (x0_count != 1 && x0_count != cfg->n_solvers) return fail(NLS_ERR_INVALID, "NelderMeadPSO: x0_count must be 1 or n_solvers");
  NLS_CUDA(cudaSetDevice(ctx->device));
  const size_t elem = elem_size(cfg->dtype);
  const u64 C = cfg->n_solvers, d = cfg->dim, N = 3 * d + 1;
  DeviceBuffers mem;
  mem.ctx = ctx;
  NMPSOState s;
  std::memset(&s, 0, sizeof(s));
  s.C = C; s.d = d; s.stride = round_up(d, 32 / elem); s.x0_count = x0_count;
  s.seed = cfg->seed; s.offset = cfg->solver_offset; s.max_iter = cfg->max_iter; s.no_change_limit = cfg->no_change_best_iter;
  s.alpha = cfg->alpha; s.gamma = cfg->gamma; s.rho = cfg->rho; s.sigma = cfg->sigma; s.inertia = cfg->inertia;
  s.cog = cfg->cognitive_coef; s.soc = cfg->social_coef; s.eps = cfg->eps; s.fm = cfg->minimize ? 1.0 : -1.0;
  s.objective = cfg->objective;
  int rc = NLS_OK;
  auto alloc = [&](void **p, size_t bytes) { if (rc == NLS_OK) rc = mem.alloc(p, bytes); };
  alloc(&s.pos, size_t(C) * N * s.stride * elem);
  alloc(&s.vel, size_t(C) * N * s.stride * elem);
  alloc(&s.tmp, size_t(C) * 4 * s.stride * elem);
  alloc(&s.x0, size_t(x0_count) * d * elem);
  alloc(&s.x_best, size_t(C) * d * elem);
  alloc(&s.f_best, size_t(C) * elem);
  alloc(reinterpret_cast<void **>(&s.iters), size_t(C) * sizeof(unsigned long long));
  alloc(reinterpret_cast<void **>(&s.evals), size_t(C) * sizeof(unsigned long long));
  cudaStream_t st = ctx->stream;
  std::vector<unsigned long long> iters(C), evals(C);
  std::vector<char> fbest(C * elem);
  auto run = [&]() -> int {
    if (rc != NLS_OK) return rc;
    NLS_CUDA(cudaMemcpyAsync(s.x0, x0_host, size_t(x0_count) * d * elem, cudaMemcpyHostToDevice, st));
    NLS_CUDA(cfg->dtype == NLS_F64 ? nmpso_launch_f64(s, st) : nmpso_launch_f32(s, st));
    if (x_best_host) NLS_CUDA(cudaMemcpyAsync(x_best_host, s.x_best, size_t(C) * d * elem, cudaMemcpyDeviceToHost, st));
    NLS_CUDA(cudaMemcpyAsync(fbest.data(), s.f_best, C * elem, cudaMemcpyDeviceToHost, st));
    NLS_CUDA(cudaMemcpyAsync(iters.data(), s.iters, C * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    NLS_CUDA(cudaMemcpyAsync(evals.data(), s.evals, C * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    NLS_CUDA(cudaStreamSynchronize(st));
    return NLS_OK;
  };
  rc = run();
  mem.release();
  if (rc != NLS_OK) return rc;
  u64 best = 0, total = 0;
  auto value = [&](u64 k) { return cfg->dtype == NLS_F64 ? reinterpret_cast<const double *>(fbest.data())[k]
                                                         : static_cast<double>(reinterpret_cast<const float *>(fbest.data())[k]); };
  for (u64 k = 0; k < C; k++) {
    if (value(k) < value(best)) best = k;
    total += evals[k];
  }
  if (f_best_host) std::memcpy(f_best_host, fbest.data(), C * elem);
  if (iterations_host) for (u64 k = 0; k < C; k++) iterations_host[k] = iters[k];
  if (function_calls_host) for (u64 k = 0; k < C; k++) function_calls_host[k] = evals[k];
  if (status) {
    std::memset(status, 0, sizeof(*status));
    status->f_value = value(best);
    status->iterations = iters[best];
    status->function_calls = total;
    status->best_index = cfg->solver_offset + best;
    status->stopped = 1;
    status->best_valid = 1;
  }
  return NLS_OK;
}

/* ================================================================ objective plugins =========================== */

int nls_load_objective(const char *plugin_path, int32_t *objective_id) {
  if (!plugin_path || !objective_id) return fail(NLS_ERR_INVALID, "nls_load_objective: NULL argument");
  void *h = dlopen(plugin_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail(NLS_ERR_INVALID, "nls_load_objective: %s", dlerror());
  typedef const ObjectivePlugin *(*entry_t)();
  entry_t entry = reinterpret_cast<entry_t>(dlsym(h, "nls_objective_plugin_v1"));
  if (!entry) { dlclose(h); return fail(NLS_ERR_INVALID, "nls_load_objective: %s exports no nls_objective_plugin_v1", plugin_path); }
  const ObjectivePlugin *pl = entry();
  // the ABI number is the first field of every plugin generation, so it is checked before anything else is read; the
  // state structs travel by value into the plugin's launchers, so their sizes must match this build exactly
  if (!pl || pl->abi != kPluginAbi || pl->de_state_bytes != sizeof(DEState) || pl->pso_state_bytes != sizeof(PSOState) ||
      pl->sann_state_bytes != sizeof(SANNState) || !pl->de_f64 || !pl->de_f32 || !pl->pso_f64 || !pl->pso_f32 ||
      !pl->sann_f64 || !pl->sann_f32) {
    dlclose(h);
    return fail(NLS_ERR_INVALID, "nls_load_objective: plugin ABI mismatch (rebuild it against this library's headers)");
  }
  std::lock_guard<std::mutex> lock(g_plugin_mutex);
  const int n = g_plugin_count.load(std::memory_order_relaxed);
  if (n >= kMaxPlugins) { dlclose(h); return fail(NLS_ERR_STATE, "nls_load_objective: at most %d plugins per process", kMaxPlugins); }
  g_plugins[n] = pl;                  // the handle stays open for the life of the process
  g_plugin_count.store(n + 1, std::memory_order_release);
  *objective_id = kFirstPluginId + n;
  return NLS_OK;
}

/* ================================================================ peer exchange =============================== */

int nls_xchg_create(nls_ctx *ctx, uint64_t record_bytes, int world, int rank, nls_xchg **out) {
  if (!ctx || !out) return fail(NLS_ERR_INVALID, "nls_xchg_create: NULL argument");
  *out = nullptr;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || record_bytes < sizeof(RecordHeader))
    return fail(NLS_ERR_INVALID, "nls_xchg_create: world must be 1..%d, 0 <= rank < world", kMaxPeers);
  NLS_CUDA(cudaSetDevice(ctx->device));
  nls_xchg *x = new nls_xchg();
  x->ctx = ctx;
  x->opened = false;
  x->attached = false;
  x->flags_offset = round_up(2 * size_t(world) * record_bytes, 256);
  x->bytes = x->flags_offset + 2 * size_t(world) * sizeof(unsigned long long);
  for (int r = 0; r < kMaxPeers; r++) {
    x->peer_base[r] = nullptr; x->w.records[r] = nullptr; x->w.flags[r] = nullptr;
    x->w.values[r] = nullptr; x->w.counts[r] = 0;
  }
  x->w.world = world; x->w.rank = rank; x->w.record_bytes = record_bytes;
  cudaError_t e = cudaMalloc(&x->base, x->bytes);     // plain cudaMalloc: the allocation must be IPC-exportable
  if (e != cudaSuccess) { delete x; return fail(NLS_ERR_NOMEM, "nls_xchg_create: %s", cudaGetErrorString(e)); }
  NLS_CUDA(cudaMemset(x->base, 0, x->bytes));
  x->w.records[rank] = static_cast<char *>(x->base);
  x->w.flags[rank] = reinterpret_cast<unsigned long long *>(static_cast<char *>(x->base) + x->flags_offset);
  if (world == 1) x->opened = true;
  *out = x;
  return NLS_OK;
}

int nls_xchg_get_handle(nls_xchg *x, void *handle_out) {
  if (!x || !handle_out) return fail(NLS_ERR_INVALID, "nls_xchg_get_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= NLS_XCHG_HANDLE_BYTES, "IPC handle larger than the ABI slot");
  NLS_CUDA(cudaSetDevice(x->ctx->device));
  cudaIpcMemHandle_t h;
  NLS_CUDA(cudaIpcGetMemHandle(&h, x->base));
  std::memset(handle_out, 0, NLS_XCHG_HANDLE_BYTES);
  std::memcpy(handle_out, &h, sizeof(h));
  return NLS_OK;
}

int nls_xchg_open_peers(nls_xchg *x, const void *handles) {
  if (!x || !handles) return fail(NLS_ERR_INVALID, "nls_xchg_open_peers: NULL argument");
  NLS_CUDA(cudaSetDevice(x->ctx->device));
  for (int r = 0; r < x->w.world; r++) {
    if (r == x->w.rank || x->peer_base[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char *>(handles) + size_t(r) * NLS_XCHG_HANDLE_BYTES, sizeof(h));
    NLS_CUDA(cudaIpcOpenMemHandle(&x->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    x->w.records[r] = static_cast<char *>(x->peer_base[r]);
    x->w.flags[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(x->peer_base[r]) + x->flags_offset);
  }
  x->opened = true;
  return NLS_OK;
}

int nls_xchg_destroy(nls_xchg *x) {
  if (!x) return NLS_OK;
  cudaSetDevice(x->ctx->device);
  cudaStreamSynchronize(x->ctx->stream);
  for (int r = 0; r < kMaxPeers; r++)
    if (x->peer_base[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
  cudaFree(x->base);
  delete x;
  return NLS_OK;
}

int nls_pso_attach_exchange(nls_pso *p, nls_xchg *x) {
  if (!p || !x) return fail(NLS_ERR_INVALID, "nls_pso_attach_exchange: NULL argument");
  if (!x->opened) return fail(NLS_ERR_STATE, "nls_pso_attach_exchange: open the peers' handles first");
  if (x->w.record_bytes != p->record_bytes) return fail(NLS_ERR_INVALID, "exchange window record size mismatch");
  if (x->attached) return fail(NLS_ERR_STATE, "nls_pso_attach_exchange: an exchange window serves one swarm; create a new one");
  x->attached = true;
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  p->xchg = x;
  if (p->first_apply_pending) {   // finish the first update_best_positions across the shards (nlsolver.h:2595)
    NLS_CUDA(p->ops->candidate_publish(p->s, x->w, 1, p->g, p->ctx->stream));
    NLS_CUDA(p->ops->gather_apply(p->s, x->w, 1, p->ctx->stream));
    p->first_apply_pending = false;
  }
  return NLS_OK;
}

int nls_de_attach_exchange(nls_de *de, nls_xchg *x) {
  if (!de || !x) return fail(NLS_ERR_INVALID, "nls_de_attach_exchange: NULL argument");
  if (!x->opened) return fail(NLS_ERR_STATE, "nls_de_attach_exchange: open the peers' handles first");
  if (x->w.record_bytes != nls_record_bytes(de->cfg.dtype, de->s.d)) return fail(NLS_ERR_INVALID, "exchange window record size mismatch");
  if (x->attached) return fail(NLS_ERR_STATE, "nls_de_attach_exchange: an exchange window serves one solver; create a new one");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  void *dev = nullptr;
  int rc = de->mem.alloc(&dev, sizeof(XchgWindow));
  if (rc != NLS_OK) return rc;
  NLS_CUDA(cudaMemcpyAsync(dev, &x->w, sizeof(XchgWindow), cudaMemcpyHostToDevice, de->ctx->stream));
  NLS_CUDA(cudaStreamSynchronize(de->ctx->stream));
  x->attached = true;
  de->xchg = x;
  de->s.xw = static_cast<const XchgWindow *>(dev);
  de->graph.reset();                                     // captured launches carry the state by value
  // publish the record of the population as it stands (a best re-scan changes nothing else)
  return de_republish(de);
}

int nls_de_read_exchange(nls_de *de, void *records_host) {
  if (!de || !records_host) return fail(NLS_ERR_INVALID, "nls_de_read_exchange: NULL argument");
  if (!de->xchg) return fail(NLS_ERR_STATE, "nls_de_read_exchange: no exchange window attached");
  NLS_CUDA(cudaSetDevice(de->ctx->device));
  const nls_xchg *x = de->xchg;
  const size_t world = size_t(x->w.world), rb = x->w.record_bytes;
  std::vector<char> win(x->bytes);
  NLS_CUDA(cudaMemcpyAsync(win.data(), x->base, x->bytes, cudaMemcpyDeviceToHost, de->ctx->stream));
  NLS_CUDA(cudaStreamSynchronize(de->ctx->stream));
  // per source island the slot with the higher sequence number: islands that stopped early (or that are a generation
  // apart) have their newest record in either parity slot
  const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(win.data() + x->flags_offset);
  for (size_t r = 0; r < world; r++) {
    const size_t parity = flags[world + r] > flags[r] ? 1 : 0;
    if (flags[parity * world + r] == 0) {
      std::memset(static_cast<char *>(records_host) + r * rb, 0, rb);      // nothing published yet: valid = 0
      continue;
    }
    std::memcpy(static_cast<char *>(records_host) + r * rb, win.data() + (parity * world + r) * rb, rb);
  }
  return NLS_OK;
}

int nls_pso_step_fused(nls_pso *p, uint64_t n_generations) {
  if (!p) return fail(NLS_ERR_INVALID, "nls_pso_step_fused: NULL handle");
  if (!p->xchg) return fail(NLS_ERR_STATE, "nls_pso_step_fused: no exchange window attached");
  if (p->first_apply_pending) return fail(NLS_ERR_STATE, "nls_pso_step_fused: initial exchange pending");
  NLS_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  for (uint64_t g = 0; g < n_generations; g++) {
    NLS_CUDA(p->ops->move(p->s, p->g, st));
    NLS_CUDA(p->ops->candidate_publish(p->s, p->xchg->w, 0, p->g, st));
    NLS_CUDA(p->ops->gather_apply(p->s, p->xchg->w, 0, st));
    p->enqueued++;
  }
  return NLS_OK;
}


/* ================================================================ device groups =============================== */

struct nls_group {
  std::vector<nls_ctx *> ctx;
};
struct nls_pso_sharded {
  nls_group *g;
  std::vector<nls_pso *> shard;
  std::vector<nls_xchg *> win;
  std::vector<GraphCache> graph;     // kGraphGens fused generations per device
  bool small;                        // launch-bound shards: replay graphs
};
struct nls_de_islands {
  nls_group *g;
  std::vector<nls_de *> isl;
  uint64_t migrate_every, k, generation;
  std::vector<void *> out_rows, out_scores, in_rows, in_scores;
  std::vector<cudaEvent_t> exported, imported;   // per island: emigrants ready / immigrants consumed
  size_t elem;
};

int nls_group_create(int n_devices, const int *devices, nls_group **out) {
  if (!out) return fail(NLS_ERR_INVALID, "nls_group_create: out is NULL");
  *out = nullptr;
  if (n_devices < 1 || n_devices > kMaxPeers) return fail(NLS_ERR_INVALID, "nls_group_create: 1 .. %d devices", kMaxPeers);
  for (int a = 0; a < n_devices; a++)
    for (int b = a + 1; b < n_devices; b++)
      if (devices && devices[a] == devices[b])
        return fail(NLS_ERR_INVALID, "nls_group_create: device %d listed twice (shards of one group wait on one another "
                                     "and must run on distinct devices)", devices[a]);
  nls_group *g = new nls_group();
  for (int r = 0; r < n_devices; r++) {
    nls_ctx *c = nullptr;
    int rc = nls_ctx_create(devices ? devices[r] : r, nullptr, &c);
    if (rc != NLS_OK) { nls_group_destroy(g); return rc; }
    g->ctx.push_back(c);
  }
  for (int a = 0; a < n_devices; a++)
    for (int b = 0; b < n_devices; b++) {
      if (a == b) continue;
      int can = 0;
      cudaSetDevice(g->ctx[a]->device);
      cudaDeviceCanAccessPeer(&can, g->ctx[a]->device, g->ctx[b]->device);
      if (!can) { nls_group_destroy(g); return fail(NLS_ERR_CUDA, "device %d cannot map the memory of device %d", g->ctx[a]->device, g->ctx[b]->device); }
      cudaError_t e = cudaDeviceEnablePeerAccess(g->ctx[b]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        nls_group_destroy(g);
        return fail(NLS_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", g->ctx[a]->device, g->ctx[b]->device, cudaGetErrorString(e));
      }
      cudaGetLastError();
    }
  *out = g;
  return NLS_OK;
}
int nls_group_destroy(nls_group *g) {
  if (!g) return NLS_OK;
  for (nls_ctx *c : g->ctx) nls_ctx_destroy(c);
  delete g;
  return NLS_OK;
}
int nls_group_size(const nls_group *g) { return g ? int(g->ctx.size()) : -1; }

/* ---- sharded swarm ---- */
static void slice_of(u64 n_global, int world, int rank, u64 *begin, u64 *end) {
  // contiguous slices; the first n_global % world ranks hold one extra element (nlsolver_b200/distributed.py: slice_bounds)
  const u64 base = n_global / world, extra = n_global % world;
  *begin = rank * base + std::min<u64>(rank, extra);
  *end = *begin + base + (u64(rank) < extra ? 1 : 0);
}

int nls_pso_sharded_destroy(nls_pso_sharded *h) {
  if (!h) return NLS_OK;
  for (nls_pso *p : h->shard)
    if (p) { cudaSetDevice(p->ctx->device); cudaStreamSynchronize(p->ctx->stream); }
  for (size_t r = 0; r < h->graph.size(); r++) {
    if (r < h->shard.size() && h->shard[r]) cudaSetDevice(h->shard[r]->ctx->device);
    h->graph[r].reset();
  }
  for (nls_pso *p : h->shard) nls_pso_destroy(p);
  for (nls_xchg *x : h->win) nls_xchg_destroy(x);
  delete h;
  return NLS_OK;
}

int nls_pso_sharded_create(nls_group *g, const nls_pso_cfg *cfg, const void *lower_host, const void *upper_host,
                           nls_pso_sharded **out) {
  if (!g || !cfg || !lower_host || !upper_host || !out) return fail(NLS_ERR_INVALID, "nls_pso_sharded_create: NULL argument");
  *out = nullptr;
  const int world = int(g->ctx.size());
  if (cfg->n_particles < u64(world)) return fail(NLS_ERR_INVALID, "nls_pso_sharded_create: fewer particles than devices");
  nls_pso_sharded *h = new nls_pso_sharded();
  h->g = g;
  h->graph.resize(world);
  int rc = NLS_OK;
  for (int r = 0; r < world && rc == NLS_OK; r++) {
    nls_pso_cfg local = *cfg;
    u64 b, e;
    slice_of(cfg->n_particles, world, r, &b, &e);
    local.n_particles = e - b;
    local.particle_offset = b;
    local.n_particles_global = cfg->n_particles;
    nls_pso *p = nullptr;
    rc = nls_pso_create(g->ctx[r], &local, lower_host, upper_host, &p);
    h->shard.push_back(p);
    nls_xchg *x = nullptr;
    if (rc == NLS_OK) rc = nls_xchg_create(g->ctx[r], nls_record_bytes(cfg->dtype, cfg->dim), world, r, &x);
    h->win.push_back(x);
  }
  if (rc != NLS_OK) { nls_pso_sharded_destroy(h); return rc; }
  // every window is addressable from every device of the group (peer access): link them directly, no IPC handles
  for (int r = 0; r < world; r++) {
    for (int q = 0; q < world; q++) {
      h->win[r]->w.records[q] = h->win[q]->w.records[q];
      h->win[r]->w.flags[q] = h->win[q]->w.flags[q];
      h->win[r]->w.values[q] = h->shard[q]->s.pbest;
      h->win[r]->w.counts[q] = h->shard[q]->s.P;
    }
    h->win[r]->opened = true;
  }
  h->small = h->shard[0]->s.P * h->shard[0]->s.d <= kGraphMaxElems;
  // the first update_best_positions across the shards (nlsolver.h:2595): enqueue every shard's publish + gather
  for (int r = 0; r < world && rc == NLS_OK; r++) rc = nls_pso_attach_exchange(h->shard[r], h->win[r]);
  if (rc != NLS_OK) { nls_pso_sharded_destroy(h); return rc; }
  *out = h;
  return NLS_OK;
}

int nls_pso_sharded_step(nls_pso_sharded *h, uint64_t n_generations) {
  if (!h) return fail(NLS_ERR_INVALID, "nls_pso_sharded_step: NULL handle");
  const int world = int(h->shard.size());
  uint64_t left = n_generations;
  // Every device gets the same work in the same order; a shard's apply kernel waits (on the device) for the records of
  // all shards, so the host must enqueue generation g on every device before it may block on anything.
  if (h->small) {
    while (left >= kGraphGens) {
      bool ok = true;
      for (int r = 0; r < world && ok; r++) {
        nls_pso *p = h->shard[r];
        cudaSetDevice(p->ctx->device);
        cudaStream_t st = p->ctx->stream;
        ok = graph_replay(h->graph[r], st, [&] {
          for (int g = 0; g < kGraphGens; g++) {
            if (p->ops->move(p->s, p->g, st) != cudaSuccess) return false;
            if (p->ops->candidate_publish(p->s, p->xchg->w, 0, p->g, st) != cudaSuccess) return false;
            if (p->ops->gather_apply(p->s, p->xchg->w, 0, st) != cudaSuccess) return false;
          }
          return true;
        });
        if (ok) p->enqueued += kGraphGens;
        else if (r > 0) return fail(NLS_ERR_CUDA, "nls_pso_sharded_step: graph launch failed on device %d after other shards were enqueued", p->ctx->device);
      }
      if (!ok) { h->small = false; break; }
      left -= kGraphGens;
    }
  }
  for (uint64_t g = 0; g < left; g++)
    for (int r = 0; r < world; r++) {
      int rc = nls_pso_step_fused(h->shard[r], 1);
      if (rc != NLS_OK) return rc;
    }
  return NLS_OK;
}

int nls_pso_sharded_sync(nls_pso_sharded *h, nls_status *status) {
  if (!h) return fail(NLS_ERR_INVALID, "nls_pso_sharded_sync: NULL handle");
  int rc = NLS_OK;
  nls_status st0;
  for (size_t r = 0; r < h->shard.size() && rc == NLS_OK; r++) {
    nls_status st;
    rc = nls_pso_sync(h->shard[r], &st);
    if (r == 0) st0 = st;      // every shard holds the same swarm-level state
  }
  if (rc == NLS_OK && status) *status = st0;
  return rc;
}
int nls_pso_sharded_read_best(nls_pso_sharded *h, void *x_host) {
  if (!h || !x_host) return fail(NLS_ERR_INVALID, "nls_pso_sharded_read_best: NULL argument");
  return nls_pso_read_best(h->shard[0], x_host);
}
int nls_pso_sharded_shard(nls_pso_sharded *h, int rank, nls_pso **shard) {
  if (!h || !shard || rank < 0 || rank >= int(h->shard.size())) return fail(NLS_ERR_INVALID, "nls_pso_sharded_shard: bad argument");
  *shard = h->shard[rank];
  return NLS_OK;
}

int nls_pso_solve_sharded(nls_group *g, const nls_pso_cfg *cfg, const void *lower_host, const void *upper_host,
                          void *x_best_host, nls_status *status) {
  if (!x_best_host) return fail(NLS_ERR_INVALID, "nls_pso_solve_sharded: x_best_host is NULL");
  nls_pso_sharded *h = nullptr;
  int rc = nls_pso_sharded_create(g, cfg, lower_host, upper_host, &h);
  if (rc != NLS_OK) return rc;
  nls_status st;
  rc = nls_pso_sharded_sync(h, &st);
  uint64_t batch = 8;
  while (rc == NLS_OK && !st.stopped) {
    const uint64_t left = cfg->max_iter > st.iterations ? cfg->max_iter - st.iterations : 1;
    rc = nls_pso_sharded_step(h, std::min<uint64_t>(batch, left));
    if (rc == NLS_OK) rc = nls_pso_sharded_sync(h, &st);
    if (batch < 256) batch *= 2;
  }
  if (rc == NLS_OK && st.best_valid) rc = nls_pso_sharded_read_best(h, x_best_host);
  if (rc == NLS_OK && status) *status = st;
  nls_pso_sharded_destroy(h);
  return rc;
}

/* ---- DE islands ---- */
int nls_de_islands_destroy(nls_de_islands *h) {
  if (!h) return NLS_OK;
  for (size_t r = 0; r < h->isl.size(); r++) {
    if (!h->isl[r]) continue;
    cudaSetDevice(h->isl[r]->ctx->device);
    cudaStreamSynchronize(h->isl[r]->ctx->stream);
  }
  for (size_t r = 0; r < h->isl.size(); r++) {
    if (h->isl[r]) cudaSetDevice(h->isl[r]->ctx->device);
    if (r < h->exported.size() && h->exported[r]) cudaEventDestroy(h->exported[r]);
    if (r < h->imported.size() && h->imported[r]) cudaEventDestroy(h->imported[r]);
    for (std::vector<void *> *v : {&h->out_rows, &h->out_scores, &h->in_rows, &h->in_scores})
      if (r < v->size() && (*v)[r]) cudaFree((*v)[r]);
    nls_de_destroy(h->isl[r]);
  }
  delete h;
  return NLS_OK;
}

int nls_de_islands_create(nls_group *g, const nls_de_cfg *cfg, const void *x0_host, uint64_t migrate_every,
                          uint64_t migrants, nls_de_islands **out) {
  if (!g || !cfg || !x0_host || !out) return fail(NLS_ERR_INVALID, "nls_de_islands_create: NULL argument");
  *out = nullptr;
  const int world = int(g->ctx.size());
  nls_de_islands *h = new nls_de_islands();
  h->g = g;
  h->migrate_every = migrate_every;
  h->k = std::min<uint64_t>(migrants, cfg->pop_size);
  h->generation = 0;
  h->elem = elem_size(cfg->dtype);
  int rc = NLS_OK;
  for (int r = 0; r < world && rc == NLS_OK; r++) {
    nls_de_cfg local = *cfg;
    local.agent_offset = cfg->agent_offset + u64(r) * cfg->pop_size;   // islands draw from disjoint streams
    nls_de *de = nullptr;
    rc = nls_de_create(g->ctx[r], &local, x0_host, &de);
    h->isl.push_back(de);
    h->exported.push_back(nullptr); h->imported.push_back(nullptr);
    h->out_rows.push_back(nullptr); h->out_scores.push_back(nullptr);
    h->in_rows.push_back(nullptr); h->in_scores.push_back(nullptr);
    if (rc != NLS_OK || h->k == 0 || world == 1) continue;
    const size_t rows = size_t(h->k) * cfg->dim * h->elem, scores = size_t(h->k) * h->elem;
    cudaError_t e = cudaSetDevice(g->ctx[r]->device);
    if (e == cudaSuccess) e = cudaMalloc(&h->out_rows[r], rows);
    if (e == cudaSuccess) e = cudaMalloc(&h->out_scores[r], scores);
    if (e == cudaSuccess) e = cudaMalloc(&h->in_rows[r], rows);
    if (e == cudaSuccess) e = cudaMalloc(&h->in_scores[r], scores);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->exported[r], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->imported[r], cudaEventDisableTiming);
    if (e != cudaSuccess) rc = fail(NLS_ERR_CUDA, "nls_de_islands_create: %s", cudaGetErrorString(e));
  }
  if (rc != NLS_OK) { nls_de_islands_destroy(h); return rc; }
  *out = h;
  return NLS_OK;
}

// ring migration after a generation that is a multiple of migrate_every: island r sends its k best rows to island
// r + 1, which overwrites its k worst with them (same selection orders as nls_de_export_top / nls_de_import_migrants)
static int islands_migrate(nls_de_islands *h) {
  const int world = int(h->isl.size());
  const size_t rows = size_t(h->k) * h->isl[0]->s.d * h->elem, scores = size_t(h->k) * h->elem;
  for (int r = 0; r < world; r++) {          // emigrants: wait until the previous batch was consumed by the receiver
    nls_de *de = h->isl[r];
    NLS_CUDA(cudaSetDevice(de->ctx->device));
    NLS_CUDA(cudaStreamWaitEvent(de->ctx->stream, h->imported[(r + 1) % world], 0));
    int rc = nls_de_export_top(de, h->k, h->out_rows[r], h->out_scores[r]);
    if (rc != NLS_OK) return rc;
    NLS_CUDA(cudaEventRecord(h->exported[r], de->ctx->stream));
  }
  for (int r = 0; r < world; r++) {          // immigrants: pull from the left neighbour over NVLink, then import
    nls_de *de = h->isl[r];
    const int src = (r + world - 1) % world;
    NLS_CUDA(cudaSetDevice(de->ctx->device));
    NLS_CUDA(cudaStreamWaitEvent(de->ctx->stream, h->exported[src], 0));
    NLS_CUDA(cudaMemcpyPeerAsync(h->in_rows[r], de->ctx->device, h->out_rows[src], h->isl[src]->ctx->device, rows, de->ctx->stream));
    NLS_CUDA(cudaMemcpyPeerAsync(h->in_scores[r], de->ctx->device, h->out_scores[src], h->isl[src]->ctx->device, scores, de->ctx->stream));
    NLS_CUDA(cudaEventRecord(h->imported[r], de->ctx->stream));
    int rc = nls_de_import_migrants(de, h->k, h->in_rows[r], h->in_scores[r]);
    if (rc != NLS_OK) return rc;
  }
  return NLS_OK;
}

int nls_de_islands_step(nls_de_islands *h, uint64_t n_generations) {
  if (!h) return fail(NLS_ERR_INVALID, "nls_de_islands_step: NULL handle");
  const int world = int(h->isl.size());
  const bool migrating = world > 1 && h->migrate_every > 0 && h->k > 0;
  uint64_t left = n_generations;
  while (left > 0) {
    // run up to the next migration point in one call per island (small islands: one launch each)
    uint64_t chunk = left;
    if (migrating) chunk = std::min<uint64_t>(chunk, h->migrate_every - h->generation % h->migrate_every);
    for (int r = 0; r < world; r++) {
      int rc = nls_de_step(h->isl[r], chunk);
      if (rc != NLS_OK) return rc;
    }
    h->generation += chunk;
    left -= chunk;
    if (migrating && h->generation % h->migrate_every == 0) {
      int rc = islands_migrate(h);
      if (rc != NLS_OK) return rc;
    }
  }
  return NLS_OK;
}

int nls_de_islands_sync(nls_de_islands *h, nls_status *status) {
  if (!h) return fail(NLS_ERR_INVALID, "nls_de_islands_sync: NULL handle");
  nls_status best;
  std::memset(&best, 0, sizeof(best));
  uint64_t calls = 0, reruns = 0, rounds = 0, accepted = 0;
  int all_stopped = 1;
  for (size_t r = 0; r < h->isl.size(); r++) {
    nls_status st;
    int rc = nls_de_sync(h->isl[r], &st);
    if (rc != NLS_OK) return rc;
    calls += st.function_calls; reruns += st.repair_reruns; rounds += st.repair_rounds; accepted += st.accepted_total;
    all_stopped &= st.stopped;
    const uint64_t iter0 = r == 0 ? st.iterations : best.iterations;
    if (r == 0 || st.f_value < best.f_value) {        // strict <: the lowest rank wins ties
      best = st;
      best.best_index = h->isl[r]->s.offset + st.best_index;
      best._reserved = int32_t(r);                    // rank of the best island
    }
    best.iterations = iter0;
  }
  best.function_calls = calls; best.repair_reruns = reruns; best.repair_rounds = rounds; best.accepted_total = accepted;
  best.stopped = all_stopped;
  if (status) *status = best;
  return NLS_OK;
}
int nls_de_islands_read_best(nls_de_islands *h, void *x_host) {
  if (!h || !x_host) return fail(NLS_ERR_INVALID, "nls_de_islands_read_best: NULL argument");
  nls_status st;
  int rc = nls_de_islands_sync(h, &st);
  if (rc != NLS_OK) return rc;
  return nls_de_read_best(h->isl[st._reserved], x_host);
}
int nls_de_islands_island(nls_de_islands *h, int rank, nls_de **island) {
  if (!h || !island || rank < 0 || rank >= int(h->isl.size())) return fail(NLS_ERR_INVALID, "nls_de_islands_island: bad argument");
  *island = h->isl[rank];
  return NLS_OK;
}
int nls_de_solve_islands(nls_group *g, const nls_de_cfg *cfg, const void *x0_host, uint64_t migrate_every,
                         uint64_t migrants, void *x_best_host, nls_status *status) {
  if (!x_best_host) return fail(NLS_ERR_INVALID, "nls_de_solve_islands: x_best_host is NULL");
  nls_de_islands *h = nullptr;
  int rc = nls_de_islands_create(g, cfg, x0_host, migrate_every, migrants, &h);
  if (rc != NLS_OK) return rc;
  nls_status st;
  rc = nls_de_islands_sync(h, &st);
  uint64_t batch = migrate_every ? migrate_every : 8;
  while (rc == NLS_OK && !st.stopped) {
    const uint64_t left = cfg->max_iter > st.iterations ? cfg->max_iter - st.iterations : 1;
    rc = nls_de_islands_step(h, std::min<uint64_t>(batch, left));
    if (rc == NLS_OK) rc = nls_de_islands_sync(h, &st);
    if (batch < 256) batch *= 2;
  }
  if (rc == NLS_OK) rc = nls_de_islands_read_best(h, x_best_host);
  if (rc == NLS_OK && status) *status = st;
  nls_de_islands_destroy(h);
  return rc;
}

}  /* extern "C" */
