// rn_abi.cu -- the extern "C" boundary declared in include/rednoise.h.
//
// Replaces the field-synthesis half of Terrain::Terrain (renderEngine/source/
// Terrain.cpp:44-88).  Host logic only: parameter validation, scratch ownership,
// algorithm dispatch, window extraction, host<->device copies.  There is no CPU
// implementation behind any compute entry point.
#include "../../include/rednoise.h"
#include "rn_kernels.h"
#include <cstdlib>
#include <stdio.h>
#include <string.h>
#include <string>
#include <new>
#include <vector>
#include <algorithm>
#include <atomic>
#include <thread>

using namespace rn;

struct rn_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  int algo = 0;
  uint64_t launches = 0;
  int sm_count = 0;
  // grow-only scratch
  void* field = nullptr;   size_t field_bytes = 0;    // full-field staging when dst is a window / host memory
  void* plane = nullptr;   size_t plane_bytes = 0;    // algorithm 1: offset plane; algorithm 2: coarse grids
  void* cells = nullptr;   size_t cells_bytes = 0;    // staging for rn_terrain2d_cells with a host destination
  void* small = nullptr;   size_t small_bytes = 0;    // staging for rn_lcg_draws_device with a host dst; status words
  // batch seeds: two buffers used alternately by consecutive rn_terrain2d_batch calls.  seeds_ready[b] is recorded on the
  // caller's stream after the upload (the side streams wait for it before their first stage), seeds_free[b] after the last
  // kernel of the call that reads buffer b (the next upload into b waits for it, whichever stream it is issued on).
  void* seeds[2] = {nullptr, nullptr};   size_t seeds_bytes[2] = {0, 0};
  cudaEvent_t seeds_ready[2] = {nullptr, nullptr};
  cudaEvent_t seeds_free[2] = {nullptr, nullptr};
  bool seeds_used[2] = {false, false};
  uint64_t batch_seq = 0;
  // algorithm 2 pipelining: the coarse stages run on a high-priority side stream into one of two
  // scratch slots, so that they overlap the output stage of the previous call
  cudaStream_t coarse_stream = nullptr;
  cudaStream_t tile_stream = nullptr;               // the tile stages of a call run here, ahead of its streamed coarse level
  cudaEvent_t tiles_done[2] = {nullptr, nullptr};   // tile stages of the call using slot s have finished
  void* slot[2] = {nullptr, nullptr};
  size_t slot_bytes[2] = {0, 0};
  cudaEvent_t coarse_done[2] = {nullptr, nullptr};  // coarse stages of the call using slot s have finished
  cudaEvent_t slot_free[2] = {nullptr, nullptr};    // the output stage that read slot s has finished
  bool slot_used[2] = {false, false};
  uint64_t seq = 0;
  int overlap = 1;
  // host destinations: the window is produced in row chunks, chunk k+1 is synthesized while chunk k crosses PCIe
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t chunk_ready[2] = {nullptr, nullptr};   // synthesis of the chunk in staging half h has finished
  cudaEvent_t chunk_copied[2] = {nullptr, nullptr};  // its device-to-host copy has finished
  void* bounce = nullptr;  size_t bounce_bytes = 0;  // pinned host halves a PAGEABLE destination is filled through (grow-only)
  // stage profiling
  int profiling = 0;
  cudaEvent_t ev[17] = {};
  int ev_stages = 0;
};

#ifndef RN_TILE_STREAM
#define RN_TILE_STREAM 1   // a third stream for the tile stages, so that they need not wait for the previous call's streamed coarse level
#endif

namespace {

thread_local std::string g_last_error;

int fail(int status, const char* what)
{
  g_last_error = what ? what : "";
  return status;
}
int fail_cuda(cudaError_t e, const char* where)
{
  char buf[512];
  snprintf(buf, sizeof buf, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  g_last_error = buf;
  return (e == cudaErrorMemoryAllocation) ? RN_ERR_OOM : RN_ERR_CUDA;
}
#define RN_CUDA(call)                                    \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
  } while (0)

int check_params(const rn_params* p)
{
  if (!p) return fail(RN_ERR_BAD_PARAM, "params is NULL");
  if (p->depth < RN_MIN_DEPTH || p->depth > RN_MAX_DEPTH) return fail(RN_ERR_BAD_PARAM, "depth out of range [3,14]");
  if (p->lacunarity != 2) return fail(RN_ERR_UNSUPPORTED, "only lacunarity 2 is supported (the reference halves the step, Terrain.cpp:81)");
  if (!(p->persistence > 0.0f) || !(p->persistence <= 1.0f)) return fail(RN_ERR_BAD_PARAM, "persistence must be in (0,1]");
  if (!(p->feature_scale != 0.0f)) return fail(RN_ERR_BAD_PARAM, "feature_scale must be non-zero");
  return RN_OK;
}

Law law_of(const rn_params* p)
{
  Law l;
  l.amplitude = p->amplitude;
  l.sideways_scale = p->sideways_scale;
  l.radiosity_strength = p->radiosity_strength;
  l.amplitude_variation = p->amplitude_variation;
  l.band_hi = 3.0f / p->feature_scale;  // Terrain.cpp:191,235
  l.band_lo = 0.5f / p->feature_scale;
  l.bulgy = p->bulgy_rocks != 0;
  l.one = 1.0f;
  l.flat_ok = law_flat_ok(l);
  return l;
}

int ensure(void** buf, size_t* have, size_t want)
{
  if (*have >= want) return RN_OK;
  if (*buf) {
    cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
  }
  cudaError_t e = cudaMalloc(buf, want);
  if (e != cudaSuccess) {
    *buf = nullptr;
    return fail_cuda(e, "cudaMalloc(scratch)");
  }
  *have = want;
  return RN_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev)
  {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Algorithm 1: the full (W+1)^2 field of `p` into device memory g (row pitch = pitch elements).
int synth_full_levelsync(rn_context* ctx, const rn_params* p, Elem* g, size_t pitch, cudaStream_t stream)
{
  const int W = 1 << p->depth;
  const Law law = law_of(p);
  int st = ensure(&ctx->plane, &ctx->plane_bytes, sizeof(float) * pitch * (size_t)(W + 1));
  if (st != RN_OK) return st;
  RN_CUDA(levelsync_generate(g, (float*)ctx->plane, pitch, p->depth, p->depth - 2, law, p->persistence, p->seed, stream, &ctx->launches));
  return RN_OK;
}

// Algorithm 2: the window `w` of `nfields` fields (seeds on the device, or p->seed) straight into device memory.
// `after` (optional): an event on the caller's stream that every stage of the call must follow (the seeds upload of a batch).
int synth_window_fused(rn_context* ctx, const rn_params* p, const rn_window& w, Elem* g, size_t pitch, size_t field_stride,
                       const uint32_t* dev_seeds, int nfields, cudaStream_t stream, cudaEvent_t after = nullptr)
{
  const Law law = law_of(p);
  const size_t need = fused_scratch_bytes(p->depth, nfields);
  const int sl = (int)(ctx->seq & 1);
  ctx->seq++;
  int st = ensure(&ctx->slot[sl], &ctx->slot_bytes[sl], need ? need : 256);
  if (st != RN_OK) return st;
  if (ctx->profiling && !ctx->ev[0])
    for (int k = 0; k < 17; k++) RN_CUDA(cudaEventCreate(&ctx->ev[k]));
  ctx->ev_stages = 0;
  const bool overlap = ctx->overlap && !ctx->profiling && ctx->coarse_stream;
  // The previous user of this scratch slot must have consumed it, and the call's inputs (batch seeds) must have landed, before
  // ANY stream of this call starts: every side stream waits, whichever of them the plan happens to use first.
  cudaStream_t first[3] = {stream, nullptr, nullptr};
  int nfirst = 1;
  if (overlap) {
    nfirst = 0;
    first[nfirst++] = ctx->coarse_stream;
    if (ctx->tile_stream) first[nfirst++] = ctx->tile_stream;
  }
  for (int k = 0; k < nfirst; k++) {
    if (ctx->slot_used[sl]) RN_CUDA(cudaStreamWaitEvent(first[k], ctx->slot_free[sl], 0));
    if (after && first[k] != stream) RN_CUDA(cudaStreamWaitEvent(first[k], after, 0));
  }
  RN_CUDA(fused_generate(g, pitch, field_stride, p->depth, w.i0, w.j0, w.ni, w.nj, law, p->persistence, p->seed, dev_seeds, nfields,
                         p->apply_volcano != 0, p->amplitude, ctx->slot[sl], stream, &ctx->launches, ctx->profiling ? ctx->ev : nullptr,
                         ctx->profiling ? &ctx->ev_stages : nullptr, overlap ? ctx->coarse_stream : nullptr,
                         overlap ? ctx->coarse_done[sl] : nullptr, overlap ? ctx->tile_stream : nullptr, overlap ? ctx->tiles_done[sl] : nullptr,
                         ctx->sm_count, ctx->algo != 3));
  RN_CUDA(cudaEventRecord(ctx->slot_free[sl], stream));
  ctx->slot_used[sl] = true;
  return RN_OK;
}

// Fields per chunk of a batch: staging (host destination) and the intermediate grids (two scratch slots) stay bounded -- the
// scratch of a chunk scales with its field count (21 MB per depth-11 field, 84 MB per depth-12 field) -- and, for a device
// destination, the field count is one for which the output level's units fill the warp slots evenly (rn_plan.h march_batch_chunk).
size_t batch_chunk(const rn_context* ctx, const rn_params* p, int ntiles, bool direct)
{
  const size_t side = (size_t)(1 << p->depth) + 1;
  const size_t field_bytes = side * side * sizeof(Elem);
  const size_t scratch_per_field = fused_scratch_bytes(p->depth, 2) / 2;
  const size_t budget = (size_t)1 << 30;
  size_t chunk = direct ? budget / (scratch_per_field ? scratch_per_field : 1) : budget / field_bytes;
  if (chunk < 1) chunk = 1;
  if (chunk > 32768) chunk = 32768;  // gridDim.z
  if (direct && ntiles > 1) {
    const char* forced = getenv("RN_BATCH_CHUNK");  // experiments
    const int want = forced ? atoi(forced) : fused_batch_chunk(p->depth, (int)chunk, ctx->sm_count);
    if (want >= 1 && (size_t)want < chunk) chunk = (size_t)want;
  }
  // (equal chunks were measured and are worse: the planner gives every field of a chunk the same whole number of (strip, segment)
  //  units, and 229 depth-10 fields x 9 units fill the 2072 warp slots where 128 x 11 do not -- 256 tiles: 115 vs 84 Gsamples/s)
  return chunk;
}

// Large host destination: the window is produced in row chunks through two staging halves, so that synthesis of chunk k+1
// overlaps the PCIe transfer of chunk k (the transfer is the bottleneck: 16 B/sample at ~55 GB/s).  Everything is enqueued
// (synthesis on `stream`, copies on ctx->copy_stream); the caller waits for ctx->copy_stream.
int enqueue_window_to_host(rn_context* ctx, const rn_params* p, const rn_window& w, rn_element* dst, size_t pitch_elems, cudaStream_t stream)
{
  int rows = (int)(((size_t)32 << 20) / ((size_t)w.nj * sizeof(Elem)));  // ~32 MB per chunk
  rows = rows < 64 ? 64 : (rows / 64) * 64;
  const size_t half = (size_t)rows * (size_t)w.nj;
  int st = ensure(&ctx->field, &ctx->field_bytes, sizeof(Elem) * 2 * half);
  if (st != RN_OK) return st;
  // A PAGEABLE destination (what `new TerrainElement[]`, Terrain.cpp:45-49, gives an unmodified caller): the runtime would stage such a
  // copy itself, at ~18 GB/s.  Instead every chunk crosses PCIe into one of two pinned halves at the full rate and a few host threads
  // move it on into the destination while the next chunk is in flight.
  cudaPointerAttributes attr{};
  const bool pageable = cudaPointerGetAttributes(&attr, dst) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  if (pageable) {
    if (ctx->bounce_bytes < sizeof(Elem) * 2 * half) {
      if (ctx->bounce) cudaFreeHost(ctx->bounce);
      ctx->bounce = nullptr;
      ctx->bounce_bytes = 0;
      cudaError_t e = cudaHostAlloc(&ctx->bounce, sizeof(Elem) * 2 * half, cudaHostAllocDefault);
      if (e != cudaSuccess) return fail_cuda(e, "cudaHostAlloc(bounce)");
      ctx->bounce_bytes = sizeof(Elem) * 2 * half;
    }
  }
  const int nchunks = (w.ni + rows - 1) / rows;
  // host side of the pageable path: chunk k may be moved once `landed` > k; `moved` counts the workers that finished a chunk
  const unsigned hc = std::thread::hardware_concurrency();
  const int nworkers = pageable ? (int)std::min<unsigned>(8u, hc > 1 ? hc / 2 : 1u) : 0;
  std::atomic<int> landed{0};
  std::vector<std::atomic<int>> moved(pageable ? (size_t)nchunks : 0);
  for (auto& m : moved) m.store(0);
  std::vector<std::thread> workers;
  auto move_chunk = [&](int k, int part, int parts) {
    const int r0 = k * rows, nr = (w.ni - r0 < rows) ? w.ni - r0 : rows;
    const Elem* src = reinterpret_cast<const Elem*>(ctx->bounce) + (size_t)(k & 1) * half;
    const int a = (int)((long long)nr * part / parts), b = (int)((long long)nr * (part + 1) / parts);
    for (int r = a; r < b; r++) memcpy(dst + (size_t)(r0 + r) * pitch_elems, src + (size_t)r * (size_t)w.nj, (size_t)w.nj * sizeof(Elem));
  };
  for (int t = 0; t < nworkers; t++)
    workers.emplace_back([&, t] {
      for (int k = 0; k < nchunks; k++) {
        while (landed.load(std::memory_order_acquire) <= k) {
          if (landed.load(std::memory_order_acquire) < 0) return;  // the call failed
          std::this_thread::yield();
        }
        move_chunk(k, t, nworkers);
        moved[(size_t)k].fetch_add(1, std::memory_order_release);
      }
    });
  auto fail_workers = [&](int status) {
    landed.store(-1, std::memory_order_release);
    for (auto& t : workers) t.join();
    return status;
  };
  auto chunk_landed = [&](int k) -> int {  // the host waits for chunk k's transfer and releases it to the workers
    cudaError_t e = cudaEventSynchronize(ctx->chunk_copied[k & 1]);
    if (e != cudaSuccess) return fail_cuda(e, "cudaEventSynchronize(chunk)");
    landed.store(k + 1, std::memory_order_release);
    return RN_OK;
  };
#define RN_CUDA_W(call)                                                            \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) return fail_workers(fail_cuda(e_, #call));              \
  } while (0)
  for (int k = 0; k < nchunks; k++) {
    const int r0 = k * rows;
    const int nr = (w.ni - r0 < rows) ? w.ni - r0 : rows;
    const int h = k & 1;
    Elem* stage = reinterpret_cast<Elem*>(ctx->field) + (size_t)h * half;
    if (pageable && k >= 2) {
      // the pinned half is free again once chunk k - 2 has left it: release chunk k - 1 first, then wait for the workers
      // (the copy queued for chunk k - 1 is already running)
      st = chunk_landed(k - 1);
      if (st != RN_OK) return fail_workers(st);
      while (moved[(size_t)(k - 2)].load(std::memory_order_acquire) < nworkers) std::this_thread::yield();
    }
    if (k >= 2) RN_CUDA_W(cudaStreamWaitEvent(stream, ctx->chunk_copied[h], 0));  // the staging half is free again
    const rn_window cw{w.i0 + r0, w.j0, nr, w.nj};
    st = synth_window_fused(ctx, p, cw, stage, (size_t)w.nj, 0, nullptr, 1, stream);
    if (st != RN_OK) return fail_workers(st);
    RN_CUDA_W(cudaEventRecord(ctx->chunk_ready[h], stream));
    RN_CUDA_W(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ready[h], 0));
    if (pageable)
      RN_CUDA_W(cudaMemcpyAsync(reinterpret_cast<Elem*>(ctx->bounce) + (size_t)h * half, stage, (size_t)nr * (size_t)w.nj * sizeof(Elem),
                                cudaMemcpyDeviceToHost, ctx->copy_stream));
    else if (pitch_elems == (size_t)w.nj)  // dense rows: one linear copy
      RN_CUDA_W(cudaMemcpyAsync(dst + (size_t)r0 * pitch_elems, stage, (size_t)nr * (size_t)w.nj * sizeof(Elem), cudaMemcpyDeviceToHost, ctx->copy_stream));
    else
      RN_CUDA_W(cudaMemcpy2DAsync(dst + (size_t)r0 * pitch_elems, pitch_elems * sizeof(Elem), stage, (size_t)w.nj * sizeof(Elem),
                                  (size_t)w.nj * sizeof(Elem), (size_t)nr, cudaMemcpyDeviceToHost, ctx->copy_stream));
    RN_CUDA_W(cudaEventRecord(ctx->chunk_copied[h], ctx->copy_stream));
  }
  if (pageable) {
    for (int k = std::max(0, nchunks - 2); k < nchunks; k++) {  // (for k >= 2 the loop released chunks 0 .. k - 1)
      if (landed.load(std::memory_order_acquire) > k) continue;
      st = chunk_landed(k);
      if (st != RN_OK) return fail_workers(st);
    }
    for (auto& t : workers) t.join();
  }
#undef RN_CUDA_W
  return RN_OK;
}

}  // namespace

extern "C" {

int rn_version(void) { return RN_VERSION; }
const char* rn_last_error_string(void) { return g_last_error.c_str(); }

const char* rn_status_string(int status)
{
  switch (status) {
    case RN_OK: return "ok";
    case RN_ERR_BAD_PARAM: return "bad parameter";
    case RN_ERR_UNSUPPORTED: return "unsupported parameter value";
    case RN_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required)";
    case RN_ERR_CUDA: return "CUDA error";
    case RN_ERR_OOM: return "out of device memory";
    default: return "unknown status";
  }
}

int rn_params_default(rn_params* p)
{
  if (!p) return fail(RN_ERR_BAD_PARAM, "params is NULL");
  p->seed = 12345u;           // the reference seeds from time(NULL) (Terrain.cpp:58)
  p->depth = 12;              // Core.cpp:63 (Release)
  p->amplitude = 0.2f;        // Core.cpp:65
  p->size = 500.0f;           // Core.cpp:65
  p->persistence = 0.5f;      // Terrain.cpp:186,231: displacement ~ step/maxWidth
  p->lacunarity = 2;          // Terrain.cpp:81
  p->sideways_scale = 0.5f;   // Terrain.cpp:142
  p->radiosity_strength = 0.5f;  // Terrain.cpp:143
  p->feature_scale = 80.0f;   // Terrain.cpp:169
  p->amplitude_variation = 0.5f;  // Terrain.cpp:170
  p->bulgy_rocks = 1;         // Terrain.cpp:189
  p->apply_volcano = 1;       // Terrain.cpp:88
  return RN_OK;
}

int rn_terrain2d_side(int depth)
{
  if (depth < RN_MIN_DEPTH || depth > RN_MAX_DEPTH) return -1;
  return (1 << depth) + 1;
}

uint64_t rn_terrain2d_draw_count(int depth)
{
  if (depth < RN_MIN_DEPTH || depth > 30) return 0;
  uint64_t total = 0;
  for (int l = 0; l < depth - 2; l++) {
    const uint64_t n = 4ull << l;
    total += 3 * n * n + 2 * n;  // n^2 generateSquare + 2n(n+1) generateDiamond calls (Terrain.cpp:66-80)
  }
  return total;
}

uint64_t rn_terrain2d_call_index(int depth, int i, int j)
{
  if (depth < RN_MIN_DEPTH || depth > 30) return UINT64_MAX;
  const int W = 1 << depth;
  if (i < 0 || j < 0 || i > W || j > W) return UINT64_MAX;
  // The level that creates (i,j) is fixed by the lowest set bit of i|j.
  const unsigned low = (unsigned)(i | j);
  if (low == 0) return UINT64_MAX;
  const int tz = __builtin_ctz(low);  // created at half-step h = 1<<tz
  const int level = depth - 3 - tz;   // s = 2h = W >> (level+2)
  if (level < 0) return UINT64_MAX;   // multiples of W/4: never displaced
  uint64_t base = 0;
  for (int l = 0; l < level; l++) {
    const uint64_t n = 4ull << l;
    base += 3 * n * n + 2 * n;
  }
  const uint64_t n = 4ull << level;
  const int s = 2 << tz;
  const uint64_t qi = (uint64_t)(i / s), qj = (uint64_t)(j / s);
  const bool io = (i % s) != 0, jo = (j % s) != 0;
  if (io && jo) return base + qi * n + qj;                      // generateSquare  (Terrain.cpp:66-68)
  if (!io) return base + n * n + qi * n + qj;                   // generateDiamond loop 1 (:75-77)
  return base + n * n + n * (n + 1) + qi * (n + 1) + qj;        // generateDiamond loop 2 (:78-80)
}

uint32_t rn_lcg_state_after(uint32_t seed, uint64_t k) { return lcg_state_after(seed, k); }
uint32_t rn_lcg_draw(uint32_t seed, uint64_t k) { return lcg_draw_from_state(lcg_state_after(seed, k + 1)); }

int rn_create(int device, rn_context** out_ctx)
{
  if (!out_ctx) return fail(RN_ERR_BAD_PARAM, "out_ctx is NULL");
  *out_ctx = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(RN_ERR_NO_DEVICE, "no CUDA device visible; rednoise_b200 has no CPU fallback");
  }
  if (device < 0) RN_CUDA(cudaGetDevice(&device));
  if (device >= count) return fail(RN_ERR_BAD_PARAM, "device index out of range");
  cudaDeviceProp prop;
  RN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10 || prop.minor != 0)  // arch-specific SASS only (no PTX): another CC 10.x part could not run it
    return fail(RN_ERR_NO_DEVICE, "device is not sm_100 (the library ships sm_100a code only)");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  rn_context* ctx = new (std::nothrow) rn_context();
  if (!ctx) return fail(RN_ERR_OOM, "host allocation failed");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return fail_cuda(e, "cudaStreamCreateWithFlags");
  }
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  e = cudaStreamCreateWithPriority(&ctx->coarse_stream, cudaStreamNonBlocking, prio_hi);
  if (e == cudaSuccess && RN_TILE_STREAM) e = cudaStreamCreateWithPriority(&ctx->tile_stream, cudaStreamNonBlocking, prio_hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; k++) {
    e = cudaEventCreateWithFlags(&ctx->coarse_done[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->tiles_done[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->slot_free[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->chunk_ready[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->chunk_copied[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->seeds_ready[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->seeds_free[k], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    rn_destroy(ctx);
    return fail_cuda(e, "side stream / events");
  }
  *out_ctx = ctx;
  return RN_OK;
}

int rn_destroy(rn_context* ctx)
{
  if (!ctx) return RN_OK;
  DeviceGuard guard(ctx->device);
  if (ctx->coarse_stream) {
    cudaStreamSynchronize(ctx->coarse_stream);
    cudaStreamDestroy(ctx->coarse_stream);
  }
  if (ctx->tile_stream) {
    cudaStreamSynchronize(ctx->tile_stream);
    cudaStreamDestroy(ctx->tile_stream);
  }
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->bounce) cudaFreeHost(ctx->bounce);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
  }
  for (int k = 0; k < 2; k++) {
    if (ctx->coarse_done[k]) cudaEventDestroy(ctx->coarse_done[k]);
    if (ctx->tiles_done[k]) cudaEventDestroy(ctx->tiles_done[k]);
    if (ctx->slot_free[k]) cudaEventDestroy(ctx->slot_free[k]);
    if (ctx->chunk_ready[k]) cudaEventDestroy(ctx->chunk_ready[k]);
    if (ctx->chunk_copied[k]) cudaEventDestroy(ctx->chunk_copied[k]);
    if (ctx->seeds_ready[k]) cudaEventDestroy(ctx->seeds_ready[k]);
    if (ctx->seeds_free[k]) cudaEventDestroy(ctx->seeds_free[k]);
    cudaFree(ctx->seeds[k]);
    cudaFree(ctx->slot[k]);
  }
  for (int k = 0; k < 17; k++)
    if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
  cudaFree(ctx->field);
  cudaFree(ctx->cells);
  cudaFree(ctx->plane);
  cudaFree(ctx->small);
  delete ctx;
  return RN_OK;
}

uint64_t rn_launch_count(const rn_context* ctx) { return ctx ? ctx->launches : 0; }

int rn_set_algorithm(rn_context* ctx, int algo)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  if (algo < 0 || algo > 3)
    return fail(RN_ERR_BAD_PARAM, "algorithm must be 0 (auto), 1 (level-synchronous), 2 (fused tiles + warp march) or 3 (fused tiles + block streaming)");
  ctx->algo = algo;
  return RN_OK;
}

int rn_set_pipelining(rn_context* ctx, int enable)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  ctx->overlap = enable != 0;
  return RN_OK;
}

int rn_set_profiling(rn_context* ctx, int enable)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  ctx->profiling = enable != 0;
  ctx->ev_stages = 0;
  return RN_OK;
}

int rn_last_stage_times(rn_context* ctx, float* ms, int cap)
{
  if (!ctx || !ms || cap <= 0) return -fail(RN_ERR_BAD_PARAM, "ctx/ms is NULL");
  if (!ctx->profiling || ctx->ev_stages == 0) return 0;
  DeviceGuard guard(ctx->device);
  if (cudaEventSynchronize(ctx->ev[ctx->ev_stages]) != cudaSuccess) return -fail(RN_ERR_CUDA, "cudaEventSynchronize failed");
  int n = ctx->ev_stages < cap ? ctx->ev_stages : cap;
  for (int k = 0; k < n; k++)
    if (cudaEventElapsedTime(&ms[k], ctx->ev[k], ctx->ev[k + 1]) != cudaSuccess) return -fail(RN_ERR_CUDA, "cudaEventElapsedTime failed");
  return n;
}

int rn_terrain2d(rn_context* ctx, const rn_params* p, const rn_window* win, rn_element* dst, size_t pitch_elems,
                 int dst_is_device, void* stream_v, float* volcano_pos)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!dst) return fail(RN_ERR_BAD_PARAM, "dst is NULL");
  const int W = 1 << p->depth, side = W + 1;
  rn_window w{0, 0, side, side};
  if (win) w = *win;
  if (w.i0 < 0 || w.j0 < 0 || w.ni <= 0 || w.nj <= 0 || w.i0 > side - w.ni || w.j0 > side - w.nj)
    return fail(RN_ERR_BAD_PARAM, "window does not lie inside the (W+1)x(W+1) lattice");
  if (pitch_elems == 0) pitch_elems = (size_t)w.nj;
  if (pitch_elems < (size_t)w.nj) return fail(RN_ERR_BAD_PARAM, "pitch_elems < nj");
  if (volcano_pos && stream_v && dst_is_device)
    return fail(RN_ERR_BAD_PARAM, "volcano_pos requires a blocking call (stream == NULL or host dst)");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
  const bool blocking = !(stream_v && dst_is_device);

  const bool aligned = dst_is_device && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  const int c = W / 2;
  const bool centre_inside = (c >= w.i0 && c < w.i0 + w.ni && c >= w.j0 && c < w.j0 + w.nj);
  Elem centre{};
  if (ctx->algo == 1) {
    // ---- algorithm 1: level-synchronous, whole field in global memory, then window extraction ----
    const bool whole = (w.i0 == 0 && w.j0 == 0 && w.ni == side && w.nj == side);
    const bool in_place = whole && aligned;
    Elem* g;
    size_t gpitch;
    if (in_place) {
      g = reinterpret_cast<Elem*>(dst);
      gpitch = pitch_elems;
    } else {
      st = ensure(&ctx->field, &ctx->field_bytes, sizeof(Elem) * (size_t)side * (size_t)side);
      if (st != RN_OK) return st;
      g = reinterpret_cast<Elem*>(ctx->field);
      gpitch = (size_t)side;
    }
    st = synth_full_levelsync(ctx, p, g, gpitch, stream);
    if (st != RN_OK) return st;
    if (p->apply_volcano) {
      Elem* wg = g + (size_t)w.i0 * gpitch + (size_t)w.j0;
      RN_CUDA(volcano_apply(wg, gpitch, p->depth, w.i0, w.j0, w.ni, w.nj, p->amplitude, stream, &ctx->launches));
    }
    if (!in_place) {
      const Elem* src = g + (size_t)w.i0 * gpitch + (size_t)w.j0;
      RN_CUDA(cudaMemcpy2DAsync(dst, pitch_elems * sizeof(Elem), src, gpitch * sizeof(Elem), (size_t)w.nj * sizeof(Elem), (size_t)w.ni,
                                dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    }
    if (volcano_pos) {
      RN_CUDA(cudaMemcpyAsync(&centre, g + (size_t)c * gpitch + (size_t)c, sizeof(Elem), cudaMemcpyDeviceToHost, stream));
      RN_CUDA(cudaStreamSynchronize(stream));
      if (p->apply_volcano && !centre_inside) centre.y += volcano_height(c, c, W, p->amplitude);
    }
  } else {
    // ---- algorithm 2 (default): fused tiles write the window directly ----
    Elem* g = nullptr;
    size_t gpitch = 0;
    bool chunked = false;
    if (aligned) {
      g = reinterpret_cast<Elem*>(dst);
      gpitch = pitch_elems;
      st = synth_window_fused(ctx, p, w, g, gpitch, 0, nullptr, 1, stream);
      if (st != RN_OK) return st;
    } else if (!dst_is_device && ctx->copy_stream && (size_t)w.ni * (size_t)w.nj >= ((size_t)1 << 22)) {
      chunked = true;
      st = enqueue_window_to_host(ctx, p, w, dst, pitch_elems, stream);
      if (st != RN_OK) return st;
      RN_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    } else {
      st = ensure(&ctx->field, &ctx->field_bytes, sizeof(Elem) * (size_t)w.ni * (size_t)w.nj);
      if (st != RN_OK) return st;
      g = reinterpret_cast<Elem*>(ctx->field);
      gpitch = (size_t)w.nj;
      st = synth_window_fused(ctx, p, w, g, gpitch, 0, nullptr, 1, stream);
      if (st != RN_OK) return st;
      RN_CUDA(cudaMemcpy2DAsync(dst, pitch_elems * sizeof(Elem), g, gpitch * sizeof(Elem), (size_t)w.nj * sizeof(Elem), (size_t)w.ni,
                                dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    }
    if (volcano_pos && chunked && centre_inside) {
      // the destination is host memory and complete: read the centre sample from it
      const rn_element& e = dst[(size_t)(c - w.i0) * pitch_elems + (size_t)(c - w.j0)];
      centre = Elem{e.x, e.y, e.z, e.shade};
    } else
    if (volcano_pos) {
      if (centre_inside) {
        RN_CUDA(cudaMemcpyAsync(&centre, g + (size_t)(c - w.i0) * gpitch + (size_t)(c - w.j0), sizeof(Elem), cudaMemcpyDeviceToHost, stream));
      } else {
        // the centre sample lies outside the window: synthesize it alone (a handful of tiles)
        st = ensure(&ctx->small, &ctx->small_bytes, 256);
        if (st != RN_OK) return st;
        const rn_window cw{c, c, 1, 1};
        st = synth_window_fused(ctx, p, cw, reinterpret_cast<Elem*>(ctx->small), 1, 0, nullptr, 1, stream);
        if (st != RN_OK) return st;
        RN_CUDA(cudaMemcpyAsync(&centre, ctx->small, sizeof(Elem), cudaMemcpyDeviceToHost, stream));
      }
      RN_CUDA(cudaStreamSynchronize(stream));
    }
  }
  if (volcano_pos) {
    // Terrain.cpp:140 + :89: volcanoPos = baseArray[W/2][W/2].pos * size
    volcano_pos[0] = centre.x * p->size;
    volcano_pos[1] = centre.y * p->size;
    volcano_pos[2] = centre.z * p->size;
  }
  if (blocking) RN_CUDA(cudaStreamSynchronize(stream));
  return RN_OK;
}

int rn_terrain2d_batch(rn_context* ctx, const rn_params* p, const uint32_t* seeds, int ntiles, rn_element* dst,
                       size_t tile_stride_elems, int dst_is_device, void* stream_v)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!seeds || !dst) return fail(RN_ERR_BAD_PARAM, "seeds/dst is NULL");
  if (ntiles < 0) return fail(RN_ERR_BAD_PARAM, "ntiles < 0");
  const size_t side = (size_t)(1 << p->depth) + 1;
  if (tile_stride_elems == 0) tile_stride_elems = side * side;
  if (tile_stride_elems < side * side) return fail(RN_ERR_BAD_PARAM, "tile_stride_elems < side*side");
  if (ntiles == 0) return RN_OK;
  if (ctx->algo == 1) {
    for (int t = 0; t < ntiles; t++) {
      rn_params q = *p;
      q.seed = seeds[t];
      st = rn_terrain2d(ctx, &q, nullptr, dst + (size_t)t * tile_stride_elems, side, dst_is_device, stream_v, nullptr);
      if (st != RN_OK) return st;
    }
    return RN_OK;
  }
  // algorithm 2: one launch chain per chunk of fields, blockIdx.z = field
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
  const bool direct = dst_is_device && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  const size_t field_bytes = side * side * sizeof(Elem);
  const size_t chunk = batch_chunk(ctx, p, ntiles, direct);
  const int sb = (int)(ctx->batch_seq++ & 1);
  st = ensure(&ctx->seeds[sb], &ctx->seeds_bytes[sb], sizeof(uint32_t) * (size_t)ntiles);
  if (st != RN_OK) return st;
  if (ctx->seeds_used[sb]) RN_CUDA(cudaStreamWaitEvent(stream, ctx->seeds_free[sb], 0));
  RN_CUDA(cudaMemcpyAsync(ctx->seeds[sb], seeds, sizeof(uint32_t) * (size_t)ntiles, cudaMemcpyHostToDevice, stream));
  RN_CUDA(cudaEventRecord(ctx->seeds_ready[sb], stream));
  const rn_window w{0, 0, (int)side, (int)side};
  for (size_t t0 = 0; t0 < (size_t)ntiles; t0 += chunk) {
    const int nf = (int)(((size_t)ntiles - t0 < chunk) ? (size_t)ntiles - t0 : chunk);
    const uint32_t* dseeds = reinterpret_cast<const uint32_t*>(ctx->seeds[sb]) + t0;
    if (direct) {
      st = synth_window_fused(ctx, p, w, reinterpret_cast<Elem*>(dst + t0 * tile_stride_elems), side, tile_stride_elems, dseeds, nf, stream,
                              ctx->seeds_ready[sb]);
      if (st != RN_OK) return st;
    } else {
      st = ensure(&ctx->field, &ctx->field_bytes, field_bytes * (size_t)nf);
      if (st != RN_OK) return st;
      st = synth_window_fused(ctx, p, w, reinterpret_cast<Elem*>(ctx->field), side, side * side, dseeds, nf, stream, ctx->seeds_ready[sb]);
      if (st != RN_OK) return st;
      RN_CUDA(cudaMemcpy2DAsync(dst + t0 * tile_stride_elems, tile_stride_elems * sizeof(Elem), ctx->field, field_bytes, field_bytes, (size_t)nf,
                                dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    }
  }
  RN_CUDA(cudaEventRecord(ctx->seeds_free[sb], stream));  // every stage of the call precedes its output stage on `stream`
  ctx->seeds_used[sb] = true;
  if (!(stream_v && dst_is_device)) RN_CUDA(cudaStreamSynchronize(stream));
  return RN_OK;
}

int rn_terrain2d_batch_chunk(rn_context* ctx, const rn_params* p, int ntiles)
{
  if (!ctx || check_params(p) != RN_OK || ntiles < 1) return 0;
  const size_t c = batch_chunk(ctx, p, ntiles, true);
  return (int)(c < (size_t)ntiles ? c : (size_t)ntiles);
}

uint64_t rn_terrain2d_cell_count(int depth)
{
  if (depth < 1 || depth > 31) return 0;
  return ((1ull << (2 * depth)) - 1) / 3;
}

uint64_t rn_terrain2d_cell_level_offset(int level)
{
  if (level < 0 || level > 31) return 0;
  return ((1ull << (2 * level)) - 1) / 3;
}

int rn_terrain2d_cells(rn_context* ctx, const rn_params* p, const rn_element* field, size_t field_pitch, int field_is_device,
                       rn_cell* cells, int cells_is_device, void* stream_v)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!cells) return fail(RN_ERR_BAD_PARAM, "cells is NULL");
  const size_t side = ((size_t)1 << p->depth) + 1;
  if (field_pitch == 0) field_pitch = side;
  if (field_pitch < side) return fail(RN_ERR_BAD_PARAM, "field_pitch < side");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
  const Elem* g = nullptr;
  size_t gpitch = field_pitch;
  if (field && field_is_device && reinterpret_cast<uintptr_t>(field) % 16 == 0) {
    g = reinterpret_cast<const Elem*>(field);
  } else {
    st = ensure(&ctx->field, &ctx->field_bytes, sizeof(Elem) * side * side);
    if (st != RN_OK) return st;
    gpitch = side;
    if (field) {
      RN_CUDA(cudaMemcpy2DAsync(ctx->field, side * sizeof(Elem), field, field_pitch * sizeof(Elem), side * sizeof(Elem), side,
                                field_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    } else {  // synthesize the field first (Terrain.cpp:44-88), whole window, into the staging buffer
      st = rn_terrain2d(ctx, p, nullptr, reinterpret_cast<rn_element*>(ctx->field), side, 1, stream, nullptr);
      if (st != RN_OK) return st;
    }
    g = reinterpret_cast<const Elem*>(ctx->field);
  }
  const uint64_t n = rn_terrain2d_cell_count(p->depth);
  void* out = cells;
  const bool direct = cells_is_device && reinterpret_cast<uintptr_t>(cells) % 16 == 0;
  if (!direct) {
    st = ensure(&ctx->cells, &ctx->cells_bytes, n * sizeof(rn_cell));
    if (st != RN_OK) return st;
    out = ctx->cells;
  }
  RN_CUDA(cells_generate(out, g, gpitch, p->depth, p->feature_scale, p->size, p->amplitude, stream, &ctx->launches));
  if (!direct)
    RN_CUDA(cudaMemcpyAsync(cells, out, n * sizeof(rn_cell), cells_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
  if (!(stream_v && cells_is_device)) RN_CUDA(cudaStreamSynchronize(stream));
  return RN_OK;
}

int rn_volume3d(rn_context* ctx, const rn_volume_params* p, int i0, int ni, float* slab, void* stream_v)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  if (!p || !slab) return fail(RN_ERR_BAD_PARAM, "params/slab is NULL");
  if (p->depth < RN_VOLUME_MIN_DEPTH || p->depth > RN_VOLUME_MAX_DEPTH) return fail(RN_ERR_BAD_PARAM, "volume depth out of range [2,11]");
  const int side = (1 << p->depth) + 1;
  if (i0 < 0 || ni <= 0 || i0 > side - ni) return fail(RN_ERR_BAD_PARAM, "slab does not lie inside the volume");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
  // The coarser level grids of the slab's dependency cone live in one of the two context-owned scratch slots (grow-only, shared
  // with the 2D path and ordered by the same events).  Call pipelining as in synth_window_fused: every level but the last runs on
  // the high-priority side stream, so the latency-bound coarse levels of this call overlap the last level of the call before.
  const int sl = (int)(ctx->seq & 1);
  ctx->seq++;
  int st = ensure(&ctx->slot[sl], &ctx->slot_bytes[sl], volume_scratch_bytes(p->depth, i0, ni));
  if (st != RN_OK) return st;
  if (ctx->profiling && !ctx->ev[0])
    for (int k = 0; k < 17; k++) RN_CUDA(cudaEventCreate(&ctx->ev[k]));
  ctx->ev_stages = 0;
  // (calls above ~1025^3 samples are throughput-bound at every level that matters: two of them side by side only compete --
  // 2049^3 measured 475 Gsamples/s serial, 447-464 overlapped -- so they stay on the caller's stream)
  const bool overlap = ctx->overlap && !ctx->profiling && ctx->coarse_stream && (uint64_t)ni * (uint64_t)side * (uint64_t)side <= 1200000000ull;
  if (ctx->slot_used[sl]) RN_CUDA(cudaStreamWaitEvent(overlap ? ctx->coarse_stream : stream, ctx->slot_free[sl], 0));
  RN_CUDA(volume_generate(slab, ctx->slot[sl], p->depth, p->seed, p->amplitude, i0, ni, stream, &ctx->launches, ctx->profiling ? ctx->ev : nullptr,
                          ctx->profiling ? &ctx->ev_stages : nullptr, overlap ? ctx->coarse_stream : nullptr, overlap ? ctx->coarse_done[sl] : nullptr));
  RN_CUDA(cudaEventRecord(ctx->slot_free[sl], stream));
  ctx->slot_used[sl] = true;
  if (!stream_v) RN_CUDA(cudaStreamSynchronize(stream));
  return RN_OK;
}

int rn_volume3d_host(rn_context* ctx, const rn_volume_params* p, int i0, int ni, float* host_slab)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  if (!p || !host_slab) return fail(RN_ERR_BAD_PARAM, "params/host_slab is NULL");
  if (p->depth < RN_VOLUME_MIN_DEPTH || p->depth > RN_VOLUME_MAX_DEPTH) return fail(RN_ERR_BAD_PARAM, "volume depth out of range [2,11]");
  const int side = (1 << p->depth) + 1;
  if (i0 < 0 || ni <= 0 || i0 > side - ni) return fail(RN_ERR_BAD_PARAM, "slab does not lie inside the volume");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  const size_t plane = (size_t)side * (size_t)side;
  int planes = (int)(((size_t)64 << 20) / (plane * sizeof(float)));  // ~64 MB per chunk, at least 4 planes (a level reaches 3 planes beyond a slab)
  planes = planes < 4 ? 4 : planes;
  planes = planes > ni ? ni : planes;
  const size_t half = (size_t)planes * plane;
  int st = ensure(&ctx->field, &ctx->field_bytes, sizeof(float) * 2 * half);
  if (st != RN_OK) return st;
  size_t scratch = 0;  // the largest cone of any chunk, allocated once (the chunks follow each other on one stream)
  for (int r0 = 0; r0 < ni; r0 += planes) {
    const size_t b = volume_scratch_bytes(p->depth, i0 + r0, ni - r0 < planes ? ni - r0 : planes);
    scratch = b > scratch ? b : scratch;
  }
  st = ensure(&ctx->plane, &ctx->plane_bytes, scratch);
  if (st != RN_OK) return st;
  cudaStream_t stream = ctx->stream;
  int k = 0;
  for (int r0 = 0; r0 < ni; r0 += planes, k++) {
    const int nr = ni - r0 < planes ? ni - r0 : planes;
    const int h = k & 1;
    float* stage = reinterpret_cast<float*>(ctx->field) + (size_t)h * half;
    if (k >= 2) RN_CUDA(cudaStreamWaitEvent(stream, ctx->chunk_copied[h], 0));  // the staging half is free again
    RN_CUDA(volume_generate(stage, ctx->plane, p->depth, p->seed, p->amplitude, i0 + r0, nr, stream, &ctx->launches));
    RN_CUDA(cudaEventRecord(ctx->chunk_ready[h], stream));
    if (ctx->copy_stream) {
      RN_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ready[h], 0));
      RN_CUDA(cudaMemcpyAsync(host_slab + (size_t)r0 * plane, stage, (size_t)nr * plane * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
      RN_CUDA(cudaEventRecord(ctx->chunk_copied[h], ctx->copy_stream));
    } else {
      RN_CUDA(cudaMemcpyAsync(host_slab + (size_t)r0 * plane, stage, (size_t)nr * plane * sizeof(float), cudaMemcpyDeviceToHost, stream));
      RN_CUDA(cudaEventRecord(ctx->chunk_copied[h], stream));
    }
  }
  if (ctx->copy_stream) RN_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  RN_CUDA(cudaStreamSynchronize(stream));
  return RN_OK;
}

uint64_t rn_volume3d_slab_elems(int depth, int ni)
{
  if (depth < RN_VOLUME_MIN_DEPTH || depth > RN_VOLUME_MAX_DEPTH || ni <= 0) return 0;
  const uint64_t side = ((uint64_t)1 << depth) + 1;
  return (uint64_t)ni * side * side;
}

uint64_t rn_rock_draw_count(int depth)
{
  if (depth < RN_ROCK_MIN_DEPTH || depth > RN_ROCK_MAX_DEPTH) return 0;
  return rock_draw_count(depth);
}

namespace {
int rock_or_cloud(rn_context* ctx, int kind, const rn_rock_params* p, rn_rock_element* faces, rn_cell* cells, int dst_is_device, void* stream_v,
                  uint32_t* lcg_state_after)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  if (!p) return fail(RN_ERR_BAD_PARAM, "params is NULL");
  const int nfaces = kind ? 2 : 6, min_depth = kind ? RN_CLOUD_MIN_DEPTH : RN_ROCK_MIN_DEPTH;
  if (p->depth < min_depth || p->depth > RN_ROCK_MAX_DEPTH) return fail(RN_ERR_BAD_PARAM, kind ? "cloud depth out of range [2,12]" : "rock depth out of range [1,12]");
  if (!faces && !cells) return fail(RN_ERR_BAD_PARAM, "faces and cells are both NULL");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
  const size_t side = ((size_t)1 << p->depth) + 1;
  const size_t face_bytes = (size_t)nfaces * side * side * sizeof(rn_rock_element);
  const size_t ncells = (size_t)nfaces * (size_t)rn_terrain2d_cell_count(p->depth);
  int st = ensure(&ctx->small, &ctx->small_bytes, 256);
  if (st != RN_OK) return st;
  void* dfaces = faces;
  if (!faces || !dst_is_device || reinterpret_cast<uintptr_t>(faces) % 8 != 0) {
    st = ensure(&ctx->field, &ctx->field_bytes, face_bytes);
    if (st != RN_OK) return st;
    dfaces = ctx->field;
  }
  void* dcells = cells;
  if (cells && (!dst_is_device || reinterpret_cast<uintptr_t>(cells) % 16 != 0)) {
    st = ensure(&ctx->cells, &ctx->cells_bytes, ncells * sizeof(rn_cell));
    if (st != RN_OK) return st;
    dcells = ctx->cells;
  }
  uint64_t draws = kind ? cloud_draw_count(p->depth) : rock_draw_count(p->depth);
  RN_CUDA(rock_generate(dfaces, dcells, (int*)ctx->small, kind, p->depth, p->lcg_state, p->amplitude, 0, stream, &ctx->launches));
  int collision = 0;
  RN_CUDA(cudaMemcpyAsync(&collision, ctx->small, sizeof collision, cudaMemcpyDeviceToHost, stream));
  RN_CUDA(cudaStreamSynchronize(stream));
  if (collision) {
    // a displaced height is exactly 0.0f: the reference's `height == 0.0f` test (Rock.cpp:181, Cloud.cpp:134) draws again and the
    // call sequence becomes value-dependent -- regenerate with the literal one-thread recursion on the device
    RN_CUDA(rock_generate(dfaces, dcells, (int*)ctx->small, kind, p->depth, p->lcg_state, p->amplitude, 1, stream, &ctx->launches));
    unsigned long long n = 0;
    RN_CUDA(cudaMemcpyAsync(&n, ctx->small, sizeof n, cudaMemcpyDeviceToHost, stream));
    RN_CUDA(cudaStreamSynchronize(stream));
    draws = n;
  }
  if (faces && dfaces != faces)
    RN_CUDA(cudaMemcpyAsync(faces, dfaces, face_bytes, dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
  if (cells && dcells != cells)
    RN_CUDA(cudaMemcpyAsync(cells, dcells, ncells * sizeof(rn_cell), dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
  RN_CUDA(cudaStreamSynchronize(stream));
  if (lcg_state_after) *lcg_state_after = rn_lcg_state_after(p->lcg_state, draws);
  return RN_OK;
}
}  // namespace

int rn_rock(rn_context* ctx, const rn_rock_params* p, rn_rock_element* faces, rn_cell* cells, int dst_is_device, void* stream_v,
            uint32_t* lcg_state_after)
{
  return rock_or_cloud(ctx, 0, p, faces, cells, dst_is_device, stream_v, lcg_state_after);
}

int rn_rock_batch(rn_context* ctx, uint32_t lcg_state, int nrocks, const int32_t* depths, float amplitude, rn_rock_element* const* faces,
                  rn_cell* const* cells, int dst_is_device, void* stream_v, int32_t* depths_out, float* exponents_out, uint32_t* states_out,
                  uint32_t* lcg_state_after)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  if (nrocks < 0 || nrocks > RN_ROCK_BATCH_MAX) return fail(RN_ERR_BAD_PARAM, "nrocks outside [0, RN_ROCK_BATCH_MAX]");
  if (!faces && !cells) return fail(RN_ERR_BAD_PARAM, "faces and cells are both NULL");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
  // 1. the host walks the rand() stream: one call decides a rock's depth (Core.cpp:72-75), the rock then consumes
  //    rn_rock_draw_count(depth) calls -- a closed form, so every rock's entry state is known before anything is generated
  std::vector<int> d((size_t)nrocks);
  std::vector<uint32_t> st((size_t)nrocks + 1);
  uint32_t s = lcg_state;
  for (int i = 0; i < nrocks; i++) {
    float exponent = 0.0f;
    if (depths) {
      d[i] = depths[i];
    } else {
      s = s * kLcgA + kLcgC;
      const int r = (int)((s >> 16) & 0x7fffu);
      exponent = -5.0f + (0.0f - -5.0f) * ((float)r / 32767.0f);  // random(-5.0f, 0.0f), basics.h:20-23
      d[i] = (int)(exponent + 7);                                   // Core.cpp:75
    }
    if (d[i] < RN_ROCK_MIN_DEPTH || d[i] > RN_ROCK_MAX_DEPTH) return fail(RN_ERR_BAD_PARAM, "rock depth out of range [1,12]");
    if (exponents_out) exponents_out[i] = exponent;
    if (depths_out) depths_out[i] = d[i];
    st[i] = s;
    s = rn::lcg_state_after(s, rock_draw_count(d[i]));
  }
  st[nrocks] = s;
  if (nrocks == 0) {
    if (lcg_state_after) *lcg_state_after = s;
    return RN_OK;
  }
  // 2. destinations: the caller's device buffers, or device staging for host destinations
  std::vector<void*> df((size_t)nrocks, nullptr), dc((size_t)nrocks, nullptr);
  std::vector<size_t> fbytes((size_t)nrocks), cbytes((size_t)nrocks);
  size_t ftotal = 0, ctotal = 0;
  for (int i = 0; i < nrocks; i++) {
    const size_t side = ((size_t)1 << d[i]) + 1;
    fbytes[i] = ((6 * side * side * sizeof(rn_rock_element) + 255) / 256) * 256;
    cbytes[i] = ((6 * (size_t)rn_terrain2d_cell_count(d[i]) * sizeof(rn_cell) + 255) / 256) * 256;
    ftotal += fbytes[i];
    ctotal += cbytes[i];
  }
  const bool stage_f = faces && !dst_is_device, stage_c = cells && !dst_is_device;
  int stt = RN_OK;
  if (stage_f && (stt = ensure(&ctx->field, &ctx->field_bytes, ftotal)) != RN_OK) return stt;
  if (stage_c && (stt = ensure(&ctx->cells, &ctx->cells_bytes, ctotal)) != RN_OK) return stt;
  {
    size_t fo = 0, co = 0;
    for (int i = 0; i < nrocks; i++) {
      if (faces) {
        if (!faces[i]) return fail(RN_ERR_BAD_PARAM, "faces[i] is NULL");
        df[i] = stage_f ? static_cast<char*>(ctx->field) + fo : (void*)faces[i];
      }
      if (cells) {
        if (!cells[i]) return fail(RN_ERR_BAD_PARAM, "cells[i] is NULL");
        dc[i] = stage_c ? static_cast<char*>(ctx->cells) + co : (void*)cells[i];
      }
      fo += fbytes[i];
      co += cbytes[i];
    }
  }
  // 3. rocks of depth <= 7 (all of the app's): one launch for the lot; deeper ones through the level kernels
  const size_t item = rock_small_item_bytes();
  if ((stt = ensure(&ctx->small, &ctx->small_bytes, 256 + (size_t)nrocks * (item + 16))) != RN_OK) return stt;
  int* flags = static_cast<int*>(ctx->small);
  void* items_dev = static_cast<char*>(ctx->small) + ((sizeof(int) * (size_t)nrocks + 255) / 256) * 256;
  std::vector<char> items_host((size_t)nrocks * item);
  std::vector<int> small_idx, sd;
  std::vector<uint32_t> ss;
  std::vector<void*> sf, sc;
  for (int i = 0; i < nrocks; i++)
    if (d[i] <= 7 && (!df[i] || reinterpret_cast<uintptr_t>(df[i]) % 8 == 0) && (!dc[i] || reinterpret_cast<uintptr_t>(dc[i]) % 16 == 0)) {
      small_idx.push_back(i); sd.push_back(d[i]); ss.push_back(st[i]); sf.push_back(df[i]); sc.push_back(dc[i]);
    }
  if (!small_idx.empty())
    RN_CUDA(rock_small_batch((int)small_idx.size(), sd.data(), ss.data(), amplitude, faces ? sf.data() : nullptr, cells ? sc.data() : nullptr, flags, items_dev,
                             items_host.data(), stream, &ctx->launches));
  std::vector<int> collided(small_idx.size(), 0);
  if (!small_idx.empty()) {
    RN_CUDA(cudaMemcpyAsync(collided.data(), flags, sizeof(int) * small_idx.size(), cudaMemcpyDeviceToHost, stream));
    RN_CUDA(cudaStreamSynchronize(stream));  // (also makes items_host reusable)
  }
  // 4. a displaced height of exactly 0.0f makes the reference's call sequence value-dependent (about 1 rock in 500): from the
  //    first such rock on, the rocks are generated one by one (rn_rock regenerates it literally and reports the state after it)
  int first_bad = nrocks;
  for (size_t q = 0; q < small_idx.size(); q++)
    if (collided[q] && small_idx[q] < first_bad) first_bad = small_idx[q];
  for (int i = 0; i < nrocks; i++) {
    const bool was_small = d[i] <= 7 && std::find(small_idx.begin(), small_idx.end(), i) != small_idx.end();
    if (was_small && i < first_bad) continue;
    if (i >= first_bad && states_out) states_out[i] = st[i];
    rn_rock_params rp{st[i], d[i], amplitude};
    uint32_t after = 0;
    stt = rock_or_cloud(ctx, 0, &rp, reinterpret_cast<rn_rock_element*>(df[i]), reinterpret_cast<rn_cell*>(dc[i]), 1, stream, &after);
    if (stt != RN_OK) return stt;
    if (i >= first_bad && after != st[i + 1]) {  // the stream moved: re-derive everything behind this rock
      uint32_t s2 = after;
      for (int j = i + 1; j < nrocks; j++) {
        float exponent = 0.0f;
        if (!depths) {
          s2 = s2 * kLcgA + kLcgC;
          const int r = (int)((s2 >> 16) & 0x7fffu);
          exponent = -5.0f + (0.0f - -5.0f) * ((float)r / 32767.0f);
          const int nd = (int)(exponent + 7);
          if (nd != d[j]) return fail(RN_ERR_UNSUPPORTED, "a colliding rock changed the depth of a later rock: generate the rocks one by one with rn_rock");
          if (exponents_out) exponents_out[j] = exponent;
        }
        st[j] = s2;
        s2 = rn::lcg_state_after(s2, rock_draw_count(d[j]));
      }
      st[nrocks] = s2;
    }
  }
  if (states_out)
    for (int i = 0; i < nrocks; i++) states_out[i] = st[i];
  // 5. host destinations
  for (int i = 0; i < nrocks; i++) {
    const size_t side = ((size_t)1 << d[i]) + 1;
    if (stage_f) RN_CUDA(cudaMemcpyAsync(faces[i], df[i], 6 * side * side * sizeof(rn_rock_element), cudaMemcpyDeviceToHost, stream));
    if (stage_c) RN_CUDA(cudaMemcpyAsync(cells[i], dc[i], 6 * (size_t)rn_terrain2d_cell_count(d[i]) * sizeof(rn_cell), cudaMemcpyDeviceToHost, stream));
  }
  RN_CUDA(cudaStreamSynchronize(stream));
  if (lcg_state_after) *lcg_state_after = st[nrocks];
  return RN_OK;
}

uint64_t rn_cloud_draw_count(int depth)
{
  if (depth < RN_CLOUD_MIN_DEPTH || depth > RN_ROCK_MAX_DEPTH) return 0;
  return cloud_draw_count(depth);
}

int rn_cloud(rn_context* ctx, const rn_rock_params* p, rn_rock_element* faces, rn_cell* cells, int dst_is_device, void* stream_v,
             uint32_t* lcg_state_after)
{
  return rock_or_cloud(ctx, 1, p, faces, cells, dst_is_device, stream_v, lcg_state_after);
}

int rn_selftest_math(rn_context* ctx, uint64_t* mismatches3)
{
  if (!ctx || !mismatches3) return fail(RN_ERR_BAD_PARAM, "ctx/mismatches3 is NULL");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  int st = ensure(&ctx->small, &ctx->small_bytes, 256);
  if (st != RN_OK) return st;
  RN_CUDA(selftest_math((unsigned long long*)ctx->small, ctx->stream, &ctx->launches));
  unsigned long long h[3];
  RN_CUDA(cudaMemcpyAsync(h, ctx->small, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 3; k++) mismatches3[k] = h[k];
  return RN_OK;
}

int rn_lcg_draws_device(rn_context* ctx, uint32_t seed, uint64_t first, uint64_t n, uint16_t* dst, int dst_is_device, void* stream_v)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  if (!dst && n) return fail(RN_ERR_BAD_PARAM, "dst is NULL");
  if (n == 0) return RN_OK;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
  if (dst_is_device) {
    RN_CUDA(lcg_draws(dst, seed, first, n, stream, &ctx->launches));
    if (!stream_v) RN_CUDA(cudaStreamSynchronize(stream));
    return RN_OK;
  }
  int st = ensure(&ctx->small, &ctx->small_bytes, n * sizeof(uint16_t));
  if (st != RN_OK) return st;
  RN_CUDA(lcg_draws((uint16_t*)ctx->small, seed, first, n, stream, &ctx->launches));
  RN_CUDA(cudaMemcpyAsync(dst, ctx->small, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, stream));
  RN_CUDA(cudaStreamSynchronize(stream));
  return RN_OK;
}

// ---------------------------------------------------------------- memory helpers ----
int rn_device_count(void)
{
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

int rn_device_alloc(rn_context* ctx, size_t bytes, void** out)
{
  if (!ctx || !out) return fail(RN_ERR_BAD_PARAM, "ctx/out is NULL");
  *out = nullptr;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  RN_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return RN_OK;
}

int rn_device_free(rn_context* ctx, void* ptr)
{
  if (!ctx) return fail(RN_ERR_BAD_PARAM, "ctx is NULL");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  RN_CUDA(cudaFree(ptr));
  return RN_OK;
}

int rn_host_alloc(size_t bytes, void** out)
{
  if (!out) return fail(RN_ERR_BAD_PARAM, "out is NULL");
  *out = nullptr;
  RN_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return RN_OK;
}

int rn_host_free(void* ptr)
{
  RN_CUDA(cudaFreeHost(ptr));
  return RN_OK;
}

int rn_copy_to_host(rn_context* ctx, void* dst_host, const void* src_device, size_t bytes)
{
  if (!ctx || (!dst_host && bytes) || (!src_device && bytes)) return fail(RN_ERR_BAD_PARAM, "ctx/dst/src is NULL");
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
  RN_CUDA(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  return RN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- several GPUs, one host thread ----
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: libnccl.so.2 is bound at run time (a host process such as PyTorch brings its own copy)

struct rn_multi {
  int n = 0;
  int devices[RN_MAX_DEVICES] = {};
  rn_context* ctx[RN_MAX_DEVICES] = {};
  bool distinct = true;             // no device listed twice (NCCL needs that)
  // NCCL, bound lazily
  void* nccl_lib = nullptr;
  bool nccl_tried = false;
  ncclComm_t comms[RN_MAX_DEVICES] = {};
  bool comms_ready = false;
  ncclResult_t (*pCommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*pCommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*pGroupStart)() = nullptr;
  ncclResult_t (*pGroupEnd)() = nullptr;
  ncclResult_t (*pAllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*pGetErrorString)(ncclResult_t) = nullptr;
};

namespace {

int bind_nccl(rn_multi* m)
{
  if (m->nccl_tried) return m->nccl_lib ? RN_OK : fail(RN_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  m->nccl_tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already carries, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail(RN_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  *(void**)&m->pCommInitAll = dlsym(h, "ncclCommInitAll");
  *(void**)&m->pCommDestroy = dlsym(h, "ncclCommDestroy");
  *(void**)&m->pGroupStart = dlsym(h, "ncclGroupStart");
  *(void**)&m->pGroupEnd = dlsym(h, "ncclGroupEnd");
  *(void**)&m->pAllGather = dlsym(h, "ncclAllGather");
  *(void**)&m->pGetErrorString = dlsym(h, "ncclGetErrorString");
  if (!m->pCommInitAll || !m->pCommDestroy || !m->pGroupStart || !m->pGroupEnd || !m->pAllGather) {
    dlclose(h);
    return fail(RN_ERR_UNSUPPORTED, "libnccl.so.2 lacks a required symbol");
  }
  m->nccl_lib = h;
  return RN_OK;
}

int fail_nccl(rn_multi* m, ncclResult_t r, const char* where)
{
  char buf[512];
  snprintf(buf, sizeof buf, "%s: %s", where, m->pGetErrorString ? m->pGetErrorString(r) : "NCCL error");
  g_last_error = buf;
  return RN_ERR_CUDA;
}

int sync_all(rn_multi* m, bool copy_streams)
{
  int status = RN_OK;
  for (int k = 0; k < m->n; k++) {
    DeviceGuard guard(m->ctx[k]->device);
    cudaError_t e = cudaStreamSynchronize(m->ctx[k]->stream);
    if (e == cudaSuccess && copy_streams) e = cudaStreamSynchronize(m->ctx[k]->copy_stream);
    if (e != cudaSuccess && status == RN_OK) status = fail_cuda(e, "cudaStreamSynchronize (multi)");
  }
  return status;
}

}  // namespace

extern "C" {

int rn_create_multi(const int* devices, int n, rn_multi** out_multi)
{
  if (!out_multi) return fail(RN_ERR_BAD_PARAM, "out_multi is NULL");
  *out_multi = nullptr;
  if (!devices || n < 1 || n > RN_MAX_DEVICES) return fail(RN_ERR_BAD_PARAM, "devices is NULL or n outside [1, RN_MAX_DEVICES]");
  rn_multi* m = new (std::nothrow) rn_multi();
  if (!m) return fail(RN_ERR_OOM, "host allocation failed");
  m->n = n;
  for (int k = 0; k < n; k++) {
    m->devices[k] = devices[k];
    for (int q = 0; q < k; q++)
      if (devices[q] == devices[k]) m->distinct = false;
    const int st = rn_create(devices[k], &m->ctx[k]);
    if (st != RN_OK) {
      const std::string keep = g_last_error;
      rn_destroy_multi(m);
      g_last_error = keep;
      return st;
    }
    m->devices[k] = m->ctx[k]->device;
  }
  // peer access for the copy form of rn_gather (best effort: NVLink-connected B200s allow it)
  for (int k = 0; k < n; k++)
    for (int q = 0; q < n; q++) {
      if (m->devices[k] == m->devices[q]) continue;
      DeviceGuard guard(m->devices[k]);
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, m->devices[k], m->devices[q]) == cudaSuccess && can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[q], 0);
        if (e != cudaSuccess) cudaGetLastError();  // already enabled is fine
      }
    }
  *out_multi = m;
  return RN_OK;
}

int rn_destroy_multi(rn_multi* m)
{
  if (!m) return RN_OK;
  if (m->comms_ready)
    for (int k = 0; k < m->n; k++)
      if (m->comms[k]) m->pCommDestroy(m->comms[k]);
  for (int k = 0; k < m->n; k++) rn_destroy(m->ctx[k]);
  // (the NCCL library stays loaded: other users of the process may hold it)
  delete m;
  return RN_OK;
}

int rn_multi_size(const rn_multi* m) { return m ? m->n : 0; }
rn_context* rn_multi_context(rn_multi* m, int k) { return (m && k >= 0 && k < m->n) ? m->ctx[k] : nullptr; }

int rn_slab_rows(int side, int nparts)
{
  if (side <= 0 || nparts <= 0) return 0;
  return (side + nparts - 1) / nparts;
}

int rn_slab_window(int side, int part, int nparts, rn_window* out)
{
  if (!out || side <= 0 || nparts <= 0 || part < 0 || part >= nparts) return fail(RN_ERR_BAD_PARAM, "bad slab arguments");
  const int rows = rn_slab_rows(side, nparts);
  const int i0 = part * rows < side ? part * rows : side;
  const int ni = rows < side - i0 ? rows : side - i0;
  out->i0 = i0; out->j0 = 0; out->ni = ni; out->nj = side;
  return RN_OK;
}

int rn_terrain2d_sharded(rn_multi* m, const rn_params* p, rn_element* const* slab_dst, size_t slab_pitch_elems, rn_element* host_dst,
                         size_t host_pitch_elems)
{
  if (!m) return fail(RN_ERR_BAD_PARAM, "multi is NULL");
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!slab_dst && !host_dst) return fail(RN_ERR_BAD_PARAM, "slab_dst and host_dst are both NULL");
  const int side = (1 << p->depth) + 1;
  if (slab_pitch_elems == 0) slab_pitch_elems = (size_t)side;
  if (host_pitch_elems == 0) host_pitch_elems = (size_t)side;
  if (slab_pitch_elems < (size_t)side || host_pitch_elems < (size_t)side) return fail(RN_ERR_BAD_PARAM, "pitch < side");
  // enqueue on every device, then wait: the devices work concurrently although one thread drives them
  for (int k = 0; k < m->n; k++) {
    rn_window w;
    rn_slab_window(side, k, m->n, &w);
    if (w.ni == 0) continue;
    rn_context* ctx = m->ctx[k];
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
    if (slab_dst) {
      if (!slab_dst[k] || reinterpret_cast<uintptr_t>(slab_dst[k]) % 16 != 0) return fail(RN_ERR_BAD_PARAM, "slab_dst[k] is NULL or not 16-byte aligned");
      st = synth_window_fused(ctx, p, w, reinterpret_cast<Elem*>(slab_dst[k]), slab_pitch_elems, 0, nullptr, 1, ctx->stream);
      if (st != RN_OK) return st;
      if (host_dst) {  // the slab is already in device memory: ship it from there
        RN_CUDA(cudaEventRecord(ctx->chunk_ready[0], ctx->stream));
        RN_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ready[0], 0));
        RN_CUDA(cudaMemcpy2DAsync(host_dst + (size_t)w.i0 * host_pitch_elems, host_pitch_elems * sizeof(Elem), slab_dst[k], slab_pitch_elems * sizeof(Elem),
                                  (size_t)side * sizeof(Elem), (size_t)w.ni, cudaMemcpyDeviceToHost, ctx->copy_stream));
      }
    } else {
      st = enqueue_window_to_host(ctx, p, w, host_dst + (size_t)w.i0 * host_pitch_elems, host_pitch_elems, ctx->stream);
      if (st != RN_OK) return st;
    }
  }
  return sync_all(m, host_dst != nullptr);
}

int rn_terrain2d_batch_sharded(rn_multi* m, const rn_params* p, const uint32_t* seeds, int ntiles, rn_element* const* tile_dst,
                               size_t tile_stride_elems)
{
  if (!m) return fail(RN_ERR_BAD_PARAM, "multi is NULL");
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!seeds || !tile_dst || ntiles < 0) return fail(RN_ERR_BAD_PARAM, "seeds/tile_dst is NULL or ntiles < 0");
  std::vector<uint32_t> mine;
  for (int k = 0; k < m->n; k++) {
    mine.clear();
    for (int t = k; t < ntiles; t += m->n) mine.push_back(seeds[t]);
    if (mine.empty()) continue;
    if (!tile_dst[k]) return fail(RN_ERR_BAD_PARAM, "tile_dst[k] is NULL");
    // (the seeds are staged by the runtime before the call returns: `mine` may be reused)
    st = rn_terrain2d_batch(m->ctx[k], p, mine.data(), (int)mine.size(), tile_dst[k], tile_stride_elems, 1, m->ctx[k]->stream);
    if (st != RN_OK) return st;
  }
  return sync_all(m, false);
}

int rn_gather(rn_multi* m, const rn_params* p, const rn_element* const* slab_src, rn_element* const* full_dst, int use_nccl)
{
  if (!m) return fail(RN_ERR_BAD_PARAM, "multi is NULL");
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!slab_src || !full_dst) return fail(RN_ERR_BAD_PARAM, "slab_src/full_dst is NULL");
  for (int k = 0; k < m->n; k++)
    if (!slab_src[k] || !full_dst[k]) return fail(RN_ERR_BAD_PARAM, "slab_src[k]/full_dst[k] is NULL");
  const int side = (1 << p->depth) + 1;
  const size_t slab_elems = (size_t)rn_slab_rows(side, m->n) * (size_t)side;
  if (use_nccl) {
    if (!m->distinct) return fail(RN_ERR_UNSUPPORTED, "NCCL needs distinct devices (a device is listed twice): use the peer-copy form");
    st = bind_nccl(m);
    if (st != RN_OK) return st;
    if (!m->comms_ready) {
      const ncclResult_t r = m->pCommInitAll(m->comms, m->n, m->devices);
      if (r != ncclSuccess) return fail_nccl(m, r, "ncclCommInitAll");
      m->comms_ready = true;
    }
    ncclResult_t r = m->pGroupStart();
    if (r != ncclSuccess) return fail_nccl(m, r, "ncclGroupStart");
    for (int k = 0; k < m->n; k++) {
      r = m->pAllGather(slab_src[k], full_dst[k], slab_elems * 4, ncclFloat, m->comms[k], m->ctx[k]->stream);
      if (r != ncclSuccess) {
        m->pGroupEnd();
        return fail_nccl(m, r, "ncclAllGather");
      }
    }
    r = m->pGroupEnd();
    if (r != ncclSuccess) return fail_nccl(m, r, "ncclGroupEnd");
  } else {
    for (int k = 0; k < m->n; k++) {
      DeviceGuard guard(m->ctx[k]->device);
      if (!guard.ok) return fail(RN_ERR_CUDA, "cudaSetDevice failed");
      for (int j = 0; j < m->n; j++)
        RN_CUDA(cudaMemcpyPeerAsync(full_dst[k] + (size_t)j * slab_elems, m->devices[k], slab_src[j], m->devices[j], slab_elems * sizeof(Elem),
                                    m->ctx[k]->stream));
    }
  }
  return sync_all(m, false);
}

}  // extern "C"
