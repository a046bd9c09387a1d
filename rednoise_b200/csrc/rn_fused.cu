// rn_fused.cu -- algorithm 2: the fused tile kernels (rn_tile.cuh), the streaming kernels (rn_stream.cuh) and their stage chain.
//
// Replaces the level loop of Terrain::Terrain (Terrain.cpp:60-82) + addVolcano (:88) with 2-6 launches per field (rn_plan.h):
// tile stages refine a complete coarse grid by 1-4 levels per CTA in shared memory and store finished tiles with 128-bit
// coalesced writes; the last two levels of a deep field are streamed.  Intermediate grids stay L2-resident between stages.
#include "rn_kernels.h"
#include "rn_plan.h"
#include <atomic>

namespace rn {

namespace {

constexpr int kStageThreads = kStageThreadsBig;

template <int L, int T, int NT, bool EDGE, int P>
__device__ __forceinline__ void run_phases(const StageParams& sp, const TileCtx& cx, float* sm)
{
  Tile<L, T, NT, EDGE>::template phase<P>(sp, cx, sm, (int)threadIdx.x);
  if constexpr (P + 1 < Tile<L, T, NT, EDGE>::NPHASES) {
    __syncthreads();
    run_phases<L, T, NT, EDGE, P + 1>(sp, cx, sm);
  }
}

#ifdef RN_TIMELINE
// Instrumented builds only: first CTA start / last CTA end (%globaltimer) of every launch of the chain.
__device__ unsigned long long g_tl_start[1024], g_tl_end[1024];
struct TlScope {
  int s;
  __device__ TlScope(int slot) : s(slot) { mark(false); }
  __device__ ~TlScope() { mark(true); }
  __device__ void mark(bool end) const
  {
    if (threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (end) atomicMax(&g_tl_end[s], t); else atomicMin(&g_tl_start[s], t);
    }
  }
};
#define RN_TL_SCOPE(slot) TlScope tl_scope__(slot)
#else
#define RN_TL_SCOPE(slot)
#endif

#ifdef RN_STREAM_TRACE
__device__ unsigned long long g_trace[4 * 8192];  // per unit: SM, start, end (%globaltimer), border flag
#endif

template <int L, int T, int NT>
__global__ void __launch_bounds__(NT, T == kTileSmall ? 8 : ((TileGeo<L, T>::BYTES <= 113 * 1024) ? 2 : 1)) k_stage(const __grid_constant__ StageParams sp)
{
  extern __shared__ __align__(16) float sm[];
  RN_TL_SCOPE(sp.tl_slot);
  TileCtx cx;
  cx.I0 = (sp.tile_i0 + (int)blockIdx.y) * T;
  cx.J0 = (sp.tile_j0 + (int)blockIdx.x) * T;
  const unsigned f = blockIdx.z;
  cx.seed = sp.seeds ? sp.seeds[f] : sp.seed;
  cx.gin = sp.gin ? sp.gin + (size_t)f * sp.gin_field_stride : nullptr;
  cx.gout = sp.gout + (size_t)f * sp.gout_field_stride;
  // interior tiles (no part of the dependency cone outside the lattice) take the lean instantiation
  if (Tile<L, T, NT, true>::is_edge(cx, 1 << sp.dout)) run_phases<L, T, NT, true, 0>(sp, cx, sm);
  else run_phases<L, T, NT, false, 0>(sp, cx, sm);
}

// Streaming form of one level (rn_stream.cuh): blockIdx.x = column strip, blockIdx.y = row segment.
template <int NT, bool EDGE, int PH>
__device__ __forceinline__ void stream_step(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int u, int c)
{
  using S = Stream<NT, EDGE>;
  S::template interval1<PH>(sp, cx, sm, th, u, c);
  __syncthreads();
  S::template interval2<PH>(sp, cx, sm, th, u, c);
  __syncthreads();
  S::template interval3<PH>(sp, cx, sm, th, u, c);
  __syncthreads();
}

template <int NT>
__device__ __forceinline__ void stream_march(const StreamParams& sp, const StreamCtx& cx, float* sm)
{
  using S = Stream<NT, true>;
  const int c = (int)threadIdx.x;
  StreamThread th;
  S::init(sp, cx, sm, th, c);
  __syncthreads();
  // Every warp executes the same three barriers per step, in whichever instantiation its columns and the step's rows allow.
  const bool warp_lean = stream_warp_is_lean(sp, cx, c) && cx.lean_lo <= cx.lean_hi;
  const int last = cx.a1e;
  int u = cx.a0 - 1;
  while (u <= last) {
    if (warp_lean && u == cx.lean_lo) {
      const int hi = cx.lean_hi;
#if RN_STREAM_UNROLL == 4
      // four steps per trip: the ring slot of every rolling row is a compile-time constant (step_of(u) & 3 = PH)
      for (;;) {
        stream_step<NT, false, 0>(sp, cx, sm, th, u, c);
        if (u == hi) break;
        stream_step<NT, false, 1>(sp, cx, sm, th, u + 1, c);
        if (u + 1 == hi) break;
        stream_step<NT, false, 2>(sp, cx, sm, th, u + 2, c);
        if (u + 2 == hi) break;
        stream_step<NT, false, 3>(sp, cx, sm, th, u + 3, c);
        if (u + 3 == hi) break;
        u += 4;
      }
#elif RN_STREAM_UNROLL == 2
      // two steps per trip: half the code of the four-step form (it no longer fits the 32 KB instruction cache)
      for (;;) {
        stream_step<NT, false, 4>(sp, cx, sm, th, u, c);
        if (u == hi) break;
        stream_step<NT, false, 5>(sp, cx, sm, th, u + 1, c);
        if (u + 1 == hi) break;
        u += 2;
      }
#else
      for (; u <= hi; u++) stream_step<NT, false, -1>(sp, cx, sm, th, u, c);
#endif
      u = hi + 1;
    } else {
      stream_step<NT, true, -1>(sp, cx, sm, th, u, c);
      u++;
    }
  }
  S::store_rows(sp, cx, sm, th, cx.a1e - 1, S::mod(S::step_of(cx, cx.a1e - 1), 2));
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_stream(const __grid_constant__ StreamParams sp)
{
  extern __shared__ __align__(16) float sm[];
  RN_TL_SCOPE(sp.tl_slot);
  const StreamCtx cx = stream_ctx(sp, (int)blockIdx.x, blockIdx.z);
#ifdef RN_STREAM_TRACE
  const unsigned bid = blockIdx.x;
  if (threadIdx.x == 0 && bid < 8192) {
    unsigned smid; unsigned long long t;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[4 * bid] = smid; g_trace[4 * bid + 1] = t; g_trace[4 * bid + 3] = !cx.col_lean;
  }
#endif
  stream_march<NT>(sp, cx, sm);
#ifdef RN_STREAM_TRACE
  if (threadIdx.x == 0 && bid < 8192) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[4 * bid + 2] = t;
  }
#endif
}

// Warp-synchronous march (rn_march.cuh): every warp is an independent (strip, segment) unit; no block barriers.
#ifndef RN_MARCH_VM
#define RN_MARCH_VM 1   // lean steps instantiated per addVolcano mode (off / far from the crater / with the crater branch)
#endif
template <bool EDGE, int VM>
__device__ __forceinline__ void march_step(const StreamParams& sp, const MarchCtx& cx, MarchLane& L, E4* wsm, int u, int lane)
{
  using M = March<EDGE>;
  M::p1(sp, cx, L, wsm, u, lane);
  M::x1(cx, L, wsm, u, lane, nullptr, nullptr);
  if (march_runs_p2(cx, u)) {  // uniform; false in the first warm-up step only
    M::p2(sp, cx, L, wsm, u, lane);
    M::x2(cx, L, wsm, u, lane, nullptr, nullptr);
  } else {
    M::skip_p2(sp, L);
  }
  if (march_runs_p3(cx, u)) {  // uniform; false in both warm-up steps
    M::template p3<VM>(sp, cx, L, wsm, u, lane);
    M::store(sp, cx, L, u, lane, nullptr);
  } else {
    M::skip_p3(sp, L);
  }
}

template <int WPB, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB) k_march(const __grid_constant__ StreamParams sp, const int nunits)
{
  extern __shared__ __align__(16) float sm[];
  RN_TL_SCOPE(sp.tl_slot);
  const int lane = (int)threadIdx.x & 31;
  // one warp per CTA (the shipped configuration): the unit index -- and with it the whole MarchCtx and the loop control -- is a
  // function of blockIdx alone, which lets the compiler keep them in uniform registers and prove the warp convergent
  const int wib = WPB == 1 ? 0 : ((int)threadIdx.x >> 5);
  const int id = (int)blockIdx.x * WPB + wib;
  if (id >= nunits) return;
  E4* wsm = reinterpret_cast<E4*>(sm) + wib * kWarpE4;  // this warp's ring, carried state and constants
  const MarchCtx cx = march_ctx(sp, id, blockIdx.z);
#ifdef RN_STREAM_TRACE
  if (lane == 0 && id < 8192) {
    unsigned smid; unsigned long long t;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[4 * id] = smid; g_trace[4 * id + 1] = t; g_trace[4 * id + 3] = !cx.col_lean;
  }
#endif
  MarchLane L;
  March<true>::init(sp, cx, L, wsm, lane, false);
  const int last = cx.a1e;
#if RN_MARCH_VM
  const int vm = !sp.volcano ? 0 : cx.far ? 1 : 2;  // addVolcano mode of the warp's lean steps (rn_march.cuh p3)
#endif
  for (int u = cx.a0 - 1; u <= last; u++) {
    if (march_step_is_lean(cx, u)) {  // uniform per warp
#if RN_MARCH_VM
      if (vm == 1) march_step<false, 1>(sp, cx, L, wsm, u, lane);
      else if (vm == 0) march_step<false, 0>(sp, cx, L, wsm, u, lane);
      else march_step<false, 2>(sp, cx, L, wsm, u, lane);
#else
      march_step<false, -1>(sp, cx, L, wsm, u, lane);
#endif
    } else march_step<true, -1>(sp, cx, L, wsm, u, lane);
  }
#ifdef RN_STREAM_TRACE
  if (lane == 0 && id < 8192) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[4 * id + 2] = t;
  }
#endif
}

cudaError_t launch_march(const StreamParams& st, dim3 grid_units, cudaStream_t stream)
{
  constexpr int kBytes = kMarchWPB * kWarpE4 * (int)sizeof(E4);
  static std::atomic<bool> configured[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(k_march<kMarchWPB, kMarchWarpsPerSM / kMarchWPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes);
    if (e != cudaSuccess) return e;
    if (dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  const int nunits = (int)grid_units.x;
  const dim3 grid((unsigned)((nunits + kMarchWPB - 1) / kMarchWPB), 1, grid_units.z);
  k_march<kMarchWPB, kMarchWarpsPerSM / kMarchWPB><<<grid, 32 * kMarchWPB, kBytes, stream>>>(st, nunits);
  return cudaGetLastError();
}

template <int NT, int MINB>
cudaError_t launch_stream_t(const StreamParams& st, dim3 grid, cudaStream_t stream)
{
  using G = StreamGeo<NT>;
  static std::atomic<bool> configured[64] = {};  // contexts may launch from several host threads
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(k_stream<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BYTES);
    if (e != cudaSuccess) return e;
    if (dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  k_stream<NT, MINB><<<grid, NT, G::BYTES, stream>>>(st);
  return cudaGetLastError();
}

cudaError_t launch_stream(const StreamParams& st, int nt, dim3 grid, cudaStream_t stream)
{
  if (nt == kStreamThreadsSmall) return launch_stream_t<kStreamThreadsSmall, kStreamMinBlocksSmall>(st, grid, stream);
  return launch_stream_t<kStreamThreads, kStreamMinBlocks>(st, grid, stream);
}

template <int L, int T, int NT>
cudaError_t launch_stage_t(const StageParams& sp, dim3 grid, cudaStream_t stream)
{
  using G = TileGeo<L, T>;
  static std::atomic<bool> configured[64] = {};  // contexts may launch from several host threads
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(k_stage<L, T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BYTES);
    if (e != cudaSuccess) return e;
    if (dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  k_stage<L, T, NT><<<grid, NT, G::BYTES, stream>>>(sp);
  return cudaGetLastError();
}

template <int L>
cudaError_t launch_stage(const StageParams& sp, int T, dim3 grid, cudaStream_t stream)
{
#if RN_SMALL_CHAIN
  if constexpr (L <= RN_SMALL_CHAIN_L) {
    if (T == kTileSmall) return launch_stage_t<L, kTileSmall, kStageThreadsSmall>(sp, grid, stream);
  }
#endif
  (void)T;
  return launch_stage_t<L, kTile, kStageThreads>(sp, grid, stream);
}

}  // namespace

#ifdef RN_STREAM_TRACE
extern "C" int rn_debug_trace(unsigned long long* host, int n)
{
  return (int)cudaMemcpyFromSymbol(host, g_trace, sizeof(unsigned long long) * (size_t)n);
}
#endif

#ifdef RN_TIMELINE
static int g_tl_next = 0;
extern "C" int rn_debug_timeline(unsigned long long* host, int reset)
{
  if (reset) {
    g_tl_next = 0;
    void *a, *b;
    cudaGetSymbolAddress(&a, g_tl_start);
    cudaGetSymbolAddress(&b, g_tl_end);
    cudaMemset(a, 0xff, sizeof(unsigned long long) * 1024);
    cudaMemset(b, 0, sizeof(unsigned long long) * 1024);
    return 0;
  }
  cudaMemcpyFromSymbol(host, g_tl_start, sizeof(unsigned long long) * 1024);
  cudaMemcpyFromSymbol(host + 1024, g_tl_end, sizeof(unsigned long long) * 1024);
  return g_tl_next;
}
#endif

size_t fused_scratch_bytes(int depth, int nfields) { return chain_scratch_bytes(depth, nfields); }
int fused_batch_chunk(int depth, int cmax, int sm_count) { return march_batch_chunk(depth, cmax, sm_count); }

// Synthesizes the window [i0,i0+ni) x [j0,j0+nj) of `nfields` depth-`depth` fields.
//   dst     element (i0, j0) of field 0; row pitch `pitch`, field stride `field_stride` (elements)
//   seeds   device pointer to nfields seeds, or nullptr (then every field uses `seed`)
//   scratch fused_scratch_bytes(depth, nfields) bytes of device memory
//   coarse_stream / coarse_done (optional): every stage but the last is launched on coarse_stream
//   and the last stage waits for it through the event, so that the small, latency-bound coarse
//   stages of one call overlap the output stage of the previous call (the caller must make sure
//   `scratch` is no longer read by an earlier call before coarse_stream reaches this one).
cudaError_t fused_generate(Elem* dst, size_t pitch, size_t field_stride, int depth, int i0, int j0, int ni, int nj, const Law& law,
                           float persistence, uint32_t seed, const uint32_t* seeds, int nfields, int volcano, float amplitude,
                           void* scratch, cudaStream_t stream, uint64_t* launches, cudaEvent_t* events, int* nstages,
                           cudaStream_t coarse_stream, cudaEvent_t coarse_done, cudaStream_t tile_stream, cudaEvent_t tiles_done, int sm_count,
                           int march)
{
  StageLaunch chain[kMaxLaunches];
  const int nl = build_chain(chain, depth, i0, j0, ni, nj, reinterpret_cast<E4*>(dst), pitch, field_stride, scratch, nfields, law,
                             persistence, seed, seeds, volcano, amplitude, sm_count > 0 ? sm_count : 148, march);
  if (nstages) *nstages = nl;
  if (events) {
    cudaError_t e0 = cudaEventRecord(events[0], stream);
    if (e0 != cudaSuccess) return e0;
  }
  bool has_coarse = false;
  for (int k = 0; k < nl; k++) has_coarse = has_coarse || !chain[k].last;
  const bool split = coarse_stream && coarse_done && coarse_stream != stream && !events && has_coarse;
  // Streams of a split call: tile stages on tile_stream (if given), streamed coarse levels on coarse_stream, the output stage on
  // the caller's stream; where the chain moves from one stream to the next an event carries the dependency.  The three streams
  // are in-order, so consecutive calls pipeline stage by stage (the caller guards the scratch slot, see rn_abi.cu).
  const bool three = split && tile_stream && tiles_done && tile_stream != stream && tile_stream != coarse_stream;
  cudaStream_t prev_s = nullptr;
  for (int k = 0; k < nl; k++) {
    StageLaunch& sl = chain[k];
#ifdef RN_TIMELINE
    sl.sp.tl_slot = sl.st.tl_slot = (g_tl_next++) & 1023;
#endif
    const dim3 grid(sl.gx, sl.gy, (unsigned)nfields);
    cudaStream_t s = stream;
    if (split && !sl.last) s = (three && sl.kind == 0) ? tile_stream : coarse_stream;
    cudaError_t e;
    if (split && k > 0 && s != prev_s) {  // the chain changes streams: this launch waits for the previous one
      cudaEvent_t ev = (three && prev_s == tile_stream) ? tiles_done : coarse_done;
      e = cudaEventRecord(ev, prev_s);
      if (e != cudaSuccess) return e;
      e = cudaStreamWaitEvent(s, ev, 0);
      if (e != cudaSuccess) return e;
    }
    prev_s = s;
    if (sl.kind == 2) {
      e = launch_march(sl.st, grid, s);
    } else if (sl.kind == 1) {
      e = launch_stream(sl.st, sl.nt, grid, s);
    } else {
      switch (sl.L) {
        case 1: e = launch_stage<1>(sl.sp, sl.T, grid, s); break;
        case 2: e = launch_stage<2>(sl.sp, sl.T, grid, s); break;
        case 3: e = launch_stage<3>(sl.sp, sl.T, grid, s); break;
        default: e = launch_stage<4>(sl.sp, sl.T, grid, s); break;
      }
    }
    if (e != cudaSuccess) return e;
    if (events) {
      e = cudaEventRecord(events[k + 1], stream);
      if (e != cudaSuccess) return e;
    }
  }
  if (launches) *launches += (uint64_t)nl;
  return cudaSuccess;
}

}  // namespace rn
