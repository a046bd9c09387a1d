// rn_volume.cu -- 3D density-volume EXTENSION (BASELINE.json configs 512^3 / 2048^3).
//
// The reference has no volume generator; the definition is the CPU specification described in DESIGN.md
// section 5.8 (3D midpoint displacement in the reference's style: cube centres, face centres, edge midpoints per level,
// amplitude law of Rock.cpp:185, random() of basics.h:16-19, the draw of a lattice point an integer hash of (seed, i, j, k)).
//
// Round 2: one launch per LEVEL, tiled in shared memory.  Level l (n = 2^l cells per axis) lives in its own dense (n+1)^3
// grid; k_vol_level refines grid l into grid l+1 (the last one is the caller's slab).  A CTA owns a tile of TI x TJ x TK
// input cells = 2TI x 2TJ x 2TK output samples: it stages the tile's input lattice plus a halo of one cell on the low side of
// every axis (what the dependency cone edge <- face <- cube <- corner reaches; recomputed, never exchanged), runs the three
// passes of the level in shared memory with one thread per CELL (its cube centre, then its three low faces, then its three
// low edges: the index arithmetic, the lattice hash's coordinate products and the corner loads are shared by the cell's
// seven new samples), and stores the finished tile as full rows, unit stride along k.  Rows of the shared-memory image keep
// even and odd k in two halves, so that both the per-cell passes (stride 2 in k) and the copy-out (stride 1) are
// conflict-free.  Tiles that touch no face of the volume run an instantiation without existence tests.
// A slab of planes needs only a bounded neighbourhood of every coarser level, so i-slabs of one volume are independent --
// the multi-GPU form recomputes that neighbourhood instead of exchanging it -- and every level grid is allocated for the
// planes the slab needs only.  Bit-exact against the oracle.
#include "rn_kernels.h"
#include "rn_volume.cuh"
#include "rn_volrow.cuh"
#include <atomic>

namespace rn {

namespace {

template <int TI, int TJ, int TK, int NT>
__global__ void __launch_bounds__(NT) k_vol_level(const __grid_constant__ VolLevel L)
{
  extern __shared__ __align__(16) float sm[];
  const int A0 = (L.tile_i0 + (int)blockIdx.z) * TI, B0 = (int)blockIdx.y * TJ, C0 = (int)blockIdx.x * TK;
  // interior tiles: every cell -1 .. T of the tile exists and no closing plane is owned
  const bool edge = A0 < 1 || B0 < 1 || C0 < 1 || A0 + TI >= L.n || B0 + TJ >= L.n || C0 + TK >= L.n;
  if (edge) vol_tile<TI, TJ, TK, NT, true>(L, sm, A0, B0, C0);
  else vol_tile<TI, TJ, TK, NT, false>(L, sm, A0, B0, C0);
}

// levels n = 2, 4, 8 in one CTA, each as ONE whole-volume tile; a level reads what the level before wrote to global memory
// (same block: ordered by the barrier)
struct VolCoarse {
  VolLevel lv[kVolFusedLevels];
};
constexpr int kVolCoarseThreads = 256;
__global__ void __launch_bounds__(kVolCoarseThreads) k_vol_coarse(const __grid_constant__ VolCoarse C)
{
  extern __shared__ __align__(16) float sm[];
  vol_tile<2, 2, 2, kVolCoarseThreads, true>(C.lv[0], sm, 0, 0, 0);
  __threadfence_block();
  __syncthreads();
  vol_tile<4, 4, 4, kVolCoarseThreads, true>(C.lv[1], sm, 0, 0, 0);
  __threadfence_block();
  __syncthreads();
  vol_tile<8, 8, 8, kVolCoarseThreads, true>(C.lv[2], sm, 0, 0, 0);
}
cudaError_t launch_coarse(const VolCoarse& C, cudaStream_t stream)
{
  k_vol_coarse<<<1, kVolCoarseThreads, VolGeo<8, 8, 8>::BYTES, stream>>>(C);
  return cudaGetLastError();
}

template <int TI, int TJ, int TK, int NT>
cudaError_t launch_level(const VolLevel& L, int tiles_i, cudaStream_t stream)
{
  using G = VolGeo<TI, TJ, TK>;
  static std::atomic<bool> configured[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(k_vol_level<TI, TJ, TK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BYTES);
    if (e != cudaSuccess) return e;
    if (dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  const dim3 grid((unsigned)((L.n + TK - 1) / TK), (unsigned)((L.n + TJ - 1) / TJ), (unsigned)tiles_i);
  k_vol_level<TI, TJ, TK, NT><<<grid, NT, G::BYTES, stream>>>(L);
  return cudaGetLastError();
}

// The warp-row kernel (rn_volrow.cuh): levels with n >= kRowMinN cells per axis.  NW warps; the phases of a level are separated
// by block barriers, the row jobs of a phase go round-robin over the warps.  The job tables are compile-time constants in
// constant memory (a warp reads its job with a warp-uniform index).
__constant__ constexpr RowJobs<kRowTI, kRowTJ, false> g_jobs_lean = make_row_jobs<kRowTI, kRowTJ, false>();
__constant__ constexpr RowJobs<kRowTI, kRowTJ, true> g_jobs_edge = make_row_jobs<kRowTI, kRowTJ, true>();
__constant__ constexpr RowJobs<kRowSmallT, kRowSmallT, false> g_jobs_small_lean = make_row_jobs<kRowSmallT, kRowSmallT, false>();
__constant__ constexpr RowJobs<kRowSmallT, kRowSmallT, true> g_jobs_small_edge = make_row_jobs<kRowSmallT, kRowSmallT, true>();
template <int TI, int TJ, bool EDGE>
__device__ __forceinline__ const RowJobs<TI, TJ, EDGE>& row_jobs()
{
  static_assert((TI == kRowTI && TJ == kRowTJ) || (TI == kRowSmallT && TJ == kRowSmallT), "one job table per tile shape");
  if constexpr (TI == kRowTI && TJ == kRowTJ) {
    if constexpr (EDGE) return g_jobs_edge;
    else return g_jobs_lean;
  } else {
    if constexpr (EDGE) return g_jobs_small_edge;
    else return g_jobs_small_lean;
  }
}

template <int TI, int TJ, int NW>
__global__ void __launch_bounds__(NW * 32) k_vol_rows(const __grid_constant__ VolLevel L)
{
  extern __shared__ __align__(16) float sm[];
  const int A0 = (L.tile_i0 + (int)blockIdx.z) * TI, B0 = (int)blockIdx.y * TJ, C0 = (int)blockIdx.x * kRowTK;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool edge = A0 < 1 || B0 < 1 || A0 + TI >= L.n || B0 + TJ >= L.n;  // block-uniform: an i or j face of the volume
  if (edge) {
    using V = VolRows<TI, TJ, true>;
    const typename V::Tile t = V::tile(L, A0, B0, C0);
    V::zero(sm, tid, NW * 32);
    __syncthreads();
    V::load(L, t, sm, warp, lane, NW, tid);
    __syncthreads();
    V::cubes(L, t, sm, row_jobs<TI, TJ, true>(), warp, lane, NW, tid);
    __syncthreads();
    V::faces(L, t, sm, row_jobs<TI, TJ, true>(), warp, lane, NW, tid);
    __syncthreads();
    V::edges(L, t, sm, row_jobs<TI, TJ, true>(), warp, lane, NW, tid);
    __syncthreads();
    V::store(L, t, sm, warp, lane, NW);
  } else {
    using V = VolRows<TI, TJ, false>;
    const typename V::Tile t = V::tile(L, A0, B0, C0);
    V::load(L, t, sm, warp, lane, NW, tid);
    __syncthreads();
    V::cubes(L, t, sm, row_jobs<TI, TJ, false>(), warp, lane, NW, tid);
    __syncthreads();
    V::faces(L, t, sm, row_jobs<TI, TJ, false>(), warp, lane, NW, tid);
    __syncthreads();
    V::edges(L, t, sm, row_jobs<TI, TJ, false>(), warp, lane, NW, tid);
    __syncthreads();
    V::store(L, t, sm, warp, lane, NW);
  }
}

template <int TI, int TJ, int NW>
cudaError_t launch_rows(const VolLevel& L, int tiles_i, cudaStream_t stream)
{
  constexpr size_t BYTES = VolRows<TI, TJ, true>::BYTES;  // the EDGE table is the longer one
  static std::atomic<bool> configured[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(k_vol_rows<TI, TJ, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BYTES);
    if (e != cudaSuccess) return e;
    if (dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  const dim3 grid((unsigned)(L.n / kRowTK), (unsigned)(L.n / TJ), (unsigned)tiles_i);
  k_vol_rows<TI, TJ, NW><<<grid, NW * 32, BYTES, stream>>>(L);
  return cudaGetLastError();
}

}  // namespace

size_t volume_scratch_bytes(int depth, int i0, int ni)
{
  int lo[32], hi[32];
  volume_ranges(depth, i0, ni, lo, hi);
  size_t total = 0;
  for (int l = 1; l < depth; l++) {
    const size_t m = ((size_t)1 << l) + 1;
    total += (((size_t)(hi[l] - lo[l] + 1) * m * m * sizeof(float) + 255) / 256) * 256;
  }
  return total ? total : 256;
}

// Planes [i0, i0+ni) of the (W+1)^3 volume into `slab` (device memory, ni x (W+1) x (W+1) floats, dense); scratch:
// volume_scratch_bytes(depth, i0, ni) bytes of device memory.
// events (optional, depth entries): events[0] is recorded before the first level, events[l] after level l; *nstages = depth - 1.
// coarse / coarse_done (optional): every level but the last runs on the side stream `coarse` and records coarse_done, which
// `stream` waits for before the last level -- so that the latency-bound coarse levels of a call overlap the last level of
// the call before (rn_abi.cu: two scratch slots).
cudaError_t volume_generate(float* slab, void* scratch, int depth, uint32_t seed, float amplitude, int i0, int ni, cudaStream_t out_stream,
                            uint64_t* launches, cudaEvent_t* events, int* nstages, cudaStream_t coarse, cudaEvent_t coarse_done)
{
  const bool split = coarse && coarse_done && depth >= 3;
  cudaStream_t stream = split ? coarse : out_stream;
  int lo[32], hi[32];
  volume_ranges(depth, i0, ni, lo, hi);
  float* grid[32] = {};
  size_t off = 0;
  for (int l = 1; l < depth; l++) {
    const size_t m = ((size_t)1 << l) + 1;
    grid[l] = reinterpret_cast<float*>(static_cast<char*>(scratch) + off);
    off += (((size_t)(hi[l] - lo[l] + 1) * m * m * sizeof(float) + 255) / 256) * 256;
  }
  grid[depth] = slab;
  // level 1: the 27 points at multiples of W/2 start at 0
  cudaError_t e = cudaMemsetAsync(grid[1], 0, (size_t)(hi[1] - lo[1] + 1) * 9 * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  if (events && (e = cudaEventRecord(events[0], stream)) != cudaSuccess) return e;
  int first = 1;
  if (vol_coarse_fused(depth, lo, hi)) {
    VolCoarse C;
    for (int l = 1; l <= kVolFusedLevels; l++) {
      int tiles_i = 0;
      C.lv[l - 1] = vol_level_make(depth, l, lo, hi, grid[l], grid[l + 1], seed, amplitude, &tiles_i);
      C.lv[l - 1].tile_i0 = 0;
    }
    if ((e = launch_coarse(C, stream)) != cudaSuccess) return e;
    for (int l = 1; l <= kVolFusedLevels; l++)  // (one launch: the levels' events coincide)
      if (events && (e = cudaEventRecord(events[l], stream)) != cudaSuccess) return e;
    if (launches) *launches += 1;
    first = kVolFusedLevels + 1;
  }
  if (nstages) *nstages = depth - 1;
  for (int l = first; l < depth; l++) {
    const int n = 1 << l;
    int tiles_i = 0;
    const VolLevel L = vol_level_make(depth, l, lo, hi, grid[l], grid[l + 1], seed, amplitude, &tiles_i);
    if (split && l == depth - 1) {  // the last level: on the caller's stream, behind the coarse levels
      if ((e = cudaEventRecord(coarse_done, coarse)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(out_stream, coarse_done, 0)) != cudaSuccess) return e;
      stream = out_stream;
    }
    if (n >= kRowBigN) e = launch_rows<kRowTI, kRowTJ, kRowWarps>(L, tiles_i, stream);
    else if (n >= kRowMinN) e = launch_rows<kRowSmallT, kRowSmallT, kRowWarps>(L, tiles_i, stream);
    else switch (n) {
      case 2: e = launch_level<2, 2, 2, 64>(L, tiles_i, stream); break;
      case 4: e = launch_level<4, 4, 4, 128>(L, tiles_i, stream); break;
#if RN_VOL_COARSE_TILE == 4
      default: e = launch_level<4, 4, 4, 128>(L, tiles_i, stream); break;
#else
      case 8: e = launch_level<8, 8, 8, kVolThreads>(L, tiles_i, stream); break;
      default: e = launch_level<kVolTI, kVolTJ, kVolTK, kVolThreads>(L, tiles_i, stream); break;
#endif
    }
    if (e != cudaSuccess) return e;
    if (events && (e = cudaEventRecord(events[l], stream)) != cudaSuccess) return e;
    if (launches) *launches += 1;
  }
  if (split && stream != out_stream) {  // every level was part of the fused coarse launch
    if ((e = cudaEventRecord(coarse_done, coarse)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(out_stream, coarse_done, 0)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace rn
