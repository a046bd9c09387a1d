// tests/host_replay/rn_replay.cpp -- TEST SCAFFOLDING ONLY (never linked into the product).
//
// Replays the fused tile algorithm (rednoise_b200/csrc/rn_tile.cuh + rn_plan.h, the very code
// the CUDA kernels are compiled from) on the CPU: every phase is executed for every thread
// index of every tile, with the barrier between phases modelled by finishing a phase for all
// threads before the next one starts.  This validates tiling, halos, windows, boundary rules,
// RNG indexing and shade ordering against the oracle in a container without a GPU.  With
// `poison` the tile's "shared memory" is filled with signalling garbage (NaN) first, proving
// that no out-of-cone value ever reaches an output sample, and every shared-memory index is
// range-checked by construction (the buffer is exactly TileGeo::WORDS long inside guard pages
// of NaN that are verified untouched).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include "../../include/rednoise.h"
#include "../../rednoise_b200/csrc/rn_plan.h"
#include "../../rednoise_b200/csrc/rn_volume.cuh"
#include "../../rednoise_b200/csrc/rn_volrow.cuh"

using namespace rn;

namespace {

constexpr int threads_of(int T) { return T == kTileSmall ? kStageThreadsSmall : kStageThreadsBig; }
static int g_reverse = 0;  // visit the threads of a phase in descending order: a phase whose threads depended on
                           // one another's writes (a race on the GPU) would give different bits
constexpr int GUARD = 4096;  // words on each side

template <int L, int T, bool EDGE, int P>
void run_phase(const StageParams& sp, const TileCtx& cx, float* sm)
{
  constexpr int NT = threads_of(T);
  for (int q = 0; q < NT; q++) Tile<L, T, NT, EDGE>::template phase<P>(sp, cx, sm, g_reverse ? NT - 1 - q : q);
  if constexpr (P + 1 < Tile<L, T, NT, EDGE>::NPHASES) run_phase<L, T, EDGE, P + 1>(sp, cx, sm);
}

template <int L, int T>
int run_stage_t(const StageLaunch& sl, int nfields, const uint32_t* seeds, int poison, int force_edge)
{
  constexpr int NT = threads_of(T);
  using G = TileGeo<L, T>;
  std::vector<float> buf((size_t)G::WORDS + 2 * GUARD + 4);
  // 16-byte align the tile's shared memory
  float* base = buf.data();
  while ((reinterpret_cast<uintptr_t>(base + GUARD) & 15u) != 0) base++;
  float* sm = base + GUARD;
  uint32_t nanbits = 0x7fc0dead;
  float nanv;
  memcpy(&nanv, &nanbits, 4);
  int bad = 0;
  for (int f = 0; f < nfields; f++)
    for (unsigned by = 0; by < sl.gy; by++)
      for (unsigned bx = 0; bx < sl.gx; bx++) {
        for (size_t i = 0; i < (size_t)G::WORDS + 2 * GUARD; i++) base[i] = poison ? nanv : 0.0f;
        TileCtx cx;
        cx.I0 = (sl.sp.tile_i0 + (int)by) * T;
        cx.J0 = (sl.sp.tile_j0 + (int)bx) * T;
        cx.seed = seeds ? seeds[f] : sl.sp.seed;
        cx.gin = sl.sp.gin ? sl.sp.gin + (size_t)f * sl.sp.gin_field_stride : nullptr;
        cx.gout = sl.sp.gout + (size_t)f * sl.sp.gout_field_stride;
        if (force_edge || Tile<L, T, NT, true>::is_edge(cx, 1 << sl.sp.dout)) run_phase<L, T, true, 0>(sl.sp, cx, sm);
        else run_phase<L, T, false, 0>(sl.sp, cx, sm);
        if (poison) {
          for (int i = 0; i < GUARD; i++) {
            uint32_t a, b;
            memcpy(&a, &base[i], 4);
            memcpy(&b, &sm[G::WORDS + i], 4);
            if (a != nanbits || b != nanbits) bad = 1;  // a write outside the tile's shared memory
          }
        }
      }
  return bad;
}

template <int L>
int run_stage(const StageLaunch& sl, int nfields, const uint32_t* seeds, int poison, int force_edge)
{
  if constexpr (L <= RN_SMALL_CHAIN_L) {
    if (sl.T == kTileSmall) return run_stage_t<L, kTileSmall>(sl, nfields, seeds, poison, force_edge);
  }
  return run_stage_t<L, kTile>(sl, nfields, seeds, poison, force_edge);
}

// One barrier interval for every thread, each in the instantiation the kernel's warp would run at this step.
template <int SNT, int K, int PH>
void stream_interval_thread(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int u, int c, bool lean)
{
  if (lean) {
    using S = Stream<SNT, false>;
    if (K == 1) S::template interval1<PH>(sp, cx, sm, th, u, c);
    if (K == 2) S::template interval2<PH>(sp, cx, sm, th, u, c);
    if (K == 3) S::template interval3<PH>(sp, cx, sm, th, u, c);
  } else {
    using S = Stream<SNT, true>;
    if (K == 1) S::template interval1<-1>(sp, cx, sm, th, u, c);
    if (K == 2) S::template interval2<-1>(sp, cx, sm, th, u, c);
    if (K == 3) S::template interval3<-1>(sp, cx, sm, th, u, c);
  }
}

template <int SNT, int K>
void stream_interval_host(const StreamParams& sp, const StreamCtx& cx, float* sm, std::vector<StreamThread>& th, int u, bool force_edge)
{
  const int ph = (u - cx.ubase) & 3;
  for (int q = 0; q < SNT; q++) {
    const int c = g_reverse ? SNT - 1 - q : q;
    const bool lean = !force_edge && stream_warp_is_lean(sp, cx, c) && stream_step_is_lean(cx, u);
    // the lean instantiation is replayed in the kernel's unrolled form (ring slots as template constants): the form
    // RN_STREAM_UNROLL selects, and the other two on alternating cells so that every form stays covered
    const int form = ((u + c) % 3 == 0) ? RN_STREAM_UNROLL : ((u + c) % 3 == 1 ? 4 : 1);
    const int sel = form == 4 ? ph : (form == 2 ? 4 + (ph & 1) : -1);
    switch (sel) {
      case 0: stream_interval_thread<SNT, K, 0>(sp, cx, sm, th[c], u, c, lean); break;
      case 1: stream_interval_thread<SNT, K, 1>(sp, cx, sm, th[c], u, c, lean); break;
      case 2: stream_interval_thread<SNT, K, 2>(sp, cx, sm, th[c], u, c, lean); break;
      case 3: stream_interval_thread<SNT, K, 3>(sp, cx, sm, th[c], u, c, lean); break;
      case 4: stream_interval_thread<SNT, K, 4>(sp, cx, sm, th[c], u, c, lean); break;
      case 5: stream_interval_thread<SNT, K, 5>(sp, cx, sm, th[c], u, c, lean); break;
      default: stream_interval_thread<SNT, K, -1>(sp, cx, sm, th[c], u, c, lean); break;
    }
  }
}

template <int SNT>
void stream_march_host(const StreamParams& sp, const StreamCtx& cx, float* sm, std::vector<StreamThread>& th, bool force_edge)
{
  using S = Stream<SNT, true>;
#define RN_ALL_THREADS(call) for (int q = 0; q < SNT; q++) { const int c = g_reverse ? SNT - 1 - q : q; call; }
  RN_ALL_THREADS(S::init(sp, cx, sm, th[c], c))
  for (int u = cx.a0 - 1; u <= cx.a1e; u++) {
    stream_interval_host<SNT, 1>(sp, cx, sm, th, u, force_edge);
    stream_interval_host<SNT, 2>(sp, cx, sm, th, u, force_edge);
    stream_interval_host<SNT, 3>(sp, cx, sm, th, u, force_edge);
  }
  RN_ALL_THREADS(S::store_rows(sp, cx, sm, th[c], cx.a1e - 1, S::mod(S::step_of(cx, cx.a1e - 1), 2)))
#undef RN_ALL_THREADS
}

template <int SNT>
int run_stream_t(const StageLaunch& sl, int nfields, const uint32_t* seeds, int poison, int force_edge)
{
  using G = StreamGeo<SNT>;
  const StreamParams& sp0 = sl.st;
  std::vector<float> buf((size_t)G::WORDS + 2 * GUARD + 4);
  float* base = buf.data();
  while ((reinterpret_cast<uintptr_t>(base + GUARD) & 15u) != 0) base++;
  float* sm = base + GUARD;
  uint32_t nanbits = 0x7fc0dead;
  float nanv;
  memcpy(&nanv, &nanbits, 4);
  int bad = 0;
  std::vector<StreamThread> th(SNT);
  for (int f = 0; f < nfields; f++)
    for (unsigned by = 0; by < sl.gy; by++)
      for (unsigned bx = 0; bx < sl.gx; bx++) {
        for (size_t i = 0; i < (size_t)G::WORDS + 2 * GUARD; i++) base[i] = poison ? nanv : 0.0f;
        StreamParams sp = sp0;
        if (seeds) { sp.seeds = nullptr; sp.seed = seeds[f]; }
        const StreamCtx cx = stream_ctx(sp, (int)bx, (unsigned)f);
        memset(th.data(), poison ? 0xff : 0, sizeof(StreamThread) * th.size());
        stream_march_host<SNT>(sp, cx, sm, th, force_edge != 0);
        if (poison) {
          for (int i = 0; i < GUARD; i++) {
            uint32_t a, b;
            memcpy(&a, &base[i], 4);
            memcpy(&b, &sm[G::WORDS + i], 4);
            if (a != nanbits || b != nanbits) bad = 1;
          }
        }
      }
  return bad;
}

// The warp-synchronous march (rn_march.cuh): the 32 lanes of a warp are replayed phase by phase in lockstep; a shuffle is a read
// of the neighbouring lane's struct after every lane has finished the phase before it.  With `poison` the lane state and the
// warp's block of "shared memory" start as NaN and the values the device shuffles in from outside the warp (lane 0's left,
// lane 31's right) are NaN too; guard zones around the block are verified untouched.
template <bool EDGE>
void march_step_host(const StreamParams& sp, const MarchCtx& cx, std::vector<MarchLane>& L, E4* wsm, int u)
{
  using M = March<EDGE>;
#define RN_LANES(call) for (int q = 0; q < 32; q++) { const int l = g_reverse ? 31 - q : q; call; }
  RN_LANES(M::p1(sp, cx, L[l], wsm, u, l))
  RN_LANES(M::x1(cx, L[l], wsm, u, l, l > 0 ? &L[l - 1] : nullptr, l < 31 ? &L[l + 1] : nullptr))
  if (march_runs_p2(cx, u)) {  // (rn_fused.cu march_step)
    RN_LANES(M::p2(sp, cx, L[l], wsm, u, l))
    RN_LANES(M::x2(cx, L[l], wsm, u, l, l > 0 ? &L[l - 1] : nullptr, l < 31 ? &L[l + 1] : nullptr))
  } else {
    RN_LANES(M::skip_p2(sp, L[l]))
  }
  if (march_runs_p3(cx, u)) {
    RN_LANES(M::template p3<-1>(sp, cx, L[l], wsm, u, l))
    RN_LANES(M::store(sp, cx, L[l], u, l, l > 0 ? &L[l - 1] : nullptr))
  } else {
    RN_LANES(M::skip_p3(sp, L[l]))
  }
#undef RN_LANES
}

int run_march(const StageLaunch& sl, int nfields, const uint32_t* seeds, int poison, int force_edge)
{
  std::vector<MarchLane> L(32);
  constexpr int G4 = 64;  // guard, in E4
  std::vector<E4> buf(kWarpE4 + 2 * G4 + 1);
  E4* wsm = buf.data() + G4;
  int bad = 0;
  for (int f = 0; f < nfields; f++)
    for (unsigned id = 0; id < sl.gx; id++) {
      StreamParams sp = sl.st;
      if (seeds) { sp.seeds = nullptr; sp.seed = seeds[f]; }
      MarchCtx cx = march_ctx(sp, (int)id, (unsigned)f);
      if (force_edge) cx.col_lean = 0;
      memset(L.data(), poison ? 0xff : 0, sizeof(MarchLane) * L.size());
      memset(buf.data(), poison ? 0xff : 0, sizeof(E4) * buf.size());
      for (int l = 0; l < 32; l++) March<true>::init(sp, cx, L[l], wsm, l, poison != 0);
      // the kernel's loop (rn_fused.cu: k_march)
      for (int u = cx.a0 - 1; u <= cx.a1e; u++) {
        if (march_step_is_lean(cx, u)) march_step_host<false>(sp, cx, L, wsm, u);
        else march_step_host<true>(sp, cx, L, wsm, u);
      }
      if (poison) {
        const unsigned char* g0 = reinterpret_cast<const unsigned char*>(buf.data());
        const unsigned char* g1 = reinterpret_cast<const unsigned char*>(wsm + kWarpE4);
        for (size_t i = 0; i < sizeof(E4) * G4; i++)
          if (g0[i] != 0xff || g1[i] != 0xff) bad = 1;  // a write outside the warp's shared memory
      }
    }
  return bad;
}

int run_stream(const StageLaunch& sl, int nfields, const uint32_t* seeds, int poison, int force_edge)
{
  if (sl.nt == kStreamThreadsSmall) return run_stream_t<kStreamThreadsSmall>(sl, nfields, seeds, poison, force_edge);
  return run_stream_t<kStreamThreads>(sl, nfields, seeds, poison, force_edge);
}

}  // namespace

// The 3D volume path (rn_volume.cuh): the kernels' own tile routine, one "thread" per tile, over the same level chain and plane
// ranges as volume_generate (rn_volume.cu).  poison: NaN-fill the level grids and every tile's shared-memory image.
template <int TI, int TJ, int TK>
void vol_level_host(const VolLevel& L, int tiles_i, int poison)
{
  using G = VolGeo<TI, TJ, TK>;
  std::vector<float> sm(G::WORDS);
  uint32_t nanbits = 0x7fc0dead;
  float nanv;
  memcpy(&nanv, &nanbits, 4);
  const int tj = (L.n + TJ - 1) / TJ, tk = (L.n + TK - 1) / TK;
  for (int z = 0; z < tiles_i; z++)
    for (int y = 0; y < tj; y++)
      for (int x = 0; x < tk; x++) {
        for (auto& w : sm) w = poison ? nanv : 0.0f;
        const int A0 = (L.tile_i0 + z) * TI, B0 = y * TJ, C0 = x * TK;
        const bool edge = A0 < 1 || B0 < 1 || C0 < 1 || A0 + TI >= L.n || B0 + TJ >= L.n || C0 + TK >= L.n;
        if (edge) vol_tile<TI, TJ, TK, 1, true>(L, sm.data(), A0, B0, C0);
        else vol_tile<TI, TJ, TK, 1, false>(L, sm.data(), A0, B0, C0);
      }
}

// The warp-row form (rn_volrow.cuh): every phase for every warp and lane in turn (g_reverse: in descending order), the barrier
// between phases modelled by finishing a phase for all of them first.
template <int TI, int TJ, int NW, bool EDGE>
void vol_rows_tile_host(const VolLevel& L, float* sm, int A0, int B0, int C0)
{
  using V = VolRows<TI, TJ, EDGE>;
  static const RowJobs<TI, TJ, EDGE> jobs = make_row_jobs<TI, TJ, EDGE>();
  const typename V::Tile t = V::tile(L, A0, B0, C0);
  constexpr int NT = NW * 32;
  auto each = [&](auto&& f) {
    for (int q = 0; q < NT; q++) {
      const int tid = g_reverse ? NT - 1 - q : q;
      f(tid, tid >> 5, tid & 31);
    }
  };
  if (EDGE) each([&](int tid, int, int) { V::zero(sm, tid, NT); });
  each([&](int tid, int w, int l) { V::load(L, t, sm, w, l, NW, tid); });
  each([&](int tid, int w, int l) { V::cubes(L, t, sm, jobs, w, l, NW, tid); });
  each([&](int tid, int w, int l) { V::faces(L, t, sm, jobs, w, l, NW, tid); });
  each([&](int tid, int w, int l) { V::edges(L, t, sm, jobs, w, l, NW, tid); });
  each([&](int, int w, int l) { V::store(L, t, sm, w, l, NW); });
}

template <int TI, int TJ, int NW>
void vol_rows_host(const VolLevel& L, int tiles_i, int poison)
{
  std::vector<float> sm(VolRows<TI, TJ, true>::BYTES / sizeof(float));
  uint32_t nanbits = 0x7fc0dead;
  float nanv;
  memcpy(&nanv, &nanbits, 4);
  for (int z = 0; z < tiles_i; z++)
    for (int y = 0; y < L.n / TJ; y++)
      for (int x = 0; x < L.n / kRowTK; x++) {
        for (auto& w : sm) w = poison ? nanv : 0.0f;
        const int A0 = (L.tile_i0 + z) * TI, B0 = y * TJ, C0 = x * kRowTK;
        const bool edge = A0 < 1 || B0 < 1 || A0 + TI >= L.n || B0 + TJ >= L.n;
        if (edge) vol_rows_tile_host<TI, TJ, NW, true>(L, sm.data(), A0, B0, C0);
        else vol_rows_tile_host<TI, TJ, NW, false>(L, sm.data(), A0, B0, C0);
      }
}

extern "C" {

// random_from_draw2h (rn_volrow.cuh) against random_from_draw for one draw
int replay_random_pair_matches(unsigned r)
{
  const F2 p = random_from_draw2h(r, 32767u - r);
  const float a = random_from_draw(r), b = random_from_draw(32767u - r);
  return memcmp(&p.a, &a, 4) == 0 && memcmp(&p.b, &b, 4) == 0;
}
void replay_set_reverse(int r) { g_reverse = r; }

// random() of basics.h:16-19 as the kernels evaluate it (rn_math.cuh: the /32767 as a multiplication + one FMA correction)
float replay_random_from_draw(unsigned r) { return random_from_draw(r); }

int replay_smem_bytes(int L)
{
  switch (L) {
    case 1: return (int)TileGeo<1, kTile>::BYTES;
    case 2: return (int)TileGeo<2, kTile>::BYTES;
    case 3: return (int)TileGeo<3, kTile>::BYTES;
    default: return (int)TileGeo<4, kTile>::BYTES;
  }
}

int replay_plan(int depth, int* din, int* L)
{
  StagePlan plan[kMaxStages];
  const int ns = plan_stages(depth, plan, 1);
  for (int k = 0; k < ns; k++) {
    din[k] = plan[k].din;
    L[k] = plan[k].L;
  }
  return ns;
}

int replay_volume3d(uint32_t seed, int depth, float amplitude, int i0, int ni, float* slab, int poison)
{
  if (!slab || depth < 2 || depth > 11) return RN_ERR_BAD_PARAM;
  const int side = (1 << depth) + 1;
  if (i0 < 0 || ni <= 0 || i0 > side - ni) return RN_ERR_BAD_PARAM;
  int lo[32], hi[32];
  volume_ranges(depth, i0, ni, lo, hi);
  std::vector<std::vector<float>> grids(depth + 1);
  float* grid[32] = {};
  for (int l = 1; l < depth; l++) {
    const size_t m = ((size_t)1 << l) + 1;
    grids[l].assign((size_t)(hi[l] - lo[l] + 1) * m * m, poison ? NAN : 0.0f);
    grid[l] = grids[l].data();
  }
  grid[depth] = slab;
  for (size_t q = 0; q < grids[1].size(); q++) grids[1][q] = 0.0f;  // the 27 undisplaced points
  int first = 1;
  if (vol_coarse_fused(depth, lo, hi)) {  // rn_volume.cu: k_vol_coarse, every level as one whole-volume tile
    for (int l = 1; l <= kVolFusedLevels; l++) {
      int tiles_i = 0;
      VolLevel L = vol_level_make(depth, l, lo, hi, grid[l], grid[l + 1], seed, amplitude, &tiles_i);
      L.tile_i0 = 0;
      if (l == 1) vol_level_host<2, 2, 2>(L, 1, poison);
      else if (l == 2) vol_level_host<4, 4, 4>(L, 1, poison);
      else vol_level_host<8, 8, 8>(L, 1, poison);
    }
    first = kVolFusedLevels + 1;
  }
  for (int l = first; l < depth; l++) {
    const int n = 1 << l;
    int tiles_i = 0;
    const VolLevel L = vol_level_make(depth, l, lo, hi, grid[l], grid[l + 1], seed, amplitude, &tiles_i);
    if (n >= kRowBigN) vol_rows_host<kRowTI, kRowTJ, kRowWarps>(L, tiles_i, poison);
    else if (n >= kRowMinN) vol_rows_host<kRowSmallT, kRowSmallT, kRowWarps>(L, tiles_i, poison);
    else switch (n) {
      case 2: vol_level_host<2, 2, 2>(L, tiles_i, poison); break;
      case 4: vol_level_host<4, 4, 4>(L, tiles_i, poison); break;
#if RN_VOL_COARSE_TILE == 4
      default: vol_level_host<4, 4, 4>(L, tiles_i, poison); break;
#else
      case 8: vol_level_host<8, 8, 8>(L, tiles_i, poison); break;
      default: vol_level_host<kVolTI, kVolTJ, kVolTK>(L, tiles_i, poison); break;
#endif
    }
  }
  return 0;
}

// Fields per chunk of a batch (rn_plan.h march_batch_chunk, what rn_terrain2d_batch_chunk reports on the device side).
int replay_batch_chunk(int depth, int cmax, int sm_count) { return march_batch_chunk(depth, cmax, sm_count); }

// kinds[k] of the launches that synthesize a whole depth-`depth` field (0 tile stage, 1 block streaming, 2 warp march).
int replay_chain_kinds(int depth, int march, int* kinds)
{
  std::vector<StageLaunch> chain(kMaxLaunches);
  Law law{};
  const int side = (1 << depth) + 1;
  std::vector<char> dummy(256);
  const int ns = build_chain(chain.data(), depth, 0, 0, side, side, nullptr, (size_t)side, 0, dummy.data(), 1, law, 0.5f, 1u, nullptr, 1, 0.2f, 148, march);
  for (int k = 0; k < ns; k++) kinds[k] = chain[k].kind;
  return ns;
}

// Same contract as rn_terrain2d with a host destination (window may be NULL).  poison bit 0: NaN-fill
// shared memory and scratch; bit 1: run every tile through the EDGE instantiation; bit 2: visit the
// threads of every phase in descending order; bit 3: streamed levels through the round-1 block kernel (rn_stream.cuh) instead of
// the warp march.  Returns 0 on
// success, 100 if a tile wrote outside its shared memory.
int replay_terrain2d(const rn_params* p, const rn_window* win, rn_element* dst, size_t pitch_elems, int poison)
{
  if (!p || !dst || p->depth < RN_MIN_DEPTH || p->depth > RN_MAX_DEPTH) return RN_ERR_BAD_PARAM;
  const int side = (1 << p->depth) + 1;
  rn_window w{0, 0, side, side};
  if (win) w = *win;
  if (pitch_elems == 0) pitch_elems = (size_t)w.nj;
  g_reverse = (poison >> 2) & 1;
  Law law;
  law.amplitude = p->amplitude;
  law.sideways_scale = p->sideways_scale;
  law.radiosity_strength = p->radiosity_strength;
  law.amplitude_variation = p->amplitude_variation;
  law.band_hi = 3.0f / p->feature_scale;
  law.band_lo = 0.5f / p->feature_scale;
  law.bulgy = p->bulgy_rocks != 0;
  law.one = 1.0f;
  law.flat_ok = law_flat_ok(law);
  const size_t sbytes = chain_scratch_bytes(p->depth, 1);
  void* scratch = nullptr;
  if (sbytes && posix_memalign(&scratch, 256, sbytes) != 0) return RN_ERR_OOM;
  if (sbytes && poison) memset(scratch, 0xff, sbytes);  // NaN: untouched intermediate samples must never be consumed
  std::vector<StageLaunch> chainv(kMaxLaunches);
  StageLaunch* chain = chainv.data();
  const int ns = build_chain(chain, p->depth, w.i0, w.j0, w.ni, w.nj, reinterpret_cast<E4*>(dst), pitch_elems, 0, scratch, 1, law,
                             p->persistence, p->seed, nullptr, p->apply_volcano, p->amplitude, 148, (poison >> 3) & 1 ? 0 : RN_MARCH);
  int bad = 0;
  for (int k = 0; k < ns; k++) {
    if (chain[k].kind == 2) {
      bad |= run_march(chain[k], 1, nullptr, poison & 1, (poison >> 1) & 1);
      continue;
    }
    if (chain[k].kind == 1) {
      bad |= run_stream(chain[k], 1, nullptr, poison & 1, (poison >> 1) & 1);
      continue;
    }
    switch (chain[k].L) {
      case 1: bad |= run_stage<1>(chain[k], 1, nullptr, poison & 1, (poison >> 1) & 1); break;
      case 2: bad |= run_stage<2>(chain[k], 1, nullptr, poison & 1, (poison >> 1) & 1); break;
      case 3: bad |= run_stage<3>(chain[k], 1, nullptr, poison & 1, (poison >> 1) & 1); break;
      default: bad |= run_stage<4>(chain[k], 1, nullptr, poison & 1, (poison >> 1) & 1); break;
    }
  }
  free(scratch);
  return bad ? 100 : 0;
}

}  // extern "C"
