// rn_plan.h -- host-side planning for the fused tile path (algorithm 2).
//
// A depth-D field (Terrain.cpp:60-82: D-2 displacement levels, steps W/4 .. 2) is produced by
// a short chain of STAGES.  Stage k refines a complete depth-din grid into the complete
// depth-(din+L) grid by running L consecutive levels inside shared memory, tile by tile
// (rn_tile.cuh).  Levels 0..d-3 of a depth-D field are exactly the depth-d field at stride
// 2^(D-d) (the initial plane, the step/maxWidth ratio and the rand() call indices of a level
// depend only on the level), so every intermediate grid is itself a valid coarser field.
//
// Plain C++ (no CUDA) so tests/host_replay can share it.
#pragma once
#include <stdint.h>
#include "rn_math.cuh"
#include "rn_tile.cuh"
#include "rn_stream.cuh"
#include "rn_march.cuh"

namespace rn {

constexpr int kMaxStages = 14;
constexpr int kMaxLaunches = 16;
#ifndef RN_STREAM_THREADS
#define RN_STREAM_THREADS 128
#endif
#ifndef RN_STREAM_MINBLOCKS
#define RN_STREAM_MINBLOCKS 8
#endif
constexpr int kStreamThreads = RN_STREAM_THREADS;      // threads (= cell columns incl. 3 halo / loader columns) of a streaming CTA
constexpr int kStreamMinBlocks = RN_STREAM_MINBLOCKS;  // resident CTAs per SM the kernel is compiled for
#ifndef RN_STREAM_UNROLL
#define RN_STREAM_UNROLL 2   // steps per trip of the streaming march: 2 = rings of two are immediates (measured best: the
                             // 4-step form, all ring slots immediate, outgrows the instruction cache), 1 = rolled loop
#endif
#ifndef RN_STREAM_LEVELS
#define RN_STREAM_LEVELS 2
#endif
constexpr int kStreamLevels = RN_STREAM_LEVELS;  // the last levels of a field are streamed (rn_stream.cuh) ...
constexpr int kStreamMinDepth = 10;   // ... when the field is at least this deep
constexpr int kStreamThreadsSmall = 64;   // narrower strips for small levels (more CTAs, fewer warm-up rows per sample)
constexpr int kStreamMinBlocksSmall = 14;
#ifndef RN_STREAM_EDGE_COST
#define RN_STREAM_EDGE_COST 1.28f
#endif
constexpr float kStreamEdgeCost = RN_STREAM_EDGE_COST;  // duration of a border-strip CTA relative to an interior one with the same rows
#ifndef RN_STREAM_MIN_SEG_ROWS
#define RN_STREAM_MIN_SEG_ROWS 8
#endif
constexpr int kStreamMinSegRows = RN_STREAM_MIN_SEG_ROWS;  // shortest row segment of a streaming CTA (2 warm-up rows on top; 6, 10, 12, 16 measured worse)
#ifndef RN_STREAM_SMALL_COLS
#define RN_STREAM_SMALL_COLS 1024
#endif
constexpr int kStreamSmallCols = RN_STREAM_SMALL_COLS;    // levels with fewer cell columns than this use the small configuration
// Warp-synchronous march (rn_march.cuh), the streamed levels' kernel since round 2: one warp per (strip of 60 cell columns, row
// segment); the launch is one wave of about kMarchWarpsPerSM independent warps per SM.
#ifndef RN_MARCH
#define RN_MARCH 1              // 0: streamed levels run the round-1 block kernel (rn_stream.cuh)
#endif
#ifndef RN_MARCH_WARPS
#define RN_MARCH_WARPS 14       // resident warps per SM the kernel is compiled for (register cap 65536 / 32 / this)
#endif
#ifndef RN_MARCH_WPB
#define RN_MARCH_WPB 1          // independent warps per CTA
#endif
#ifndef RN_MARCH_FREE
#define RN_MARCH_FREE 2         // warps per SM the output launch leaves to the next call's small coarse stages
#endif
#ifndef RN_MARCH_FREE_COARSE
#define RN_MARCH_FREE_COARSE 2   // the same for the streamed levels before the last one (12 warps per SM = 3 per scheduler: 1025^2 -> 2049^2 in 30.2 instead of 32.4 us)
#endif
#ifndef RN_MARCH_FREE_ALWAYS
#define RN_MARCH_FREE_ALWAYS 0   // the same without a small chain (batches, depth 10): measured worse (64 depth-10 tiles: output level 438 vs 413-420 us)
#endif
#ifndef RN_MARCH_EDGE_COST
#define RN_MARCH_EDGE_COST 1.22f // duration of a border-strip warp relative to an interior one with the same rows (unit trace: 2.0 vs
                                 // 1.66 us per row); at depth 12 it gives 1772 units of 41 (border: 34) rows for the 1776 warp slots
#endif
#ifndef RN_MARCH_MIN_SEG_ROWS
#define RN_MARCH_MIN_SEG_ROWS 8
#endif
constexpr int kMarchWarpsPerSM = RN_MARCH_WARPS, kMarchWPB = RN_MARCH_WPB;
constexpr int kTile = 64;  // output tile edge of a stage (and the unit in which streamed stages are cut out of a window)
// Small chain (RN_SMALL_CHAIN): for one deep field per call the coarse levels run as stages of RN_SMALL_CHAIN_L levels with 32x32
// tiles and 128-thread CTAs (<= 64 registers, 30.9 KB) that fit into ONE CTA slot of the streaming output kernel, which leaves one
// slot per SM free: the coarse chain of the next call then runs underneath the output kernel of this one instead of after it
// (measured with profiles/micro/timeline.py: 133 vs 143 us per depth-12 field; the 512-thread tile CTAs cannot become resident
// while the output wave holds the slots).
#ifndef RN_SMALL_CHAIN
#define RN_SMALL_CHAIN 1
#endif
#ifndef RN_SMALL_CHAIN_MIN
#define RN_SMALL_CHAIN_MIN 11
#endif
#ifndef RN_SMALL_CHAIN_MAX
#define RN_SMALL_CHAIN_MAX 14
#endif
#ifndef RN_SMALL_CHAIN_FREE
#define RN_SMALL_CHAIN_FREE 1   // CTA slots per SM the output kernel leaves to the next call's small stages
#endif
#ifndef RN_SMALL_CHAIN_L
#define RN_SMALL_CHAIN_L 2   // levels per small stage: TileGeo<2,32> (30.9 KB) still fits beside 7 streaming CTAs per SM
#endif
constexpr int kTileSmall = 32;
constexpr int kStageThreadsBig = 512, kStageThreadsSmall = 128;

struct StagePlan {
  int din;  // depth of the input grid (2 = the analytic 5x5 initial lattice, Terrain.cpp:46-56)
  int L;    // levels fused in this stage (1..4)
};

// Split the depth-2 levels into stages.  The last stage -- the one that touches every output
// sample -- fuses kLastStageLevels levels (one in-place level + the unit-stride last level: the
// in-place levels are the latency-bound part of a tile, and the coarser grid it then starts from
// is 1/16 of the output and stays in L2); every stage before it fuses 3, and the first stage
// (its input is the analytic initial lattice) takes the remaining 1..3.
constexpr int kLastStageLevels = 2;
inline int plan_tile_levels(int NL, int last, StagePlan* out)
{
  int ns = 0;
  if (NL <= 4) {
    out[ns++] = StagePlan{2, NL};
    return ns;
  }
  const int rest = NL - last;                   // >= 3 for every caller
  const int first = (rest - 1) % 3 + 1;         // 1..3
  const int k = (rest - first) / 3;             // stages of 3 levels in between
  int d = 2;
  out[ns++] = StagePlan{d, first};
  d += first;
  for (int q = 0; q < k; q++) {
    out[ns++] = StagePlan{d, 3};
    d += 3;
  }
  if (last > 0) out[ns++] = StagePlan{d, last};
  return ns;
}

// Deep fields: tile stages up to depth - kStreamLevels, then one single-level stage per remaining
// level; the interior of those is streamed (rn_stream.cuh), their frame of border tiles goes through
// the L = 1 tile kernel.
inline bool plan_uses_streaming(int depth) { return depth >= kStreamMinDepth; }
inline bool plan_small_chain(int depth, int nfields) { return RN_SMALL_CHAIN && depth >= RN_SMALL_CHAIN_MIN && depth <= RN_SMALL_CHAIN_MAX && nfields <= 1; }
inline int plan_stages(int depth, StagePlan* out, int nfields = 0)
{
  if (plan_small_chain(depth, nfields)) {
    int ns = 0;
    for (int d = 2; d < depth - kStreamLevels; d += RN_SMALL_CHAIN_L) {
      const int rest = depth - kStreamLevels - d;
      out[ns++] = StagePlan{d, rest < RN_SMALL_CHAIN_L ? rest : RN_SMALL_CHAIN_L};
    }
    for (int q = kStreamLevels; q > 0; q--) out[ns++] = StagePlan{depth - q, 1};
    return ns;
  }
  if (!plan_uses_streaming(depth)) return plan_tile_levels(depth - 2, kLastStageLevels, out);
  int ns = plan_tile_levels(depth - 2 - kStreamLevels, 0, out);
  for (int q = kStreamLevels; q > 0; q--) out[ns++] = StagePlan{depth - q, 1};
  return ns;
}

// Per-level constants of a stage (level index = din-2+m, m = 0..L-1).
inline void fill_stage_levels(StageParams& sp, float persistence)
{
  for (int m = 0; m < sp.dout - sp.din; m++) {
    const int level = sp.din - 2 + m;
    uint64_t base = 0;
    for (int l = 0; l < level; l++) {
      const uint64_t n = 4ull << l;
      base += 3 * n * n + 2 * n;
    }
    const uint64_t n = 4ull << level;
    const float gain = level_gain(persistence, level);
    StageLevel& lv = sp.lev[m];
    lv.n = (int)n;
    lv.kS = (uint32_t)(base + 1);
    lv.kD1 = (uint32_t)(base + n * n + 1);
    lv.kD2 = (uint32_t)(base + n * n + n * (n + 1) + 1);
    lv.blahS = (1.414f * 2.0f) * gain;   // Terrain.cpp:231
    lv.blahD = 2.0f * (gain * 1.0f);     // Terrain.cpp:186, X extent not clamped
    lv.blahDh = 2.0f * (gain * 0.5f);    // Terrain.cpp:186, X extent clamped (halved)
  }
}

// Inclusive rectangle of lattice coordinates of some grid.
struct Rect {
  int i0, i1, j0, j1;
};

// Tiles of a depth-d grid: tile t owns rows [t*T, t*T+T-1]; the last tile also owns row W.
inline int tiles_per_side(int d, int T = kTile) { const int W = 1 << d; return W >= T ? W / T : 1; }
inline int tile_of(int d, int coord, int T = kTile) { const int t = coord / T, nt = tiles_per_side(d, T); return t < nt ? t : nt - 1; }

// Input-grid rectangle (depth d-L) a stage needs in order to produce tiles [t0,t1] of its output grid.
inline void input_range_1d(int d, int L, int t0, int t1, int T, int* lo, int* hi)
{
  const int W = 1 << d;
  const int H = stage_halo(L);
  int a = t0 * T - H, b = (t1 + 1) * T + H;
  if (a < 0) a = 0;
  if (b > W) b = W;
  *lo = a >> L;
  *hi = (b + (1 << L) - 1) >> L;
  if (*hi > (W >> L)) *hi = W >> L;
}

// One launch of the chain.
struct StageLaunch {
  int kind;               // 0 = tile kernel k_stage<L>, 1 = streaming block kernel k_stream, 2 = warp-synchronous k_march
  int stage;              // index of the stage this launch belongs to
  int last;               // belongs to the last stage (runs on the caller's stream)
  StageParams sp;         // kind 0
  StreamParams st;        // kind 1
  int L;
  int T;                  // kind 0: tile edge (kTile, or kTileSmall with kStageThreadsSmall threads)
  int nt;                 // kind 1: threads per CTA (kStreamThreads or kStreamThreadsSmall)
  unsigned gx, gy;        // kind 0: tiles in j, tiles in i; kind 1: strips, segments
};

// Intermediate grids have an EVEN row pitch (side + 1 elements) and 256-byte aligned bases, so that rn_march.cuh can read a
// lane's two lattice columns with one 256-bit load.
inline size_t stage_grid_pitch(int d) { return ((size_t)1 << d) + 2; }
inline size_t stage_grid_bytes(int d) { const size_t side = ((size_t)1 << d) + 1; return ((side * stage_grid_pitch(d) * sizeof(E4) + 255) / 256) * 256; }

inline size_t chain_scratch_bytes(int depth, int nfields)
{
  StagePlan plan[kMaxStages];
  const int ns = plan_stages(depth, plan, nfields);
  size_t total = 0;
  for (int k = 0; k + 1 < ns; k++) total += stage_grid_bytes(plan[k].din + plan[k].L) * (size_t)nfields;
  return total;
}

// Row segments of a streamed level: the smallest interior segment length r (>= min_rows) whose unit count -- ni interior strips
// and ne border strips per field -- fits per_field units; border strips get r / edge_cost rows.
inline void segment_rows(int per_field, int rows, int ni, int ne, int min_rows, float edge_cost, int* r_out, int* re_out)
{
  int r = min_rows < rows ? min_rows : rows, re = r;
  for (;; r++) {
    re = (int)((float)r / edge_cost + 0.5f);
    re = re < min_rows ? (min_rows < rows ? min_rows : rows) : re;
    if (r >= rows || ni * ((rows + r - 1) / r) + ne * ((rows + re - 1) / re) <= per_field) break;
  }
  *r_out = r;
  *re_out = re;
}

// How many whole fields a batch chunk of the warp march should hold (at most cmax): every field of a chunk gets the same whole number
// of (strip, segment) units, so the duration of the output level per field is a sawtooth in the chunk size -- 229 depth-10 fields
// x 9 full-height units fill the warp slots, but the two border strips of every field then run 1.22x as long as the rest; 103
// fields x 20 units (2 segments per interior strip, 3 per border strip) finish together.  Model: the longest unit of the chunk
// (rows + 3 steps of start-up, border rows x edge cost) per field; the smallest chunk of at least 4 fields within 2 % of the best
// wins (less scratch).  Depth 11, 32 tiles: chunks of 2 / 4 / 8 / 14 / 28 give 102 / 123 / 134 / 135 / 131; depth 12, 8 tiles: 2 / 4 / 8: 143 / 149 / 146.  Measured, 1024 depth-10 tiles: chunks of 229 / 103 / 69 / 27 / 14 fields give 111 / 109 / 119 / 120 / 108 Gsamples/s.
inline int march_batch_chunk(int depth, int cmax, int sm_count)
{
  if (!RN_MARCH || !plan_uses_streaming(depth) || cmax <= 1) return cmax;
  const int n = 1 << (depth - 1);
  const int nstrips = (n + 1 + kMarchCols - 1) / kMarchCols;
  const int ne = nstrips > 1 ? 2 : 1, ni = nstrips - ne;
  const int target = sm_count * (kMarchWarpsPerSM - (RN_MARCH_FREE_ALWAYS ? RN_MARCH_FREE : 0));
  float best = 0.0f;
  int best_c = cmax;
  for (int pass = 0; pass < 2; pass++)
    for (int c = cmax; c >= (cmax < 4 ? 1 : 4); c--) {
      int r, re;
      segment_rows(target / c > 0 ? target / c : 1, n, ni, ne, RN_MARCH_MIN_SEG_ROWS, RN_MARCH_EDGE_COST, &r, &re);
      const int units = ni * ((n + r - 1) / r) + ne * ((n + re - 1) / re);
      const float waves = (float)(units * c) / (float)target;  // > 1 when even whole strips do not fit
      const float longest = (float)(r + 3) > (float)(re + 3) * RN_MARCH_EDGE_COST ? (float)(r + 3) : (float)(re + 3) * RN_MARCH_EDGE_COST;
      const float cost = longest * (waves > 1.0f ? waves : 1.0f) / (float)c;
      if (pass == 0) { if (best == 0.0f || cost < best) best = cost; }
      else if (cost <= best * 1.02f && (c >= 4 || cmax < 4)) best_c = c;  // descending c: ends at the smallest chunk (of at least 4
                                                                            // fields: the fixed cost of a chunk) that is good enough
    }
  return best_c;
}

// Builds the launches that synthesize the window [i0,i0+ni) x [j0,j0+nj) of `nfields` depth-`depth`
// fields into dst (element (i0,j0) of field 0; row pitch / field stride in elements).  Earlier
// stages only compute the tiles the window's dependency cone touches.  Returns the launch count.
inline int build_chain(StageLaunch* out, int depth, int i0, int j0, int ni, int nj, E4* dst, size_t pitch, size_t field_stride,
                       void* scratch, int nfields, const Law& law, float persistence, uint32_t seed, const uint32_t* seeds, int volcano,
                       float amplitude, int sm_count = 148, int march = RN_MARCH)
{
  StagePlan plan[kMaxStages];
  const int ns = plan_stages(depth, plan, nfields);
  const bool streaming = plan_uses_streaming(depth);
  const bool small_chain = plan_small_chain(depth, nfields);
  int t_i0[kMaxStages], t_i1[kMaxStages], t_j0[kMaxStages], t_j1[kMaxStages], Tk[kMaxStages];
  for (int k = 0; k < ns; k++) Tk[k] = (small_chain && k < ns - kStreamLevels) ? kTileSmall : kTile;
  t_i0[ns - 1] = tile_of(depth, i0, Tk[ns - 1]);
  t_i1[ns - 1] = tile_of(depth, i0 + ni - 1, Tk[ns - 1]);
  t_j0[ns - 1] = tile_of(depth, j0, Tk[ns - 1]);
  t_j1[ns - 1] = tile_of(depth, j0 + nj - 1, Tk[ns - 1]);
  for (int k = ns - 1; k > 0; k--) {
    const int d = plan[k].din + plan[k].L;
    int lo, hi;
    input_range_1d(d, plan[k].L, t_i0[k], t_i1[k], Tk[k], &lo, &hi);
    t_i0[k - 1] = tile_of(plan[k].din, lo, Tk[k - 1]);
    t_i1[k - 1] = tile_of(plan[k].din, hi, Tk[k - 1]);
    input_range_1d(d, plan[k].L, t_j0[k], t_j1[k], Tk[k], &lo, &hi);
    t_j0[k - 1] = tile_of(plan[k].din, lo, Tk[k - 1]);
    t_j1[k - 1] = tile_of(plan[k].din, hi, Tk[k - 1]);
  }
  const E4* gin = nullptr;
  size_t gin_pitch = 0, gin_stride = 0, off = 0;
  int nl = 0;
  for (int k = 0; k < ns; k++) {
    StageParams sp{};
    sp.din = plan[k].din;
    sp.dout = plan[k].din + plan[k].L;
    sp.gin = gin;
    sp.gin_pitch = gin_pitch;
    sp.gin_field_stride = gin_stride;
    const size_t side = ((size_t)1 << sp.dout) + 1;
    const bool last = (k == ns - 1);
    if (last) {
      sp.gout = dst;
      sp.gout_pitch = pitch;
      sp.gout_field_stride = field_stride;
      sp.win_i0 = i0; sp.win_j0 = j0; sp.win_ni = ni; sp.win_nj = nj;
      sp.volcano = volcano;
      sp.volcano_gain = amplitude * 0.008f;  // Terrain.cpp:137
    } else {
      const size_t fbytes = stage_grid_bytes(sp.dout);
      sp.gout = reinterpret_cast<E4*>(static_cast<char*>(scratch) + off);
      sp.gout_pitch = stage_grid_pitch(sp.dout);
      sp.gout_field_stride = fbytes / sizeof(E4);
      off += fbytes * (size_t)nfields;
      sp.win_i0 = 0; sp.win_j0 = 0; sp.win_ni = (int)side; sp.win_nj = (int)side;
    }
    sp.tile_i0 = t_i0[k];
    sp.tile_j0 = t_j0[k];
    sp.seed = seed;
    sp.seeds = seeds;
    sp.law = law;
    fill_stage_levels(sp, persistence);
    // a streamed single-level stage: the whole needed rectangle goes through k_stream (no tile launch)
    if (streaming && plan[k].L == 1 && k >= ns - kStreamLevels) {
      const int nt = tiles_per_side(sp.dout);
      const int n = 1 << (sp.dout - 1);
      StageLaunch& sl = out[nl++];
      sl = StageLaunch{};
      sl.kind = 1; sl.stage = k; sl.last = last; sl.L = 1;
      StreamParams& st = sl.st;
      st.gin = sp.gin; st.gin_pitch = sp.gin_pitch; st.gin_field_stride = sp.gin_field_stride;
      st.gout = sp.gout; st.gout_pitch = sp.gout_pitch; st.gout_field_stride = sp.gout_field_stride;
      st.dout = sp.dout;
      st.win_i0 = sp.win_i0; st.win_j0 = sp.win_j0; st.win_ni = sp.win_ni; st.win_nj = sp.win_nj;
      st.a_lo = t_i0[k] * (kTile / 2); st.a_hi = (t_i1[k] + 1) * (kTile / 2) < n ? (t_i1[k] + 1) * (kTile / 2) : n;
      st.b_lo = t_j0[k] * (kTile / 2); st.b_hi = (t_j1[k] + 1) * (kTile / 2) < n ? (t_j1[k] + 1) * (kTile / 2) : n;
      st.close_i = t_i1[k] == nt - 1;
      st.close_j = t_j1[k] == nt - 1;
      const int cols = st.b_hi - st.b_lo, rows = st.a_hi - st.a_lo;
      int ctas_target, min_rows;
      float edge_cost;
      if (march) {
        // warp-synchronous march: strips of 60 cell columns (the closing output column counts as one)
        sl.kind = 2;
        sl.nt = 32 * kMarchWPB;
        st.strip_cols = kMarchCols;
        st.nstrips = (cols + st.close_j + kMarchCols - 1) / kMarchCols;
        auto strip_is_edge = [&](int s) {  // same rule as MarchCtx::col_lean
          const int B0 = st.b_lo + s * kMarchCols;
          const int CW = kMarchCols < st.b_hi - B0 ? kMarchCols : st.b_hi - B0;
          return B0 - 2 < 0 || B0 + 62 > n || (st.close_j && B0 + CW == st.b_hi);
        };
        st.edge_first = strip_is_edge(0);
        st.edge_last = st.nstrips > 1 && strip_is_edge(st.nstrips - 1);
        ctas_target = sm_count * (kMarchWarpsPerSM - ((small_chain || RN_MARCH_FREE_ALWAYS) ? (last ? RN_MARCH_FREE : RN_MARCH_FREE_COARSE) : 0));
        min_rows = RN_MARCH_MIN_SEG_ROWS;
        edge_cost = RN_MARCH_EDGE_COST;
      } else {
        sl.nt = cols < kStreamSmallCols ? kStreamThreadsSmall : kStreamThreads;
        ctas_target = sm_count * ((cols < kStreamSmallCols ? kStreamMinBlocksSmall : kStreamMinBlocks) - ((small_chain && last) ? RN_SMALL_CHAIN_FREE : 0));
        const int strips = (cols + (sl.nt - 3) - 1) / (sl.nt - 3);
        st.strip_cols = (cols + strips - 1) / strips;
        st.nstrips = (cols + st.strip_cols - 1) / st.strip_cols;
        // strips on a column border (same rule as StreamCtx::col_lean)
        st.edge_first = st.b_lo < 1 || (st.nstrips == 1 && (st.b_hi > n - 1 || st.close_j));
        st.edge_last = st.nstrips > 1 && (st.b_hi > n - 1 || st.close_j);
        min_rows = kStreamMinSegRows;
        edge_cost = kStreamEdgeCost;
      }
      const int ne = st.edge_first + st.edge_last, ni = st.nstrips - ne;
      // One wave: about ctas_target CTAs (warps) of equal duration.  A segment costs 2 warm-up steps, but small levels are
      // latency-bound: filling every slot wins (measured: 1025 -> 2049 level 40 us with 8-row segments, 52 us
      // with 24-row segments).  A border strip takes edge_cost times as long per row (measured), so it
      // gets that much fewer rows.
      const int per_field = ctas_target / (nfields > 0 ? nfields : 1);
      int r, re;
      segment_rows(per_field, rows, ni, ne, min_rows, edge_cost, &r, &re);
      st.seg_rows = r;
      st.nseg = (rows + r - 1) / r;
      st.seg_rows_edge = re;
      st.nseg_edge = (rows + re - 1) / re;
      st.seed = seed; st.seeds = seeds; st.law = law;
      st.lev = sp.lev[0];
      st.jn = jump_of((uint64_t)st.lev.n);
      st.jn1 = jump_of((uint64_t)st.lev.n + 1);
      st.volcano = sp.volcano; st.volcano_gain = sp.volcano_gain;
      sl.gx = (unsigned)stream_grid_size(st);  // kind 2: warps
      sl.gy = 1;
      gin = sp.gout;
      gin_pitch = sp.gout_pitch;
      gin_stride = sp.gout_field_stride;
      continue;
    }
    StageLaunch& sl = out[nl++];
    sl = StageLaunch{};
    sl.kind = 0; sl.stage = k; sl.last = last; sl.L = plan[k].L; sl.T = Tk[k];
    sl.sp = sp;
    sl.gx = (unsigned)(t_j1[k] - t_j0[k] + 1);
    sl.gy = (unsigned)(t_i1[k] - t_i0[k] + 1);
    gin = sp.gout;
    gin_pitch = sp.gout_pitch;
    gin_stride = sp.gout_field_stride;
  }
  return nl;
}

}  // namespace rn
