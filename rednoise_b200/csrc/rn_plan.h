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

namespace rn {

constexpr int kMaxStages = 8;
constexpr int kTile = 64;  // output tile edge of every stage

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
inline int plan_stages(int depth, StagePlan* out)
{
  const int NL = depth - 2;
  int ns = 0;
  if (NL <= 4) {
    out[ns++] = StagePlan{2, NL};
    return ns;
  }
  const int rest = NL - kLastStageLevels;       // >= 3
  const int first = (rest - 1) % 3 + 1;         // 1..3
  const int k = (rest - first) / 3;             // stages of 3 levels in between
  int d = 2;
  out[ns++] = StagePlan{d, first};
  d += first;
  for (int q = 0; q < k; q++) {
    out[ns++] = StagePlan{d, 3};
    d += 3;
  }
  out[ns++] = StagePlan{d, kLastStageLevels};
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
inline int tiles_per_side(int d) { const int W = 1 << d; return W >= kTile ? W / kTile : 1; }
inline int tile_of(int d, int coord) { const int t = coord / kTile, nt = tiles_per_side(d); return t < nt ? t : nt - 1; }

// Input-grid rectangle (depth d-L) a stage needs in order to produce tiles [t0,t1] of its output grid.
inline void input_range_1d(int d, int L, int t0, int t1, int* lo, int* hi)
{
  const int W = 1 << d;
  const int H = stage_halo(L);
  int a = t0 * kTile - H, b = (t1 + 1) * kTile + H;
  if (a < 0) a = 0;
  if (b > W) b = W;
  *lo = a >> L;
  *hi = (b + (1 << L) - 1) >> L;
  if (*hi > (W >> L)) *hi = W >> L;
}

// One launch of the chain: the stage parameters (pointers resolved), its tile grid and where
// its output lives (scratch offset in bytes, or the caller's destination for the last stage).
struct StageLaunch {
  StageParams sp;
  int L;
  unsigned gx, gy;        // tiles in j, tiles in i
  size_t scratch_offset;  // intermediate stages: byte offset of field 0 inside the scratch buffer
  size_t field_bytes;     // intermediate stages: bytes per field (256-byte aligned)
};

inline size_t stage_grid_bytes(int d) { const size_t side = ((size_t)1 << d) + 1; return ((side * side * sizeof(E4) + 255) / 256) * 256; }

inline size_t chain_scratch_bytes(int depth, int nfields)
{
  StagePlan plan[kMaxStages];
  const int ns = plan_stages(depth, plan);
  size_t total = 0;
  for (int k = 0; k + 1 < ns; k++) total += stage_grid_bytes(plan[k].din + plan[k].L) * (size_t)nfields;
  return total;
}

// Builds the launches that synthesize the window [i0,i0+ni) x [j0,j0+nj) of `nfields` depth-`depth`
// fields into dst (element (i0,j0) of field 0; row pitch / field stride in elements).  Earlier
// stages only compute the tiles the window's dependency cone touches.
inline int build_chain(StageLaunch* out, int depth, int i0, int j0, int ni, int nj, E4* dst, size_t pitch, size_t field_stride,
                       void* scratch, int nfields, const Law& law, float persistence, uint32_t seed, const uint32_t* seeds, int volcano,
                       float amplitude)
{
  StagePlan plan[kMaxStages];
  const int ns = plan_stages(depth, plan);
  int t_i0[kMaxStages], t_i1[kMaxStages], t_j0[kMaxStages], t_j1[kMaxStages];
  t_i0[ns - 1] = tile_of(depth, i0);
  t_i1[ns - 1] = tile_of(depth, i0 + ni - 1);
  t_j0[ns - 1] = tile_of(depth, j0);
  t_j1[ns - 1] = tile_of(depth, j0 + nj - 1);
  for (int k = ns - 1; k > 0; k--) {
    const int d = plan[k].din + plan[k].L;
    int lo, hi;
    input_range_1d(d, plan[k].L, t_i0[k], t_i1[k], &lo, &hi);
    t_i0[k - 1] = tile_of(plan[k].din, lo);
    t_i1[k - 1] = tile_of(plan[k].din, hi);
    input_range_1d(d, plan[k].L, t_j0[k], t_j1[k], &lo, &hi);
    t_j0[k - 1] = tile_of(plan[k].din, lo);
    t_j1[k - 1] = tile_of(plan[k].din, hi);
  }
  const E4* gin = nullptr;
  size_t gin_pitch = 0, gin_stride = 0, off = 0;
  for (int k = 0; k < ns; k++) {
    StageLaunch& sl = out[k];
    sl = StageLaunch{};
    StageParams& sp = sl.sp;
    sl.L = plan[k].L;
    sp.din = plan[k].din;
    sp.dout = plan[k].din + plan[k].L;
    sp.gin = gin;
    sp.gin_pitch = gin_pitch;
    sp.gin_field_stride = gin_stride;
    const size_t side = ((size_t)1 << sp.dout) + 1;
    if (k == ns - 1) {
      sp.gout = dst;
      sp.gout_pitch = pitch;
      sp.gout_field_stride = field_stride;
      sp.win_i0 = i0; sp.win_j0 = j0; sp.win_ni = ni; sp.win_nj = nj;
      sp.volcano = volcano;
      sp.volcano_gain = amplitude * 0.008f;  // Terrain.cpp:137
    } else {
      sl.scratch_offset = off;
      sl.field_bytes = stage_grid_bytes(sp.dout);
      sp.gout = reinterpret_cast<E4*>(static_cast<char*>(scratch) + off);
      sp.gout_pitch = side;
      sp.gout_field_stride = sl.field_bytes / sizeof(E4);
      off += sl.field_bytes * (size_t)nfields;
      sp.win_i0 = 0; sp.win_j0 = 0; sp.win_ni = (int)side; sp.win_nj = (int)side;
    }
    sp.tile_i0 = t_i0[k];
    sp.tile_j0 = t_j0[k];
    sp.seed = seed;
    sp.seeds = seeds;
    sp.law = law;
    fill_stage_levels(sp, persistence);
    sl.gx = (unsigned)(t_j1[k] - t_j0[k] + 1);
    sl.gy = (unsigned)(t_i1[k] - t_i0[k] + 1);
    gin = sp.gout;
    gin_pitch = sp.gout_pitch;
    gin_stride = sp.gout_field_stride;
  }
  return ns;
}

}  // namespace rn
