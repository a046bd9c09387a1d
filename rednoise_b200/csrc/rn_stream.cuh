// rn_stream.cuh -- algorithm 2, streaming form of ONE level of the diamond-square path
// (Terrain.cpp:60-82 for a single `step`), used for the last (largest) levels of a field.
//
// A CTA owns a column strip of the output and MARCHES down it, one row of cells per step, keeping
// only a few rolling rows of every sample class in shared memory:
//   * nothing is recomputed in the marching direction (a tile recomputes its halo on all 4 sides),
//   * every thread owns one cell column and has one cell of work in every barrier interval, so
//     there are no low-parallelism phases, no per-tile start-up and no index arithmetic,
//   * the LCG state of a column advances by a constant jump per row (one IMAD, no tables).
// The level's input is the complete coarser grid in global memory (L2-resident: it is 1/4 of the
// level's output); the output is stored as full 128-bit coalesced rows, like the tile kernels.
//
// Only INTERIOR regions are streamed (every cell the strip touches exists); the frame of 64-sample
// tiles along the lattice border is produced by the tile kernel (rn_tile.cuh) in the same stage.
//
// Step u of a strip (cell row u; cell (u,c) has its corner at output (2u, 2c')):
//   interval 1   S(u): centre C(u,c) from lattice rows u, u+1        (averageSquare+generateSquare)
//                + store of the two output rows finished by G(u-1)
//   interval 2   D(u): mids M1(u,c) and M2(u-1,c)                     (averageDiamond+generateDiamond)
//                + global load of lattice row u+2 into registers
//   interval 3   G(u): final shade of corner K(u-1,c) and centre C(u-1,c) (ordered gather)
//                + the prefetched lattice row goes to shared memory
// Rolling rows: lattice x4, centres x2, M1 x2, M2, final K, final C (float4) and the loss planes
// TS x3, TD1 x2, TD2 x2 (float).  See rn_tile.cuh for the vocabulary and the shade-order argument.
//
// Compiles as plain C++ for tests/host_replay (per-thread registers live in StreamThread).
#pragma once
#include "rn_tile.cuh"

namespace rn {

struct StreamParams {
  const E4* gin;   // complete depth-(dout-1) grid: lattice point (r, k) <-> output sample (2r, 2k)
  size_t gin_pitch, gin_field_stride;
  E4* gout;        // element (win_i0, win_j0) of the depth-dout grid
  size_t gout_pitch, gout_field_stride;
  int dout;
  int win_i0, win_j0, win_ni, win_nj;
  int a_lo, a_hi;  // owned cell rows [a_lo, a_hi) of the streamed region
  int b_lo, b_hi;  // owned cell columns
  int seg_rows;    // cell rows per CTA (blockIdx.y)
  int strip_cols;  // owned cell columns per CTA (blockIdx.x), <= NT-2
  uint32_t seed;
  const uint32_t* seeds;
  Law law;
  StageLevel lev;
  Jump jn, jn1;    // LCG jumps over n and n+1 draws (one cell row of generateSquare / loop 1, of loop 2)
  int volcano;
  float volcano_gain;
};

// registers of one thread that live across barrier intervals
struct StreamThread {
  uint32_t sS, sD1, sD2;  // LCG states whose extraction is the draw of this column's next S / loop-1 / loop-2 call
  E4 pre, pre_x;          // prefetched lattice point(s) of row u+2 (pre_x: the extra last column, thread 0)
  int have_pre;
};

struct StreamCtx {
  int B0, CW;   // first owned cell column, owned columns of this strip
  int a0, a1;   // owned cell rows [a0, a1)
  uint32_t seed;
  const E4* gin;
  E4* gout;
};

template <int NT>
struct StreamGeo {
  static constexpr int P = NT + 2;  // row length: c = -1 .. NT
  // E4 rows
  static constexpr int KROW = 0;                 // 4 lattice rows
  static constexpr int CROW = KROW + 4 * P;      // 2 centre rows
  static constexpr int M1ROW = CROW + 2 * P;     // 2 loop-1 mid rows
  static constexpr int M2ROW = M1ROW + 2 * P;    // 1 loop-2 mid row
  static constexpr int OUTK = M2ROW + P;         // final corners of the row being stored
  static constexpr int OUTC = OUTK + P;          // final centres
  static constexpr int E4ROWS = OUTC + P;
  // float rows (word offsets)
  static constexpr int TS = 4 * E4ROWS;          // 3 rows
  static constexpr int TD1 = TS + 3 * P;         // 2 rows
  static constexpr int TD2 = TD1 + 2 * P;        // 2 rows
  static constexpr int WORDS = ((TD2 + 2 * P + 3) / 4) * 4;
  static constexpr size_t BYTES = (size_t)WORDS * 4;
};

template <int NT>
struct Stream {
  using G = StreamGeo<NT>;
  static RN_HD E4* e4(float* sm, int row_off, int c) { return reinterpret_cast<E4*>(sm) + row_off + (c + 1); }
  static RN_HD float* f32(float* sm, int row_off, int c) { return sm + row_off + (c + 1); }
  static RN_HD int mod(int u, int m) { return ((u % m) + m) % m; }

  static RN_HD bool active(const StreamCtx& cx, int c) { return c <= cx.CW + 1; }

  // lattice point (row r, cell column index c) from the coarser grid
  static RN_HD E4 lattice(const StreamParams& sp, const StreamCtx& cx, int r, int c)
  {
    return cx.gin[(size_t)r * sp.gin_pitch + (size_t)(cx.B0 - 1 + c)];
  }

  // ------------------------------------------------------------------ start-up ----
  static RN_HD void init(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int c)
  {
    const int u0 = cx.a0 - 1;
    const uint32_t n = (uint32_t)sp.lev.n;
    const uint32_t bcol = (uint32_t)(cx.B0 - 1 + c);
    th.sS = lcg_skip(cx.seed, sp.lev.kS + (uint32_t)u0 * n + bcol);
    th.sD1 = lcg_skip(cx.seed, sp.lev.kD1 + (uint32_t)u0 * n + bcol);
    th.sD2 = lcg_skip(cx.seed, sp.lev.kD2 + (uint32_t)(u0 - 1) * (n + 1u) + bcol);
    th.have_pre = 0;
    if (active(cx, c)) {
      *e4(sm, G::KROW + mod(u0, 4) * G::P, c) = lattice(sp, cx, u0, c);
      *e4(sm, G::KROW + mod(u0 + 1, 4) * G::P, c) = lattice(sp, cx, u0 + 1, c);
      if (c == 0) {  // the column to the right of the last cell
        *e4(sm, G::KROW + mod(u0, 4) * G::P, cx.CW + 2) = lattice(sp, cx, u0, cx.CW + 2);
        *e4(sm, G::KROW + mod(u0 + 1, 4) * G::P, cx.CW + 2) = lattice(sp, cx, u0 + 1, cx.CW + 2);
      }
    }
  }

  // ------------------------------------------------------------------ interval 1 ----
  static RN_HD void store_rows(const StreamParams& sp, const StreamCtx& cx, float* sm, int a, int c)
  {
    // output rows 2a (corners K(a,.) / mids M1(a,.)) and 2a+1 (mids M2(a,.) / centres C(a,.)),
    // columns 2*B0 .. 2*(B0+CW)-1.  Lane -> column permutation inside every 32-column group: each
    // quarter warp reads 8 consecutive slots of one class row, the warp stores 512 contiguous bytes.
    const int Wout = 1 << sp.dout;
    const float inv_w = 1.0f / (float)Wout;
    const int lane = c & 31, q = lane >> 3, k = lane & 7;
    const int c32 = (q & 1) * 16 + 2 * k + (q >> 1);
    const int m1slot = mod(a, 2);
    for (int pass = 0; pass < 2; pass++) {
      const int j = pass * NT + (c - lane) + c32;  // column relative to 2*B0
      if (j >= 2 * cx.CW) continue;
      const int cc = (j >> 1) + 1;                 // thread index of the cell that owns the sample
      const int gj = 2 * cx.B0 + j;
      const int wj = gj - sp.win_j0;
      if ((unsigned)wj >= (unsigned)sp.win_nj) continue;
      const float fy = (float)(gj - Wout / 2);
      const float fy2 = fy * fy;
      for (int r = 0; r < 2; r++) {
        const int gi = 2 * a + r;
        const int wi = gi - sp.win_i0;
        if ((unsigned)wi >= (unsigned)sp.win_ni) continue;
        int row;
        if (r == 0) row = (j & 1) ? G::M1ROW + m1slot * G::P : G::OUTK;
        else row = (j & 1) ? G::OUTC : G::M2ROW;
        E4 e = *e4(sm, row, cc);
        if (sp.volcano) {
          const float fx = (float)(gi - Wout / 2);
          e.y = e.y + volcano_height_r2<false>(fx * fx + fy2, inv_w, sp.volcano_gain);
        }
        cx.gout[(size_t)wi * sp.gout_pitch + (size_t)wj] = e;
      }
    }
  }

  static RN_HD void interval1(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int u, int c)
  {
    if (active(cx, c)) {
      // S(u): averageSquare (Terrain.cpp:216-222) + generateSquare (:224-258)
      const int r0 = G::KROW + mod(u, 4) * G::P, r1 = G::KROW + mod(u + 1, 4) * G::P;
      const E4 k00 = *e4(sm, r0, c), k01 = *e4(sm, r0, c + 1), k10 = *e4(sm, r1, c), k11 = *e4(sm, r1, c + 1);
      const V3 avg = vmul(vadd(vadd(vadd(pos4(k00), pos4(k10)), pos4(k01)), pos4(k11)), 0.25f);
      const float sh = (k00.w + k10.w + k01.w + k11.w) * 0.25f;
      const uint32_t draw = lcg_draw_from_state(th.sS);
      th.sS = jump_apply(sp.jn, th.sS);
      const Displaced dp = displace<true>(sp.law, avg, vsub(pos4(k11), pos4(k00)), vsub(pos4(k10), pos4(k01)), sp.lev.blahS, draw);
      *e4(sm, G::CROW + mod(u, 2) * G::P, c) = e4_of(dp.pos, sh + shade_gain(sp.law, dp.offset));
      *f32(sm, G::TS + mod(u, 3) * G::P, c) = shade_loss(sp.law, dp.offset);
    }
    // the rows G(u-1) finished: cell row u-2
    if (u - 2 >= cx.a0 && u - 2 < cx.a1) store_rows(sp, cx, sm, u - 2, c);
  }

  // ------------------------------------------------------------------ interval 2 ----
  static RN_HD E4 mid(const Law& law, const E4& ex0, const E4& ex1, const E4& ey0, const E4& ey1, float blah, uint32_t draw, float* loss_out)
  {
    const V3 avg = vmul(vadd(vadd(vadd(pos4(ex0), pos4(ex1)), pos4(ey0)), pos4(ey1)), 0.25f);
    const float sh = (ex0.w + ex1.w + ey0.w + ey1.w) * 0.25f;
    const Displaced dp = displace<true>(law, avg, vsub(pos4(ex0), pos4(ex1)), vsub(pos4(ey1), pos4(ey0)), blah, draw);
    *loss_out = shade_loss(law, dp.offset);
    return e4_of(dp.pos, sh + shade_gain(law, dp.offset));
  }

  static RN_HD void interval2(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int u, int c)
  {
    // prefetch lattice row u+2 (needed by S(u+1)); it lands in shared memory in interval 3
    th.have_pre = 0;
    if (active(cx, c) && u + 1 <= cx.a1) {
      th.pre = lattice(sp, cx, u + 2, c);
      if (c == 0) th.pre_x = lattice(sp, cx, u + 2, cx.CW + 2);
      th.have_pre = 1;
    }
    if (!active(cx, c)) return;
    const int ts2 = G::TS + mod(u - 2, 3) * G::P, ts1 = G::TS + mod(u - 1, 3) * G::P, ts0 = G::TS + mod(u, 3) * G::P;
    const int kr = G::KROW + mod(u, 4) * G::P, krm = G::KROW + mod(u - 1, 4) * G::P;
    const int cr = G::CROW + mod(u, 2) * G::P, crm = G::CROW + mod(u - 1, 2) * G::P;
    // corner shades as generateSquare left them (Terrain.cpp:252-255): minus the scatter of the 4 surrounding centres
    const float t1m = *f32(sm, ts1, c - 1), t10 = *f32(sm, ts1, c), t0m = *f32(sm, ts0, c - 1), t00 = *f32(sm, ts0, c);
    E4 k00 = *e4(sm, kr, c);  // K(u, c)
    k00.w = (((k00.w - t1m) - t10) - t0m) - t00;
    {  // loop 1: M1(u, c) at output (2u, 2c'+1): x neighbours C(u-1,c), C(u,c); y neighbours K(u,c), K(u,c+1)
      E4 k01 = *e4(sm, kr, c + 1);
      k01.w = (((k01.w - t10) - *f32(sm, ts1, c + 1)) - t00) - *f32(sm, ts0, c + 1);
      const uint32_t draw = lcg_draw_from_state(th.sD1);
      th.sD1 = jump_apply(sp.jn, th.sD1);
      float loss;
      const E4 m = mid(sp.law, *e4(sm, crm, c), *e4(sm, cr, c), k00, k01, sp.lev.blahD, draw, &loss);
      *e4(sm, G::M1ROW + mod(u, 2) * G::P, c) = m;
      *f32(sm, G::TD1 + mod(u, 2) * G::P, c) = loss;
    }
    {  // loop 2: M2(u-1, c) at output (2u-1, 2c'): x neighbours K(u-1,c), K(u,c); y neighbours C(u-1,c-1), C(u-1,c)
      E4 km = *e4(sm, krm, c);  // K(u-1, c)
      km.w = (((km.w - *f32(sm, ts2, c - 1)) - *f32(sm, ts2, c)) - t1m) - t10;
      const uint32_t draw = lcg_draw_from_state(th.sD2);
      th.sD2 = jump_apply(sp.jn1, th.sD2);
      float loss;
      const E4 m = mid(sp.law, km, k00, *e4(sm, crm, c - 1), *e4(sm, crm, c), sp.lev.blahD, draw, &loss);
      *e4(sm, G::M2ROW, c) = m;
      *f32(sm, G::TD2 + mod(u - 1, 2) * G::P, c) = loss;
    }
  }

  // ------------------------------------------------------------------ interval 3 ----
  static RN_HD void interval3(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int u, int c)
  {
    if (active(cx, c)) {
      const int ts2 = G::TS + mod(u - 2, 3) * G::P, ts1 = G::TS + mod(u - 1, 3) * G::P;
      const int td1m = G::TD1 + mod(u - 1, 2) * G::P, td10 = G::TD1 + mod(u, 2) * G::P;
      const int td2mm = G::TD2 + mod(u - 2, 2) * G::P, td2m = G::TD2 + mod(u - 1, 2) * G::P;
      // corner K(u-1, c): generateSquare scatter of 4 squares, then loop 1 (same row), loop 2 (same column)
      E4 k = *e4(sm, G::KROW + mod(u - 1, 4) * G::P, c);
      float sh = k.w;
      sh -= *f32(sm, ts2, c - 1);
      sh -= *f32(sm, ts2, c);
      sh -= *f32(sm, ts1, c - 1);
      sh -= *f32(sm, ts1, c);
      sh -= *f32(sm, td1m, c - 1);
      sh -= *f32(sm, td1m, c);
      sh -= *f32(sm, td2mm, c);
      sh -= *f32(sm, td2m, c);
      k.w = sh;
      *e4(sm, G::OUTK, c) = k;
      // centre C(u-1, c): loop-1 mids above and below, loop-2 mids left and right
      E4 ce = *e4(sm, G::CROW + mod(u - 1, 2) * G::P, c);
      float shc = ce.w;
      shc -= *f32(sm, td1m, c);
      shc -= *f32(sm, td10, c);
      shc -= *f32(sm, td2m, c);
      shc -= *f32(sm, td2m, c + 1);
      ce.w = shc;
      *e4(sm, G::OUTC, c) = ce;
    }
    if (th.have_pre) {
      *e4(sm, G::KROW + mod(u + 2, 4) * G::P, c) = th.pre;
      if (c == 0) *e4(sm, G::KROW + mod(u + 2, 4) * G::P, cx.CW + 2) = th.pre_x;
    }
  }
};

}  // namespace rn
