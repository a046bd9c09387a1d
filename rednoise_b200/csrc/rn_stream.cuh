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
// Strips on a lattice column border, and the steps of any strip that touch the first / last lattice rows, run the EDGE
// instantiation (existence tests, the border forms of averageDiamond / generateDiamond, the closing row and column W); all
// other (strip, step) pairs run the lean one, in which the two midpoints of a cell are evaluated as packed FP32 pairs
// (rn_math.cuh: F2, displace2) and the thread's own samples stay in registers across the barrier intervals (RN_STREAM_CARRY).
//
// Step u of a strip (cell row u; cell (u,c) has its corner at output (2u, 2c')):
//   interval 1   S(u): centre C(u,c) from lattice rows u, u+1        (averageSquare+generateSquare)
//                + store of the two output rows finished by G(u-1)
//   interval 2   D(u): mids M1(u,c) and M2(u-1,c)                     (averageDiamond+generateDiamond)
//                + global load of lattice row u+2 into registers
//   interval 3   G(u): final shade of corner K(u-1,c) and centre C(u-1,c) (ordered gather)
//                + the prefetched lattice row goes to shared memory
// Rolling rows: lattice x4, centres x2, M1 x2, M2, final K, final C (float4) and the loss planes
// TS x4, TD1 x2, TD2 x2 (float).  See rn_tile.cuh for the vocabulary and the shade-order argument.
//
// Compiles as plain C++ for tests/host_replay (per-thread registers live in StreamThread).
#pragma once
#include "rn_tile.cuh"

#ifndef RN_STREAM_CARRY
#define RN_STREAM_CARRY 6   // (6 measured best: bit 0 costs spills at 64 registers) bit 0: K(u+1,c) from the prefetch register, bit 1: K(u,c), C(u,c) interval 1 -> 2, bit 2: K(u-1,c), C(u-1,c) interval 2 -> 3
#endif

namespace rn {

struct StreamParams {
  const E4* gin;   // complete depth-(dout-1) grid: lattice point (r, k) <-> output sample (2r, 2k)
  size_t gin_pitch, gin_field_stride;
  E4* gout;        // element (win_i0, win_j0) of the depth-dout grid
  size_t gout_pitch, gout_field_stride;
  int dout;
  int win_i0, win_j0, win_ni, win_nj;
  int a_lo, a_hi;  // cell rows [a_lo, a_hi) of the streamed region (a_hi <= n); close_i: output row W as well
  int b_lo, b_hi;  // cell columns; close_j: output column W as well
  int close_i, close_j;
  // Decomposition: blockIdx.x enumerates (strip, segment) pairs, the strips on a lattice column border first.  Those CTAs
  // run the EDGE instantiation for every step, which is slower per step, so their strips are cut into more, shorter
  // segments: all CTAs of the (single-wave) launch then take about the same time.
  int strip_cols;     // cell columns per strip, <= NT-3
  int nstrips;
  int edge_first, edge_last;  // strip 0 / strip nstrips-1 touches a column border
  int seg_rows, nseg;            // interior strips: cell rows per segment, segments
  int seg_rows_edge, nseg_edge;  // border strips
  uint32_t seed;
  const uint32_t* seeds;
  Law law;
  StageLevel lev;
  Jump jn, jn1;    // LCG jumps over n and n+1 draws (one cell row of generateSquare / loop 1, of loop 2)
  int volcano;
#ifdef RN_TIMELINE
  int tl_slot;          // instrumented builds: launch number (profiles/micro/timeline.py)
#endif
  float volcano_gain;
};

// registers of one thread that live across barrier intervals
struct StreamThread {
  uint32_t sS, sD1, sD2;  // LCG states whose extraction is the draw of this column's next S / loop-1 / loop-2 call
  E4 pre;                 // prefetched lattice point of row u+2 in this thread's column (kept until it is K(u+1, c) of the next step)
  // lean instantiation: values of the thread's own column that the next interval would otherwise re-read from shared memory
  // (the kernel is bound by the shared-memory pipe)
  E4 own_k, own_c;        // K(u, c) and C(u, c) from interval 1 for interval 2
  E4 own_km, own_cm;      // K(u-1, c) and C(u-1, c) from interval 2 for interval 3
  const E4* pre_src;      // where the next prefetch reads (advances one lattice row per step)
  E4* out_dst;            // output row 2a, column out_j of the next cell row to be stored (advances two rows per store)
  float sg0, sg1;         // shade of corners K(u,c) / K(u-1,c) after generateSquare's scatter (carried from step to step)
  // copy-out: this thread stores the output columns out_j and out_j + NT (relative to 2*B0) of every finished row
  int out_j;
  int out_ok;             // bit p: column out_j + p*NT is owned by the strip and inside the stored window
  float out_fy2[2];       // squared column distance from the volcano
};

struct StreamCtx {
  int B0, CW;   // first cell column, cell columns of this strip: cells [B0, B0+CW)
  int CWe;      // CW + 1 if the strip also owns the closing output column W
  int a0, a1;   // cell rows [a0, a1)
  int a1e;      // a1 + 1 if the segment also owns the closing output row W
  uint32_t seed;
  const E4* gin;
  E4* gout;
  int far;      // every sample of this CTA lies outside the volcano's crater (d >= minRad)
  // Steps [lean_lo, lean_hi] touch no lattice row border (every row they read or write exists), so warps whose columns
  // touch no column border run them in the lean instantiation; all other (warp, step) pairs run the EDGE one.
  int lean_lo, lean_hi;
  int col_lean;  // no thread column of the CTA touches a column border
  int ubase;     // ring rows are indexed by u - ubase (the first lean step has ring phase 0)
};

// crater test for a CTA: index distance of the nearest sample of its box from the volcano
RN_HD int stream_box_is_far(int i_lo, int i_hi, int j_lo, int j_hi, int W)
{
  const int ci = W / 2;
  const int di = imax_(imax_(i_lo - ci, ci - i_hi), 0), dj = imax_(imax_(j_lo - ci, ci - j_hi), 0);
  const float dmin = sqrtf((float)di * (float)di + (float)dj * (float)dj) / (float)W;
  return dmin > 0.0205f;
}

RN_HD int stream_grid_size(const StreamParams& sp)
{
  const int ne = sp.edge_first + sp.edge_last;
  return ne * sp.nseg_edge + (sp.nstrips - ne) * sp.nseg;
}

// id: blockIdx.x in [0, stream_grid_size)
RN_HD StreamCtx stream_ctx(const StreamParams& sp, int id, unsigned f)
{
  StreamCtx cx;
  const int ne = sp.edge_first + sp.edge_last;
  int strip, seg, rows;
  if (id < ne * sp.nseg_edge) {  // border strips: consecutive ids alternate between the two
    seg = ne == 2 ? id >> 1 : id;
    strip = (ne == 2 ? (id & 1) : sp.edge_last) ? sp.nstrips - 1 : 0;
    if (ne == 1 && sp.edge_first) strip = 0;
    rows = sp.seg_rows_edge;
  } else {
    const int k = id - ne * sp.nseg_edge, ni = sp.nstrips - ne;
    seg = k / ni;
    strip = sp.edge_first + (k - seg * ni);
    rows = sp.seg_rows;
  }
  cx.B0 = sp.b_lo + strip * sp.strip_cols;
  cx.CW = imin_(sp.strip_cols, sp.b_hi - cx.B0);
  cx.CWe = cx.CW + ((sp.close_j && cx.B0 + cx.CW == sp.b_hi) ? 1 : 0);
  cx.a0 = sp.a_lo + seg * rows;
  cx.a1 = imin_(cx.a0 + rows, sp.a_hi);
  cx.a1e = cx.a1 + ((sp.close_i && cx.a1 == sp.a_hi) ? 1 : 0);
  cx.seed = sp.seeds ? sp.seeds[f] : sp.seed;
  cx.gin = sp.gin + (size_t)f * sp.gin_field_stride;
  cx.gout = sp.gout + (size_t)f * sp.gout_field_stride;
  cx.far = stream_box_is_far(2 * cx.a0, 2 * cx.a1e - 1, 2 * cx.B0, 2 * (cx.B0 + cx.CWe) - 1, 1 << sp.dout);
  const int n = sp.lev.n;
  cx.col_lean = !(cx.B0 < 1 || cx.B0 + cx.CW > n - 1 || cx.CWe != cx.CW);
  // step u reads lattice rows u-1 .. u+2 (the last step of a CTA does not prefetch) and cell rows u-2 .. u
  cx.lean_lo = imax_(cx.a0 - 1, 1);
  cx.lean_hi = cx.a1e <= n - 1 ? cx.a1e : n - 2;
  cx.ubase = cx.lean_lo <= cx.lean_hi ? cx.lean_lo : cx.a0 - 1;
  return cx;
}
// A warp all of whose thread columns have existing cells, neighbours and lattice points can run the lean instantiation
// (c: any thread of the warp).
RN_HD bool stream_warp_is_lean(const StreamParams& sp, const StreamCtx& cx, int c)
{
  // (a per-warp rule -- lean warps beside the border warp of an edge strip -- was measured slower: 107 vs 97 us per CTA)
  (void)sp; (void)c;
  return cx.col_lean;
}
RN_HD bool stream_step_is_lean(const StreamCtx& cx, int u) { return u >= cx.lean_lo && u <= cx.lean_hi; }

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
  static constexpr int TS = 4 * E4ROWS;          // 4 rows (3 live; a power of two keeps the ring index a mask)
  static constexpr int TD1 = TS + 4 * P;         // 2 rows
  static constexpr int TD2 = TD1 + 2 * P;        // 2 rows
  static constexpr int WORDS = ((TD2 + 2 * P + 3) / 4) * 4;
  static constexpr size_t BYTES = (size_t)WORDS * 4;
};

template <int NT, bool EDGE>
struct Stream {
  using G = StreamGeo<NT>;
  static RN_HD E4* e4(float* sm, int row_off, int c) { return reinterpret_cast<E4*>(sm) + row_off + (c + 1); }
  static RN_HD float* f32(float* sm, int row_off, int c) { return sm + row_off + (c + 1); }
  // Ring rows are indexed by the step number t relative to the CTA's first step (ring sizes are powers of two).  The
  // unrolled march passes PH = t & 3 as a template argument, which turns every ring offset into an immediate;
  // PH < 0 computes it from t at run time.
  static RN_HD int mod(int t, int m) { return t & (m - 1); }
  // PH = 0..3: t & 3 is known; PH = 4, 5: only the parity t & 1 = PH - 4 is known (two steps per trip: the rings of
  // two are immediates, the rings of four are computed); PH < 0: nothing is known.
  template <int PH>
  static RN_HD int ring(int t, int k, int m)
  {
    if (PH >= 4) return m == 2 ? ((PH + k) & 1) : ((t + k) & (m - 1));
    return (PH >= 0 ? PH + k : t + k) & (m - 1);
  }
  static RN_HD int step_of(const StreamCtx& cx, int u) { return u - cx.ubase; }

  // threads with a cell; the lean instantiation lets the few surplus lanes compute on garbage (all in-bounds, never stored)
  static RN_HD bool active(const StreamCtx& cx, int c) { return !EDGE || c <= cx.CW + 1; }
  static RN_HD bool loader(const StreamCtx& cx, int c) { return c <= cx.CW + 2; }   // threads with a lattice column

  // lattice point (row r, thread column c) from the coarser grid; EDGE: only if it exists
  static RN_HD bool lattice_exists(const StreamParams& sp, const StreamCtx& cx, int r, int c)
  {
    const int k = cx.B0 - 1 + c, n = sp.lev.n;
    return !EDGE || (r >= 0 && r <= n && k >= 0 && k <= n);
  }
  static RN_HD E4 lattice(const StreamParams& sp, const StreamCtx& cx, int r, int c)
  {
    return cx.gin[(size_t)r * sp.gin_pitch + (size_t)(cx.B0 - 1 + c)];
  }
  static RN_HD void load_row_now(const StreamParams& sp, const StreamCtx& cx, float* sm, int r, int c)
  {
    if (lattice_exists(sp, cx, r, c)) *e4(sm, G::KROW + mod(step_of(cx, r), 4) * G::P, c) = lattice(sp, cx, r, c);
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
    th.sg0 = th.sg1 = 0.0f;
    {
      // Lane -> column permutation inside every 32-column group: each quarter warp reads 8 consecutive
      // slots of one class row (conflict-free LDS.128), the warp stores 512 contiguous bytes.
      const int Wout = 1 << sp.dout;
      const int lane = c & 31, q = lane >> 3, k = lane & 7;
      th.out_j = (c - lane) + (q & 1) * 16 + 2 * k + (q >> 1);
      th.out_ok = 0;
      for (int pass = 0; pass < 2; pass++) {
        const int j = th.out_j + pass * NT;  // column relative to 2*B0
        const int gj = 2 * cx.B0 + j;
        const float fy = (float)(gj - Wout / 2);
        const bool owned = j < 2 * cx.CW || (j == 2 * cx.CW && cx.CWe != cx.CW);
        if (owned && (unsigned)(gj - sp.win_j0) < (unsigned)sp.win_nj) th.out_ok |= 1 << pass;
        th.out_fy2[pass] = fy * fy;
      }
    }
    // (both pointers may start outside their grids; they are only dereferenced where the row / column exists)
    th.pre_src = cx.gin + ((ptrdiff_t)(u0 + 2) * (ptrdiff_t)sp.gin_pitch + (ptrdiff_t)(cx.B0 - 1 + c));
    th.out_dst = cx.gout + ((ptrdiff_t)(2 * cx.a0 - sp.win_i0) * (ptrdiff_t)sp.gout_pitch + (ptrdiff_t)(2 * cx.B0 - sp.win_j0 + th.out_j));
    if (loader(cx, c)) {
      load_row_now(sp, cx, sm, u0, c);
      load_row_now(sp, cx, sm, u0 + 1, c);
      if (lattice_exists(sp, cx, u0 + 1, c)) th.pre = lattice(sp, cx, u0 + 1, c);  // K(u0+1, c): interval 1 of the first step reads it from here
    }
  }

  // ------------------------------------------------------------------ interval 1 ----
  template <bool VOLCANO, bool FAR>
  static RN_HD void store_rows_t(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int a, int m1slot)
  {
    // output rows 2a (corners K(a,.) / mids M1(a,.)) and 2a+1 (mids M2(a,.) / centres C(a,.))
    const int Wout = 1 << sp.dout;
    const float inv_w = 1.0f / (float)Wout;
    const float gain_w = sp.volcano_gain * (float)Wout;
    const E4* smE = reinterpret_cast<const E4*>(sm);
    const bool odd = th.out_j & 1;
    const int cc0 = (th.out_j >> 1) + 2, cc1 = cc0 + NT / 2;  // slot of the owning cell inside a class row (its c + 1)
    const bool ok0 = th.out_ok & 1, ok1 = th.out_ok & 2;
    E4* dst = th.out_dst;
    th.out_dst = dst + 2 * sp.gout_pitch;
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int gi = 2 * a + r;
      const int wi = gi - sp.win_i0;
      if ((unsigned)wi < (unsigned)sp.win_ni && gi <= Wout) {  // uniform
        E4* row = r == 0 ? dst : dst + sp.gout_pitch;
        const float fx = (float)(gi - Wout / 2);
        const float fx2 = fx * fx;
        const int src = r == 0 ? (odd ? G::M1ROW + m1slot * G::P : G::OUTK) : (odd ? G::OUTC : G::M2ROW);
        float h0 = 0.0f, h1 = 0.0f;
        if (VOLCANO && FAR) {  // both columns of the thread at once (packed pair)
          const F2 h = f2mul(rcp_rn_normal2(sqrt_rn_normal2(f2add(f2s(fx2), f2(th.out_fy2[0], th.out_fy2[1])))), f2s(gain_w));
          h0 = f2lo(h);
          h1 = f2hi(h);
        } else if (VOLCANO) {
          h0 = volcano_height_r2<false>(fx2 + th.out_fy2[0], inv_w, sp.volcano_gain, gain_w);
          h1 = volcano_height_r2<false>(fx2 + th.out_fy2[1], inv_w, sp.volcano_gain, gain_w);
        }
        if (ok0) {
          E4 e = smE[src + cc0];
          if (VOLCANO) e.y = e.y + h0;
          row[0] = e;
        }
        if (ok1) {
          E4 e = smE[src + cc1];
          if (VOLCANO) e.y = e.y + h1;
          row[NT] = e;
        }
      }
    }
  }
  // m1slot: ring slot of the M1 row of cell row a
  static RN_HD void store_rows(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int a, int m1slot)
  {
    if (!sp.volcano) store_rows_t<false, true>(sp, cx, sm, th, a, m1slot);
    else if (cx.far) store_rows_t<true, true>(sp, cx, sm, th, a, m1slot);
    else store_rows_t<true, false>(sp, cx, sm, th, a, m1slot);
  }

  template <int PH = -1>
  static RN_HD void interval1(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int u, int c)
  {
    const int t = step_of(cx, u);
    if (active(cx, c)) {
      const int n = sp.lev.n;
      const int bg = cx.B0 - 1 + c;
      const bool ex = !EDGE || (u >= 0 && u < n && bg >= 0 && bg < n);
      float loss = 0.0f;
      const uint32_t draw = lcg_draw_from_state(th.sS);
      th.sS = jump_apply(sp.jn, th.sS);
      if (ex) {
        // S(u): averageSquare (Terrain.cpp:216-222) + generateSquare (:224-258)
        const int r0 = G::KROW + ring<PH>(t, 0, 4) * G::P, r1 = G::KROW + ring<PH>(t, 1, 4) * G::P;
        const E4 k00 = *e4(sm, r0, c), k01 = *e4(sm, r0, c + 1), k10 = (EDGE || !(RN_STREAM_CARRY & 1)) ? *e4(sm, r1, c) : th.pre, k11 = *e4(sm, r1, c + 1);
        const E4 a4 = avg4_packed(k00, k10, k01, k11);
        const V3 avg = pos4(a4);
        const float sh = a4.w;
        const Displaced dp = displace<true>(sp.law, avg, vsub(pos4(k11), pos4(k00)), vsub(pos4(k10), pos4(k01)), sp.lev.blahS, draw);
        const E4 cnew = e4_of(dp.pos, sh + shade_gain(sp.law, dp.offset));
        *e4(sm, G::CROW + ring<PH>(t, 0, 2) * G::P, c) = cnew;
        if (!EDGE && (RN_STREAM_CARRY & 2)) { th.own_k = k00; th.own_c = cnew; }
        loss = shade_loss(sp.law, dp.offset);
      }
      *f32(sm, G::TS + ring<PH>(t, 0, 4) * G::P, c) = loss;
    }
    // the rows G(u-1) finished: cell row u-2
    if (u - 2 >= cx.a0 && u - 2 < cx.a1e) store_rows(sp, cx, sm, th, u - 2, ring<PH>(t, -2, 2));
  }

  // ------------------------------------------------------------------ interval 2 ----
  template <int PH = -1>
  static RN_HD void interval2(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int u, int c)
  {
    const int t = step_of(cx, u);
    // prefetch lattice row u+2 (needed by S(u+1) and D(u+2)); it lands in shared memory in interval 3
    if (loader(cx, c) && u + 1 <= cx.a1e && lattice_exists(sp, cx, u + 2, c)) th.pre = *th.pre_src;
    th.pre_src += sp.gin_pitch;
    if (!active(cx, c)) return;
    const int n = sp.lev.n;
    const int bg = cx.B0 - 1 + c;
    bool v1 = true, v2 = true;  // M1(u, c), M2(u-1, c) exist
    if (EDGE) {
      v1 = u >= 0 && u <= n && bg >= 0 && bg < n;
      v2 = u - 1 >= 0 && u - 1 < n && bg >= 0 && bg <= n;
    }
    const uint32_t draw1 = lcg_draw_from_state(th.sD1), draw2 = lcg_draw_from_state(th.sD2);
    th.sD1 = jump_apply(sp.jn, th.sD1);
    th.sD2 = jump_apply(sp.jn1, th.sD2);
    float loss1 = 0.0f, loss2 = 0.0f;
    // corner shades as generateSquare left them (Terrain.cpp:252-255): minus the scatter of the 4 surrounding
    // centres.  sg(u,c) is computed here once and carried to the next step (as sg(u-1,c)) and to interval 3.
    const int ts1 = G::TS + ring<PH>(t, -1, 4) * G::P, ts0 = G::TS + ring<PH>(t, 0, 4) * G::P;
    const int kr = G::KROW + ring<PH>(t, 0, 4) * G::P;
    const float t10 = *f32(sm, ts1, c), t00 = *f32(sm, ts0, c);
    E4 k00 = (EDGE || !(RN_STREAM_CARRY & 2)) ? *e4(sm, kr, c) : th.own_k;  // K(u, c)
    k00.w = (((k00.w - *f32(sm, ts1, c - 1)) - t10) - *f32(sm, ts0, c - 1)) - t00;
    th.sg1 = th.sg0;   // sg(u-1, c)
    th.sg0 = k00.w;    // sg(u, c)
    if (!EDGE) {
      // both midpoints exist and have all four neighbours: one packed evaluation (rn_tile.cuh mid_interior2)
      const int krm = G::KROW + ring<PH>(t, -1, 4) * G::P;
      const int crm = G::CROW + ring<PH>(t, -1, 2) * G::P;
      E4 k01 = *e4(sm, kr, c + 1);
      k01.w = (((k01.w - t10) - *f32(sm, ts1, c + 1)) - t00) - *f32(sm, ts0, c + 1);
      E4 km = *e4(sm, krm, c);  // K(u-1, c)
      km.w = th.sg1;
      const E4 cm = *e4(sm, crm, c);  // C(u-1, c): x neighbour of M1, y neighbour of M2
      E4 m1, m2;
      if (RN_STREAM_CARRY & 4) { th.own_km = km; th.own_cm = cm; }
      mid_interior2(sp.law, cm, (RN_STREAM_CARRY & 2) ? th.own_c : *e4(sm, G::CROW + ring<PH>(t, 0, 2) * G::P, c), k00, k01, km, k00, *e4(sm, crm, c - 1), cm, sp.lev.blahD, draw1, draw2, &m1, &m2, &loss1, &loss2);
      *e4(sm, G::M1ROW + ring<PH>(t, 0, 2) * G::P, c) = m1;
      *e4(sm, G::M2ROW, c) = m2;
    } else if (v1 || v2) {
      const int krm = G::KROW + ring<PH>(t, -1, 4) * G::P;
      const int cr = G::CROW + ring<PH>(t, 0, 2) * G::P, crm = G::CROW + ring<PH>(t, -1, 2) * G::P;
      if (v1) {  // loop 1: M1(u, c) at output (2u, 2c'+1): x neighbours C(u-1,c), C(u,c); y neighbours K(u,c), K(u,c+1)
        E4 k01 = *e4(sm, kr, c + 1);
        k01.w = (((k01.w - t10) - *f32(sm, ts1, c + 1)) - t00) - *f32(sm, ts0, c + 1);
        E4 m;
        if (EDGE && (u == 0 || u == n)) {
          const E4 dummy = k00;
          m = mid_border(sp.law, u == 0 ? dummy : *e4(sm, crm, c), u == n ? dummy : *e4(sm, cr, c), k00, k01, u == 0, u == n, false, false,
                         sp.lev.blahDh, draw1, &loss1);
        } else {
          m = mid_interior(sp.law, *e4(sm, crm, c), *e4(sm, cr, c), k00, k01, sp.lev.blahD, draw1, &loss1);
        }
        *e4(sm, G::M1ROW + ring<PH>(t, 0, 2) * G::P, c) = m;
      }
      if (v2) {  // loop 2: M2(u-1, c) at output (2u-1, 2c'): x neighbours K(u-1,c), K(u,c); y neighbours C(u-1,c-1), C(u-1,c)
        E4 km = *e4(sm, krm, c);  // K(u-1, c)
        km.w = th.sg1;
        E4 m;
        if (EDGE) {  // column borders by selects: no divergence inside the border warp
          m = mid_select_y(sp.law, km, k00, *e4(sm, crm, c - 1), *e4(sm, crm, c), bg == 0, bg == n, sp.lev.blahD, draw2, &loss2);
        } else {
          m = mid_interior(sp.law, km, k00, *e4(sm, crm, c - 1), *e4(sm, crm, c), sp.lev.blahD, draw2, &loss2);
        }
        *e4(sm, G::M2ROW, c) = m;
      }
    }
    *f32(sm, G::TD1 + ring<PH>(t, 0, 2) * G::P, c) = loss1;
    *f32(sm, G::TD2 + ring<PH>(t, -1, 2) * G::P, c) = loss2;
  }

  // ------------------------------------------------------------------ interval 3 ----
  template <int PH = -1>
  static RN_HD void interval3(const StreamParams& sp, const StreamCtx& cx, float* sm, StreamThread& th, int u, int c)
  {
    const int t = step_of(cx, u);
    if (active(cx, c)) {
      const int n = sp.lev.n;
      const int bg = cx.B0 - 1 + c;
      const int td1m = G::TD1 + ring<PH>(t, -1, 2) * G::P, td10 = G::TD1 + ring<PH>(t, 0, 2) * G::P;
      const int td2mm = G::TD2 + ring<PH>(t, -2, 2) * G::P, td2m = G::TD2 + ring<PH>(t, -1, 2) * G::P;
      if (!EDGE || (u - 1 >= 0 && u - 1 <= n && bg >= 0 && bg <= n)) {
        // corner K(u-1, c): generateSquare's scatter (carried sg), then loop 1 (same row), loop 2 (same column)
        E4 k = (EDGE || !(RN_STREAM_CARRY & 4)) ? *e4(sm, G::KROW + ring<PH>(t, -1, 4) * G::P, c) : th.own_km;
        float sh = th.sg1;
        sh -= *f32(sm, td1m, c - 1);
        sh -= *f32(sm, td1m, c);
        sh -= *f32(sm, td2mm, c);
        sh -= *f32(sm, td2m, c);
        k.w = sh;
        *e4(sm, G::OUTK, c) = k;
      }
      if (!EDGE || (u - 1 >= 0 && u - 1 < n && bg >= 0 && bg < n)) {
        // centre C(u-1, c): loop-1 mids above and below, loop-2 mids left and right
        E4 ce = (EDGE || !(RN_STREAM_CARRY & 4)) ? *e4(sm, G::CROW + ring<PH>(t, -1, 2) * G::P, c) : th.own_cm;
        float shc = ce.w;
        shc -= *f32(sm, td1m, c);
        shc -= *f32(sm, td10, c);
        shc -= *f32(sm, td2m, c);
        shc -= *f32(sm, td2m, c + 1);
        ce.w = shc;
        *e4(sm, G::OUTC, c) = ce;
      }
    }
    // the row prefetched in interval 2
    if (loader(cx, c) && u + 1 <= cx.a1e && lattice_exists(sp, cx, u + 2, c)) *e4(sm, G::KROW + ring<PH>(t, 2, 4) * G::P, c) = th.pre;
  }
};

}  // namespace rn
