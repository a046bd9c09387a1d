// rn_march.cuh -- algorithm 2, warp-synchronous streaming form of ONE level of the diamond-square path
// (Terrain.cpp:60-82 for a single `step`): the output kernel of deep fields since round 2.
//
// Same march as rn_stream.cuh (a strip of cell columns walked row by row, nothing recomputed in the marching
// direction, ordered shade gather), re-cut for the instruction-issue bound that kernel ran into:
//   * ONE WARP owns a strip and nothing is shared between warps: no block barriers.  What a lane needs from its
//     neighbours (one centre, two generateSquare losses, two generateDiamond losses per step) travels by warp shuffles.
//   * a lane owns TWO adjacent cell columns, so its two square centres, its two loop-1 midpoints and its two loop-2
//     midpoints are each evaluated as ONE packed-FP32 instruction stream (rn_math.cuh: F2, displace2), the values the two
//     cells share are loaded once, and the per-step bookkeeping is paid once per eight output samples;
//   * a lane's four output columns of a row are contiguous: finished rows go from registers to global memory as full
//     32-byte sectors (st.global.v8.f32 -> STG.E.ENL2.256).  Rows whose first sample sits on an odd 16-byte slot (every
//     other row of a (W+1)-pitched field) are shifted by one sample with four shuffles; the two samples at the warp's
//     ends are stored as 16-byte halves.
//   * registers hold a value only WITHIN a step.  What step u hands to step u+1 (centres, loop-1 mids, the loss terms) is
//     parked in lane-private shared-memory slots (two banks, element e of lane l at [e][l]: every access is a
//     conflict-free 128-bit one), and the input lattice arrives through a warp-private ring of four rows filled by
//     cp.async one step ahead, so the L2 latency of the input never reaches the issue stream.  That keeps the kernel at
//     128 registers -- 16 resident warps per SM, which is what the issue rate needs -- without spills.
//
// Lanes: lane l holds the lattice / cell columns k0 = L0 + 2l and k1 = k0 + 1 (L0 even) and reads column k1 + 1, its right
// neighbour's k0, as well.  Lanes 0 and 31 are halo: their samples are incomplete (they have no left / right neighbour) and
// are never stored, so a warp finishes the 60 cell columns [L0 + 2, L0 + 62) = 120 output columns per row.
//
// Step u (cell row u; vocabulary of rn_tile.cuh):
//   P1   S(u): centres C(u, k0), C(u, k1) from lattice rows u, u+1        (averageSquare + generateSquare);
//        lattice row u+2 is requested (cp.async) into the ring slot row u-2 occupied
//   X1   losses TS(u, .) to both neighbours, C(u, k1) to the right neighbour
//   P2   D(u): mids M1(u, k0), M1(u, k1) and M2(u-1, k0), M2(u-1, k1)       (averageDiamond + generateDiamond)
//   X2   losses TD1(u, k1) to the right, TD2(u-1, k0) to the left neighbour
//   P3   G(u-1): final shade of corners K(u-1, .) and centres C(u-1, .) (ordered gather), addVolcano, store of output
//        rows 2(u-1) and 2(u-1)+1
// Steps that touch the first / last lattice rows, strips on a lattice column border and strips cut by the window run the
// EDGE instantiation (existence tests, border forms, per-sample stores).
//
// Compiles as plain C++ for tests/host_replay (a warp = 32 MarchLane structs + one block of "shared memory", a shuffle = a
// read of the neighbour's struct).
#pragma once
#include "rn_stream.cuh"

namespace rn {

constexpr int kMarchCols = 60;  // cell columns a warp finishes
#ifndef RN_EXP
#define RN_EXP 0           // timing experiments only (results are wrong): bit 0 no global stores, bit 1 no addVolcano
#endif
#ifndef RN_MARCH_CARRY
#define RN_MARCH_CARRY 0   // values kept in registers from phase to phase instead of being re-read from shared memory (lean
                           // instantiation): bit 0 K(u-1, k0/k1) P2 -> P3, bit 1 C(u-1, k0/k1) P2 -> P3, bit 2 K(u, .) P1 -> P2
#endif

// Warp-private shared memory, in E4 (16-byte) units.
//   ring    four input lattice rows (row r in slot r & 3), even and odd columns apart so that every access of a warp is one
//           contiguous run: lane l's column k0 at [l] -- so [l+1] is its right neighbour's k0, [32] lane 31's third
//           column -- and its column k0+1 at [kRingOdd + l]
//   state   two banks (step parity) of the values a step hands to the next: element e of lane l at [e][l]
//   const   per-lane constants: the squared column distances of the lane's four output columns from the volcano
constexpr int kRingOdd = 36;     // 33 even-column entries; 36 = 4 (mod 8): the even and the odd columns a quarter-warp's coalesced
                                 // cp.async writes (4 + 4 elements) fall on disjoint banks
constexpr int kRingRowE4 = 68;
constexpr int kRingE4 = 4 * kRingRowE4;
enum { S_C0 = 0, S_C1, S_C2, S_M1A, S_M1B, S_TS, S_TD, S_TE, S_COUNT };  // C(u; k0-1, k0, k1), M1(u; k0, k1), TS(u; 4),
                                                                          // {TD1(u; 3), sg(u, k0)}, {TD2(u-1; 3), sg(u, k1)}
constexpr int kBankE4 = S_COUNT * 32;
constexpr int kStateE4 = kRingE4;
constexpr int kConstE4 = kStateE4 + 2 * kBankE4;
constexpr int kWarpE4 = kConstE4 + 32;  // 808 E4 = 12.6 KB per warp
RN_HD E4* ring_row(E4* wsm, int r) { return wsm + (r & 3) * kRingRowE4; }
RN_HD const E4& ring_at(const E4* row_lane, int q) { return row_lane[q == 1 ? kRingOdd : (q >> 1)]; }  // row_lane = ring_row(wsm, r) + lane
RN_HD int ring_slot(int e) { return (e & 1) ? kRingOdd + (e >> 1) : (e >> 1); }  // ring entry of the warp's lattice column L0 + e, e = 0 .. 64
RN_HD E4* state_at(E4* wsm, int bank, int e, int lane) { return wsm + kStateE4 + bank * kBankE4 + e * 32 + lane; }

// registers of a lane: what lives across steps, and what one phase of a step hands to the next
struct MarchLane {
  uint32_t sS, sD1, sD2;  // LCG states whose extraction is the draw of column k0's next S / loop-1 / loop-2 call
  const E4* src;   // lattice (row of the next cp.async, column L0 + lane)
  E4* dst;         // output (row 2(u-1), column 2*k0) of the step in progress, relative to the window
  int ok;          // bit j: output column 2*k0 + j belongs to the strip and lies inside the stored window
  // within a step
  E4 cen[3];       // C(u; k0-1, k0, k1)
  E4 m1[2];        // M1(u; k0, k1)
  float ts[4];     // TS(u; k0-1, k0, k1, k1+1)
  float td1[3];    // TD1(u; k0-1, k0, k1)
  float td2[3];    // TD2(u-1; k0, k1, k1+1)
  float sg[2];     // shade of corner K(u, k0/k1) as generateSquare's scatter left it
  E4 tdm, tem;     // what step u-1 left in its S_TD / S_TE slots (read in P2, used again in P3)
  E4 c_km[2], c_cm[2], c_ka[3];  // RN_MARCH_CARRY
  E4 out[2][4];    // the two finished output rows of the lane
};

struct MarchCtx {
  int L0;          // lattice column of lane 0's first column (even)
  int B0, CW, CWe; // owned cell columns [B0, B0+CW), B0 = L0 + 2; CWe = CW + 1 with the closing output column W
  int a0, a1, a1e; // owned cell rows [a0, a1); a1e = a1 + 1 with the closing output row W
  uint32_t seed;
  const E4* gin;
  E4* gout;
  int far;         // every sample of the warp lies outside the volcano's crater
  int lean_lo, lean_hi;  // steps that touch no lattice row border
  int col_lean;    // all 65 lattice columns of the warp exist and it owns 60 whole cell columns inside the window (sector stores):
                   // its interior steps run the lean instantiation
  int ubase;       // bank of step u = (u - ubase) & 1
};

RN_HD void ring_copy16(E4* dst_smem, const E4* src_global)
{
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src_global) : "memory");
#else
  *dst_smem = *src_global;
#endif
}
RN_HD void ring_commit()
{
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
// everything but the most recent group of EVERY lane has landed and is visible to the whole warp
RN_HD void ring_wait_all_but_1()
{
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncwarp();
#endif
}

// id enumerates (strip, segment) pairs exactly like stream_ctx: the strips on a lattice column border first, cut into
// shorter segments because every one of their steps runs the slower EDGE instantiation.
RN_HD MarchCtx march_ctx(const StreamParams& sp, int id, unsigned f)
{
  MarchCtx cx;
  const int ne = sp.edge_first + sp.edge_last;
  int strip, seg, rows;
  if (id < ne * sp.nseg_edge) {
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
  const int n = sp.lev.n;
  cx.B0 = sp.b_lo + strip * sp.strip_cols;
  cx.L0 = cx.B0 - 2;
  cx.CW = imin_(sp.strip_cols, sp.b_hi - cx.B0);
  cx.CWe = cx.CW + ((sp.close_j && cx.B0 + cx.CW == sp.b_hi) ? 1 : 0);
  cx.a0 = sp.a_lo + seg * rows;
  cx.a1 = imin_(cx.a0 + rows, sp.a_hi);
  cx.a1e = cx.a1 + ((sp.close_i && cx.a1 == sp.a_hi) ? 1 : 0);
  cx.seed = sp.seeds ? sp.seeds[f] : sp.seed;
  cx.gin = sp.gin + (size_t)f * sp.gin_field_stride;
  cx.gout = sp.gout + (size_t)f * sp.gout_field_stride;
  cx.far = stream_box_is_far(2 * cx.a0, 2 * cx.a1e - 1, 2 * cx.B0, 2 * (cx.B0 + cx.CWe) - 1, 1 << sp.dout);
  const bool fast = cx.CW == kMarchCols && cx.CWe == cx.CW && 2 * cx.B0 >= sp.win_j0 && 2 * (cx.B0 + cx.CW) <= sp.win_j0 + sp.win_nj;
  cx.col_lean = cx.L0 >= 0 && cx.L0 + 64 <= n && fast;
  cx.lean_lo = imax_(cx.a0 - 1, 1);
  cx.lean_hi = cx.a1e <= n - 1 ? cx.a1e : n - 2;
  // a lean step stores both output rows 2a, 2a + 1 of its cell row a = u - 1 without looking at the window
  const int a_in_lo = (sp.win_i0 + 1) >> 1, a_in_hi = (sp.win_i0 + sp.win_ni - 2) >> 1;
  if (a_in_lo > cx.a0) cx.lean_lo = imax_(cx.lean_lo, a_in_lo + 1);
  cx.lean_hi = imin_(cx.lean_hi, a_in_hi + 1);
  cx.ubase = cx.a0 - 1;
  return cx;
}
// The two warm-up steps of a unit (u = a0 - 1, a0) exist for what its first OWNED cell row a0 takes from the row above: the centres
// C(a0 - 1) (P1 of step a0 - 1: their generateSquare losses reach the corners K(a0), their positions the midpoints M2(a0 - 1) and
// M1(a0)) and the loss terms of M2(a0 - 1) (P2 of step a0).  Everything else those steps would compute -- M1(a0 - 1) and M2(a0 - 2)
// in P2 of step a0 - 1, the final shades, addVolcano and store of the rows a0 - 2 and a0 - 1 in P3 -- only ever reaches samples
// the unit does not own, so it is skipped (the LCG states and the output pointer still advance).
// Measured (per-unit %globaltimer trace, profiles/micro/cta_trace.py, four A/B pairs): with the skip the units of the output kernel take
// 71.6 us (median) against 69.5 us without it and the kernel 78.5-79.7 against 77.8-78.3 us -- less work, no gain -- so the shipped
// build runs every phase in every step.
#ifndef RN_MARCH_WARMUP_SKIP
#define RN_MARCH_WARMUP_SKIP 0
#endif
RN_HD bool march_runs_p2(const MarchCtx& cx, int u) { return !RN_MARCH_WARMUP_SKIP || u >= cx.a0; }
RN_HD bool march_runs_p3(const MarchCtx& cx, int u) { return !RN_MARCH_WARMUP_SKIP || u > cx.a0; }
RN_HD bool march_step_is_lean(const MarchCtx& cx, int u) { return cx.col_lean && u >= cx.lean_lo && u <= cx.lean_hi; }

// ---------------------------------------------------------------- lane <-> lane and lane <-> memory ----
// On the device a value from the neighbouring lane is a shuffle; in the host replay the caller passes the neighbour's struct
// (nullptr at the ends of the warp, where the device shuffle returns the lane's own value: such values only ever reach
// the halo lanes' own, never stored, samples -- the replay feeds NaN there to prove it).
RN_HD float march_nan()
{
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x7fc0beef);
#else
  union { uint32_t u; float f; } v;
  v.u = 0x7fc0beefu;
  return v.f;
#endif
}
RN_HD float from_left(float mine, const float* theirs)   // the value `mine` names, as held by lane - 1
{
#if defined(__CUDA_ARCH__)
  (void)theirs;
  return __shfl_up_sync(0xffffffffu, mine, 1);
#else
  (void)mine;
  return theirs ? *theirs : march_nan();
#endif
}
RN_HD float from_right(float mine, const float* theirs)  // ... by lane + 1
{
#if defined(__CUDA_ARCH__)
  (void)theirs;
  return __shfl_down_sync(0xffffffffu, mine, 1);
#else
  (void)mine;
  return theirs ? *theirs : march_nan();
#endif
}
RN_HD E4 from_left4(const E4& mine, const E4* theirs)
{
  E4 r;
  r.x = from_left(mine.x, theirs ? &theirs->x : nullptr);
  r.y = from_left(mine.y, theirs ? &theirs->y : nullptr);
  r.z = from_left(mine.z, theirs ? &theirs->z : nullptr);
  r.w = from_left(mine.w, theirs ? &theirs->w : nullptr);
  return r;
}

RN_HD void st128(E4* p, const E4& a) { *p = a; }
RN_HD void st256(E4* p, const E4& a, const E4& b)  // p is 32-byte aligned: one full sector
{
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y),
               "f"(b.z), "f"(b.w)
               : "memory");
#else
  p[0] = a;
  p[1] = b;
#endif
}
// A vote of the warp (device) / the lane's own value (host replay: the two tails it chooses between are bit-identical wherever both
// apply, so a per-lane choice replays the same field and exercises both).
RN_HD bool warp_all(bool pred)
{
#if defined(__CUDA_ARCH__)
  return __all_sync(0xffffffffu, pred) != 0;
#else
  return pred;
#endif
}

// Two samples displaced at once from their averaged elements and direction vectors (the tail of generateSquare /
// generateDiamond, Terrain.cpp:185-213,230-257), in two halves so that a step evaluates the independent halves of several pairs
// side by side and decides ONCE, for the whole warp, between the general tail (unit normal by square root and reciprocal) and the
// flat one: at the last levels of a deep field every |dir1 x dir2|^2 is below the 1e-10 threshold of Vector3::normalise and the
// normal is (0, 0, 1) -- only z moves (rn_math.cuh shift_flat2_z).  Bit for bit what mid_interior2 / displace2 do.
struct PairWork {
  F2 a_xy, a_zw, b_xy, b_zw;  // averaged elements of sample A and sample B
  V3P c;                      // dir1 x dir2
  F2 f;                       // |c|^2
  Offset2 o;
};
RN_HD PairWork pair_begin(const Law& law, F2 a_xy, F2 a_zw, F2 b_xy, F2 b_zw, const V3P& dir1, const V3P& dir2, float blah, uint32_t draw_a,
                          uint32_t draw_b)
{
  PairWork w;
  w.a_xy = a_xy; w.a_zw = a_zw; w.b_xy = b_xy; w.b_zw = b_zw;
  w.c = vcross2(dir1, dir2, law.one);
  w.f = vnorm2_sq(w.c, law.one);
  w.o = offset2(law, f2lo(a_xy), f2hi(a_xy), f2lo(b_xy), f2hi(b_xy), blah, draw_a, draw_b);
  return w;
}
template <bool FLAT>
RN_HD void pair_finish(const Law& law, const PairWork& w, float blah, E4* out_a, E4* out_b, float* loss_a, float* loss_b)
{
  const F2 rs = f2s(law.radiosity_strength);
  const F2 gain = f2mul(w.o.offset, rs);
  const F2 loss = f2mul(f2mul(f2s(0.25f), w.o.offset), rs);
  *loss_a = f2lo(loss);
  *loss_b = f2hi(loss);
  if (FLAT) {
    const F2 sz = shift_flat2_z(blah, w.o);
    *out_a = E4{f2lo(w.a_xy), f2hi(w.a_xy), f2lo(w.a_zw) + f2lo(sz), f2hi(w.a_zw) + f2lo(gain)};
    *out_b = E4{f2lo(w.b_xy), f2hi(w.b_xy), f2lo(w.b_zw) + f2hi(sz), f2hi(w.b_zw) + f2hi(gain)};
  } else {
    const V3P sh = shift_general2(w.c, w.f, blah, w.o);
    *out_a = E4{f2lo(w.a_xy) + f2lo(sh.x), f2hi(w.a_xy) + f2lo(sh.y), f2lo(w.a_zw) + f2lo(sh.z), f2hi(w.a_zw) + f2lo(gain)};
    *out_b = E4{f2lo(w.b_xy) + f2hi(sh.x), f2hi(w.b_xy) + f2hi(sh.y), f2lo(w.b_zw) + f2hi(sh.z), f2hi(w.b_zw) + f2hi(gain)};
  }
}

// First half of averageSquare + generateSquare (Terrain.cpp:216-258) for the two cells of a lane: cell A has corners a00 (u, k),
// a10 (u+1, k), a01 (u, k+1), a11 (u+1, k+1); cell B likewise.  Same operand order as the scalar form in rn_stream.cuh interval1.
RN_HD PairWork centre_pair_begin(const Law& law, const E4& a00, const E4& a10, const E4& a01, const E4& a11, const E4& b00, const E4& b10,
                                 const E4& b01, const E4& b11, float blah, uint32_t draw_a, uint32_t draw_b)
{
  const F2 q = f2s(0.25f);
  const F2 a_xy = f2mul(f2add(f2add(f2add(e4_xy(a00), e4_xy(a10)), e4_xy(a01)), e4_xy(a11)), q);
  const F2 a_zw = f2mul(f2add(f2add(f2add(e4_zw(a00), e4_zw(a10)), e4_zw(a01)), e4_zw(a11)), q);
  const F2 b_xy = f2mul(f2add(f2add(f2add(e4_xy(b00), e4_xy(b10)), e4_xy(b01)), e4_xy(b11)), q);
  const F2 b_zw = f2mul(f2add(f2add(f2add(e4_zw(b00), e4_zw(b10)), e4_zw(b01)), e4_zw(b11)), q);
  const V3P dir1 = v3p(vsub(pos4(a11), pos4(a00)), vsub(pos4(b11), pos4(b00)));
  const V3P dir2 = v3p(vsub(pos4(a10), pos4(a01)), vsub(pos4(b10), pos4(b01)));
  return pair_begin(law, a_xy, a_zw, b_xy, b_zw, dir1, dir2, blah, draw_a, draw_b);
}
// First half of averageDiamond + generateDiamond for two interior midpoints (rn_tile.cuh mid_interior2: x neighbours ax0, ax1,
// y neighbours ay0, ay1 of sample A; likewise B)
RN_HD PairWork mid_pair_begin(const Law& law, const E4& ax0, const E4& ax1, const E4& ay0, const E4& ay1, const E4& bx0, const E4& bx1,
                              const E4& by0, const E4& by1, float blah, uint32_t draw_a, uint32_t draw_b)
{
  const F2 q = f2s(0.25f);
  const F2 a_xy = f2mul(f2add(f2add(f2add(e4_xy(ax0), e4_xy(ax1)), e4_xy(ay0)), e4_xy(ay1)), q);
  const F2 a_zw = f2mul(f2add(f2add(f2add(e4_zw(ax0), e4_zw(ax1)), e4_zw(ay0)), e4_zw(ay1)), q);
  const F2 b_xy = f2mul(f2add(f2add(f2add(e4_xy(bx0), e4_xy(bx1)), e4_xy(by0)), e4_xy(by1)), q);
  const F2 b_zw = f2mul(f2add(f2add(f2add(e4_zw(bx0), e4_zw(bx1)), e4_zw(by0)), e4_zw(by1)), q);
  const V3P dir1 = v3p(vsub(pos4(ax0), pos4(ax1)), vsub(pos4(bx0), pos4(bx1)));
  const V3P dir2 = v3p(vsub(pos4(ay1), pos4(ay0)), vsub(pos4(by1), pos4(by0)));
  return pair_begin(law, a_xy, a_zw, b_xy, b_zw, dir1, dir2, blah, draw_a, draw_b);
}

// addVolcano (Terrain.cpp:129-137) for two samples at once, crater branch included: volcano_height_r2<false> lane by lane
// (same operations in the same order; the sums after a packed product go through f2add_product, see rn_math.cuh).
RN_HD F2 volcano_pair_general(F2 r2, float inv_w, float gain, float gain_w, float one)
{
  const float minRad = 0.02f;
  const F2 s = sqrt_rn_normal2(r2);  // not a number for r2 == 0 (the centre sample): replaced below
  const F2 far = f2mul(rcp_rn_normal2(s), f2s(gain_w));
  F2 d = f2mul(s, f2s(inv_w));
  d = f2(f2lo(r2) == 0.0f ? 0.0f : f2lo(d), f2hi(r2) == 0.0f ? 0.0f : f2hi(d));
  const F2 t = f2add_product(f2mul(d, d), f2s(-(minRad * minRad)), one);
  const F2 crater = f2mul(f2add_product(f2mul(f2s(60000.f), t), f2s(1.0f / minRad), one), f2s(gain));
  return f2(!(f2lo(d) >= minRad) ? f2lo(crater) : f2lo(far), !(f2hi(d) >= minRad) ? f2hi(crater) : f2hi(far));
}

template <bool EDGE>
struct March {
  static RN_HD int bank_of(const MarchCtx& cx, int u) { return (u - cx.ubase) & 1; }

  // cp.async of lattice row r into the ring (from p = &gin[r][L0 + lane]): the warp's 65 lattice columns as two coalesced
  // runs of 32 elements (lane l copies columns L0 + l and L0 + 32 + l) and column L0 + 64 by lane 0; always exactly one
  // commit group
  static RN_HD void prefetch_row(const StreamParams& sp, const MarchCtx& cx, E4* wsm, const E4* p, int r, int lane)
  {
    if (r <= cx.a1e + 1 && r >= 0 && r <= sp.lev.n) {  // uniform: a later step reads the row, and it exists
      E4* row = ring_row(wsm, r);
      const int k = cx.L0 + lane, n = sp.lev.n;
      if (!EDGE || (k >= 0 && k <= n)) ring_copy16(row + ring_slot(lane), p);
      if (!EDGE || (k + 32 >= 0 && k + 32 <= n)) ring_copy16(row + ring_slot(lane + 32), p + 32);
      if (lane == 0 && (!EDGE || (k + 64 >= 0 && k + 64 <= n))) ring_copy16(row + ring_slot(64), p + 64);
    }
    ring_commit();
  }

  // ------------------------------------------------------------------ start-up ----
  // wsm: the warp's block of shared memory (kWarpE4 elements)
  static RN_HD void init(const StreamParams& sp, const MarchCtx& cx, MarchLane& L, E4* wsm, int lane, bool poison)
  {
    const int u0 = cx.a0 - 1;
    const int k0 = cx.L0 + 2 * lane;
    const uint32_t n = (uint32_t)sp.lev.n;
    const float z = poison ? march_nan() : 0.0f;
    const E4 z4 = E4{z, z, z, z};
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int e = 0; e < S_COUNT; e++) *state_at(wsm, b, e, lane) = z4;
#pragma unroll
    for (int q = 0; q < 3; q++) { L.cen[q] = z4; L.td1[q] = z; L.td2[q] = z; }
#pragma unroll
    for (int q = 0; q < 2; q++) { L.sg[q] = z; L.m1[q] = z4; }
#pragma unroll
    for (int q = 0; q < 4; q++) { L.ts[q] = z; L.out[0][q] = z4; L.out[1][q] = z4; }
    L.sS = lcg_skip(cx.seed, sp.lev.kS + (uint32_t)u0 * n + (uint32_t)k0);
    L.sD1 = lcg_skip(cx.seed, sp.lev.kD1 + (uint32_t)u0 * n + (uint32_t)k0);
    L.sD2 = lcg_skip(cx.seed, sp.lev.kD2 + (uint32_t)(u0 - 1) * (n + 1u) + (uint32_t)k0);
    const int Wout = 1 << sp.dout;
    L.ok = 0;
    float fy2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gj = 2 * k0 + j;
      const int b = k0 + (j >> 1);  // owning cell column
      const float fy = (float)(gj - Wout / 2);
      fy2[j] = fy * fy;
      const bool owned = (b >= cx.B0 && b < cx.B0 + cx.CW) || (gj == 2 * (cx.B0 + cx.CW) && cx.CWe != cx.CW);
      if (owned && (unsigned)(gj - sp.win_j0) < (unsigned)sp.win_nj) L.ok |= 1 << j;
    }
    wsm[kConstE4 + lane] = E4{fy2[0], fy2[1], fy2[2], fy2[3]};
    // (both pointers may start outside their grids; they are only dereferenced where the row / column exists)
    L.src = cx.gin + ((ptrdiff_t)u0 * (ptrdiff_t)sp.gin_pitch + (ptrdiff_t)(cx.L0 + lane));
    L.dst = cx.gout + ((ptrdiff_t)(2 * (u0 - 1) - sp.win_i0) * (ptrdiff_t)sp.gout_pitch + (ptrdiff_t)(2 * k0 - sp.win_j0));  // the first step finishes cell row u0 - 1
    for (int b = 0; b < 2; b++) {  // rows u0, u0+1: one commit group each; step u requests row u+2
      prefetch_row(sp, cx, wsm, L.src, u0 + b, lane);
      L.src += sp.gin_pitch;
    }
  }

  // ------------------------------------------------------------------ P1: S(u) ----
  static RN_HD void p1(const StreamParams& sp, const MarchCtx& cx, MarchLane& L, E4* wsm, int u, int lane)
  {
    // lattice row u+2 sets out now (into the slot row u-2 occupied; the ring keeps rows u-1 .. u+2); rows u and u+1 have
    // landed once at most this one group is still in flight
    prefetch_row(sp, cx, wsm, L.src, u + 2, lane);
    L.src += sp.gin_pitch;
    ring_wait_all_but_1();
    const E4* ra = ring_row(wsm, u) + lane;
    const E4* rb = ring_row(wsm, u + 1) + lane;
    const E4 ka[3] = {ring_at(ra, 0), ring_at(ra, 1), ring_at(ra, 2)}, kb[3] = {ring_at(rb, 0), ring_at(rb, 1), ring_at(rb, 2)};
    const uint32_t sB = kLcgA * L.sS + kLcgC;
    const uint32_t draw_a = lcg_draw_from_state(L.sS), draw_b = lcg_draw_from_state(sB);
    L.sS = jump_apply(sp.jn, L.sS);
    if (!EDGE) {
      if (RN_MARCH_CARRY & 4) { L.c_ka[0] = ka[0]; L.c_ka[1] = ka[1]; L.c_ka[2] = ka[2]; }
      const PairWork w = centre_pair_begin(sp.law, ka[0], kb[0], ka[1], kb[1], ka[1], kb[1], ka[2], kb[2], sp.lev.blahS, draw_a, draw_b);
      if (warp_all(sp.law.flat_ok && vnorm2_both_degenerate(w.f))) pair_finish<true>(sp.law, w, sp.lev.blahS, &L.cen[1], &L.cen[2], &L.ts[1], &L.ts[2]);
      else pair_finish<false>(sp.law, w, sp.lev.blahS, &L.cen[1], &L.cen[2], &L.ts[1], &L.ts[2]);
    } else {
      const int n = sp.lev.n;
      const int k0 = cx.L0 + 2 * lane;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int bg = k0 + q;
        float loss = 0.0f;
        if (u >= 0 && u < n && bg >= 0 && bg < n) {
          const E4 k00 = ka[q], k01 = ka[q + 1], k10 = kb[q], k11 = kb[q + 1];
          const E4 a4 = avg4_packed(k00, k10, k01, k11);
          const Displaced dp = displace<true>(sp.law, pos4(a4), vsub(pos4(k11), pos4(k00)), vsub(pos4(k10), pos4(k01)), sp.lev.blahS, q ? draw_b : draw_a);
          L.cen[1 + q] = e4_of(dp.pos, a4.w + shade_gain(sp.law, dp.offset));
          loss = shade_loss(sp.law, dp.offset);
        }
        L.ts[1 + q] = loss;
      }
    }
  }
  // X1 + the step's centres and generateSquare losses go to the bank of step u
  static RN_HD void x1(const MarchCtx& cx, MarchLane& L, E4* wsm, int u, int lane, const MarchLane* left, const MarchLane* right)
  {
    L.ts[0] = from_left(L.ts[2], left ? &left->ts[2] : nullptr);
    L.ts[3] = from_right(L.ts[1], right ? &right->ts[1] : nullptr);
    L.cen[0] = from_left4(L.cen[2], left ? &left->cen[2] : nullptr);
    const int b = bank_of(cx, u);
    *state_at(wsm, b, S_C0, lane) = L.cen[0];
    *state_at(wsm, b, S_C1, lane) = L.cen[1];
    *state_at(wsm, b, S_C2, lane) = L.cen[2];
    *state_at(wsm, b, S_TS, lane) = E4{L.ts[0], L.ts[1], L.ts[2], L.ts[3]};
  }

  // ------------------------------------------------------------------ P2: D(u) ----
  static RN_HD void p2(const StreamParams& sp, const MarchCtx& cx, MarchLane& L, E4* wsm, int u, int lane)
  {
    const int pb = bank_of(cx, u) ^ 1;  // what step u-1 left
    const E4* ra = ring_row(wsm, u) + lane;
    const E4* rm = ring_row(wsm, u - 1) + lane;
    const E4 tsm = *state_at(wsm, pb, S_TS, lane);
    L.tdm = *state_at(wsm, pb, S_TD, lane);  // kept for P3
    L.tem = *state_at(wsm, pb, S_TE, lane);
    const E4 cenm[3] = {*state_at(wsm, pb, S_C0, lane), *state_at(wsm, pb, S_C1, lane), *state_at(wsm, pb, S_C2, lane)};
    const float* tsc = L.ts;
    const E4* cenc = L.cen;
    // corner shades as generateSquare left them (Terrain.cpp:252-255): minus the scatter of the 4 surrounding centres,
    // in call order (u-1, k-1), (u-1, k), (u, k-1), (u, k)
    const bool cka = !EDGE && (RN_MARCH_CARRY & 4);
    E4 K0 = cka ? L.c_ka[0] : ring_at(ra, 0), K1 = cka ? L.c_ka[1] : ring_at(ra, 1), K2 = cka ? L.c_ka[2] : ring_at(ra, 2);
    K0.w = (((K0.w - tsm.x) - tsm.y) - tsc[0]) - tsc[1];
    K1.w = (((K1.w - tsm.y) - tsm.z) - tsc[1]) - tsc[2];
    K2.w = (((K2.w - tsm.z) - tsm.w) - tsc[2]) - tsc[3];
    // K'(u-1, k0/k1): position from the ring, shade as step u-1 computed it
    E4 kcm[2] = {ring_at(rm, 0), ring_at(rm, 1)};
    if (!EDGE && (RN_MARCH_CARRY & 1)) { L.c_km[0] = kcm[0]; L.c_km[1] = kcm[1]; }
    if (!EDGE && (RN_MARCH_CARRY & 2)) { L.c_cm[0] = cenm[1]; L.c_cm[1] = cenm[2]; }
    kcm[0].w = L.tdm.w;
    kcm[1].w = L.tem.w;
    const uint32_t s1B = kLcgA * L.sD1 + kLcgC, s2B = kLcgA * L.sD2 + kLcgC;
    const uint32_t d1a = lcg_draw_from_state(L.sD1), d1b = lcg_draw_from_state(s1B);
    const uint32_t d2a = lcg_draw_from_state(L.sD2), d2b = lcg_draw_from_state(s2B);
    L.sD1 = jump_apply(sp.jn, L.sD1);
    L.sD2 = jump_apply(sp.jn1, L.sD2);
    if (!EDGE) {
      // loop 1: M1(u, k) at output (2u, 2k+1): x neighbours C(u-1, k), C(u, k); y neighbours K'(u, k), K'(u, k+1)
      const PairWork w1 = mid_pair_begin(sp.law, cenm[1], cenc[1], K0, K1, cenm[2], cenc[2], K1, K2, sp.lev.blahD, d1a, d1b);
      // loop 2: M2(u-1, k) at output (2u-1, 2k): x neighbours K'(u-1, k), K'(u, k); y neighbours C(u-1, k-1), C(u-1, k)
      const PairWork w2 = mid_pair_begin(sp.law, kcm[0], K0, cenm[0], cenm[1], kcm[1], K1, cenm[1], cenm[2], sp.lev.blahD, d2a, d2b);
      // one decision for the four midpoints and the whole warp.  Lane 0's M2(u-1, k0) is built on C(u-1, k0-1), which a halo lane does
      // not have (its sample is never stored nor passed on): it does not vote.
      const bool flat2 = f2hi(w2.f) < 1e-10f && (lane == 0 || f2lo(w2.f) < 1e-10f);
      if (warp_all(sp.law.flat_ok && vnorm2_both_degenerate(w1.f) && flat2)) {
        pair_finish<true>(sp.law, w1, sp.lev.blahD, &L.m1[0], &L.m1[1], &L.td1[1], &L.td1[2]);
        pair_finish<true>(sp.law, w2, sp.lev.blahD, &L.out[1][0], &L.out[1][2], &L.td2[0], &L.td2[1]);
      } else {
        pair_finish<false>(sp.law, w1, sp.lev.blahD, &L.m1[0], &L.m1[1], &L.td1[1], &L.td1[2]);
        pair_finish<false>(sp.law, w2, sp.lev.blahD, &L.out[1][0], &L.out[1][2], &L.td2[0], &L.td2[1]);
      }
    } else {
      const int n = sp.lev.n;
      const int k0 = cx.L0 + 2 * lane;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int bg = k0 + q;
        const bool v1 = u >= 0 && u <= n && bg >= 0 && bg < n;
        const bool v2 = u - 1 >= 0 && u - 1 < n && bg >= 0 && bg <= n;
        const E4 k00 = q ? K1 : K0, k01 = q ? K2 : K1;
        float loss1 = 0.0f, loss2 = 0.0f;
        if (v1) {
          if (u == 0 || u == n) {
            const E4 dummy = k00;
            L.m1[q] = mid_border(sp.law, u == 0 ? dummy : cenm[1 + q], u == n ? dummy : cenc[1 + q], k00, k01, u == 0, u == n, false, false, sp.lev.blahDh,
                                 q ? d1b : d1a, &loss1);
          } else {
            L.m1[q] = mid_interior(sp.law, cenm[1 + q], cenc[1 + q], k00, k01, sp.lev.blahD, q ? d1b : d1a, &loss1);
          }
        }
        if (v2) L.out[1][2 * q] = mid_select_y(sp.law, kcm[q], k00, cenm[q], cenm[1 + q], bg == 0, bg == n, sp.lev.blahD, q ? d2b : d2a, &loss2);
        L.td1[1 + q] = loss1;
        L.td2[q] = loss2;
      }
    }
    L.sg[0] = K0.w;
    L.sg[1] = K1.w;
  }
  static RN_HD void skip_p2(const StreamParams& sp, MarchLane& L)
  {
    L.sD1 = jump_apply(sp.jn, L.sD1);
    L.sD2 = jump_apply(sp.jn1, L.sD2);
  }
  static RN_HD void skip_p3(const StreamParams& sp, MarchLane& L) { L.dst += 2 * sp.gout_pitch; }
  // X2 + the step's loop-1 mids, generateDiamond losses and corner shades go to the bank of step u
  static RN_HD void x2(const MarchCtx& cx, MarchLane& L, E4* wsm, int u, int lane, const MarchLane* left, const MarchLane* right)
  {
    L.td1[0] = from_left(L.td1[2], left ? &left->td1[2] : nullptr);
    L.td2[2] = from_right(L.td2[0], right ? &right->td2[0] : nullptr);
    const int b = bank_of(cx, u);
    *state_at(wsm, b, S_M1A, lane) = L.m1[0];
    *state_at(wsm, b, S_M1B, lane) = L.m1[1];
    *state_at(wsm, b, S_TD, lane) = E4{L.td1[0], L.td1[1], L.td1[2], L.sg[0]};
    *state_at(wsm, b, S_TE, lane) = E4{L.td2[0], L.td2[1], L.td2[2], L.sg[1]};
  }

  // ------------------------------------------------------------------ P3: G(u-1), addVolcano ----
  // VM: what the caller knows about addVolcano at compile time -- -1 nothing (sp.volcano and cx.far decide), 0 off, 1 on and every
  // sample of the warp outside the crater, 2 on with the crater branch.  Known modes leave no branch between the shade sums and
  // the reciprocal square roots, so the compiler interleaves the two dependent chains.
  template <int VM>
  static RN_HD void p3(const StreamParams& sp, const MarchCtx& cx, MarchLane& L, E4* wsm, int u, int lane)
  {
    const bool volcano_on = VM < 0 ? (sp.volcano != 0) : VM > 0;
    const bool far_away = VM < 0 ? (cx.far != 0) : VM == 1;
    const int pb = bank_of(cx, u) ^ 1;
    const E4 tdm = L.tdm;  // TD1(u-1; k0-1, k0, k1), sg(u-1, k0)
    const E4 tem = L.tem;  // TD2(u-2; k0, k1, .), sg(u-1, k1)
    const float* tdc = L.td1;
    const float* t2 = L.td2;
    // corner K(u-1, k): generateSquare's scatter (carried), then loop 1 (same row: k-1, k), loop 2 (same column: u-2, u-1)
    const E4* rm = ring_row(wsm, u - 1) + lane;
    const bool ckm = !EDGE && (RN_MARCH_CARRY & 1), ccm = !EDGE && (RN_MARCH_CARRY & 2);
    E4 c0 = ckm ? L.c_km[0] : ring_at(rm, 0), c1 = ckm ? L.c_km[1] : ring_at(rm, 1);
    c0.w = (((tdm.w - tdm.x) - tdm.y) - tem.x) - t2[0];
    c1.w = (((tem.w - tdm.y) - tdm.z) - tem.y) - t2[1];
    // centre C(u-1, k): loop-1 mids above and below, loop-2 mids left and right
    E4 e0 = ccm ? L.c_cm[0] : *state_at(wsm, pb, S_C1, lane), e1 = ccm ? L.c_cm[1] : *state_at(wsm, pb, S_C2, lane);
    e0.w = (((e0.w - tdm.y) - tdc[1]) - t2[0]) - t2[1];
    e1.w = (((e1.w - tdm.z) - tdc[2]) - t2[1]) - t2[2];
    L.out[0][0] = c0;
    L.out[0][1] = *state_at(wsm, pb, S_M1A, lane);
    L.out[0][2] = c1;
    L.out[0][3] = *state_at(wsm, pb, S_M1B, lane);
    L.out[1][1] = e0;
    L.out[1][3] = e1;
    if (volcano_on && !(RN_EXP & 2)) {  // uniform
      const int Wout = 1 << sp.dout;
      const float gain_w = sp.volcano_gain * (float)Wout;
      const E4 fy2 = wsm[kConstE4 + lane];
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const float fx = (float)(2 * (u - 1) + r - Wout / 2);
        const float fx2 = fx * fx;
        if (!EDGE) {  // two columns at once (packed pair); the few warps that touch the crater evaluate its branch as a select
          const F2 r01 = f2add(f2s(fx2), f2(fy2.x, fy2.y)), r23 = f2add(f2s(fx2), f2(fy2.z, fy2.w));
          F2 h01, h23;
          if (far_away) {  // uniform
            h01 = f2mul(rcp_rn_normal2(sqrt_rn_normal2(r01)), f2s(gain_w));
            h23 = f2mul(rcp_rn_normal2(sqrt_rn_normal2(r23)), f2s(gain_w));
          } else {
            const float inv_w = 1.0f / (float)Wout;
            h01 = volcano_pair_general(r01, inv_w, sp.volcano_gain, gain_w, sp.law.one);
            h23 = volcano_pair_general(r23, inv_w, sp.volcano_gain, gain_w, sp.law.one);
          }
          L.out[r][0].y = L.out[r][0].y + f2lo(h01);
          L.out[r][1].y = L.out[r][1].y + f2hi(h01);
          L.out[r][2].y = L.out[r][2].y + f2lo(h23);
          L.out[r][3].y = L.out[r][3].y + f2hi(h23);
        } else {
          const float inv_w = 1.0f / (float)Wout;
          L.out[r][0].y = L.out[r][0].y + volcano_height_r2<false>(fx2 + fy2.x, inv_w, sp.volcano_gain, gain_w);
          L.out[r][1].y = L.out[r][1].y + volcano_height_r2<false>(fx2 + fy2.y, inv_w, sp.volcano_gain, gain_w);
          L.out[r][2].y = L.out[r][2].y + volcano_height_r2<false>(fx2 + fy2.z, inv_w, sp.volcano_gain, gain_w);
          L.out[r][3].y = L.out[r][3].y + volcano_height_r2<false>(fx2 + fy2.w, inv_w, sp.volcano_gain, gain_w);
        }
      }
    }
  }

  // ------------------------------------------------------------------ store of cell row a = u-1 ----
  // `left`: host replay only (the lane to the left); on the device the shifted rows shuffle.
  static RN_HD void store(const StreamParams& sp, const MarchCtx& cx, MarchLane& L, int u, int lane, const MarchLane* left)
  {
    const int a = u - 1;
    const int Wout = 1 << sp.dout;
    E4* dst = L.dst;
    L.dst = dst + 2 * sp.gout_pitch;
    if (a < cx.a0 || a >= cx.a1e) return;  // uniform
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int gi = 2 * a + r;
      if (EDGE && ((unsigned)(gi - sp.win_i0) >= (unsigned)sp.win_ni || gi > Wout)) continue;  // uniform; lean steps: march_ctx
      E4* row = r == 0 ? dst : dst + sp.gout_pitch;
      const E4* v = L.out[r];
      if (!EDGE) {  // lean warps own 60 whole cell columns inside the window
        const bool mine = lane >= 1 && lane <= 30 && (!(RN_EXP & 1) || sp.law.one != 1.0f);
        if ((reinterpret_cast<uintptr_t>(row) & 16u) == 0) {  // uniform: the lane's four samples are two whole sectors
          if (mine) {
            st256(row, v[0], v[1]);
            st256(row + 2, v[2], v[3]);
          }
        } else {  // sectors straddle the lanes: take the left neighbour's last sample, leave the own last one to the right neighbour
          const E4 lv = from_left4(v[3], left ? &left->out[r][3] : nullptr);
          if (mine) {
            if (lane == 1) st128(row, v[0]);
            else st256(row - 1, lv, v[0]);
            st256(row + 1, v[1], v[2]);
            if (lane == 30) st128(row + 3, v[3]);
          }
        }
      } else {
        // closing row W (a == n) and closing column W hold corners and one mid class only; `ok` knows the columns
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (L.ok & (1 << j)) row[j] = v[j];
      }
    }
  }
};

}  // namespace rn
