// rn_tile.cuh -- algorithm 2: the fused tile form of the diamond-square path.
//
// One CTA refines one T x T output tile through L consecutive levels of
// Terrain.cpp:60-82 entirely in shared memory and writes the finished tile once,
// as 128-bit coalesced stores.  Everything a tile needs beyond its own samples (the
// dependency cone of its border) is RECOMPUTED from the input lattice, never
// exchanged, so tiles -- and therefore slabs, windows and GPUs -- are independent.
//
// Vocabulary (all per level, s = step, h = s/2, n = W/s squares per side):
//   corner K(a,b)  lattice point (a*s, b*s)          -- exists before the level
//   centre C(a,b)  (a*s+h, b*s+h)   averageSquare + generateSquare   Terrain.cpp:216-258
//   mid    M1(a,b) (a*s,   b*s+h)   averageDiamond + generateDiamond Terrain.cpp:144-214, loop 1 (:69-71,:75-77)
//   mid    M2(a,b) (a*s+h, b*s)     same, loop 2 (:72-74,:78-80)
//   "cell" (a,b) = {K, C, M1, M2}(a,b): one thread works on one cell per phase.
// Shade is accumulated by ORDERED GATHER, so it is bit-identical to the
// reference's sequential scatter; each new point stores its scatter term
// (0.25f*offset)*radiosityStrength ("loss", Terrain.cpp:208-211,252-255) in a float
// plane, with 0.0f for points that do not exist (x - 0.0f == x), which makes every
// gather branch-free.
//
// Per level three barrier-separated phases:
//   S  centres:  reads K (pos+shade of the previous level), writes C and its loss TS
//   D  mids:     reads K (shade corrected on the fly by the 4 surrounding TS), C; writes M1, M2, losses TD1, TD2
//   G  gather:   K.shade -= 4 TS, 4 TD ; C.shade -= 4 TD   (reference call order)
// The levels above the last run in place on the step-2 lattice LAT; the last level
// (75 % of the work) reads LAT with unit stride and writes C / M1 / M2 into their own
// unit-stride arrays; a final phase interleaves the four classes into full output
// rows (+ the addVolcano epilogue, Terrain.cpp:121-141) for the coalesced store.
//
// Halo analysis (fine units, tile = [0,T]): after a level with half-step h, samples
// within extent a_h = 2h-2 of the tile are final.  That needs mids within 3h-2,
// centres within 4h-2, input shade within 4h-2 and input pos within 5h-2; the
// coarsest fused level has h = 2^(L-1), hence the halo H = stage_halo(L).  Samples
// outside those extents may hold garbage; it never propagates inwards
// (tests/host_replay fills shared memory with NaN to prove it).
//
// The file compiles as plain C++ too: tests/host_replay replays the phases thread by
// thread on the CPU to validate the tiling logic without a GPU.  That replay is test
// scaffolding and is never linked into the product.
#pragma once
#include "rn_math.cuh"

namespace rn {

#if defined(__CUDACC__)
typedef float4 E4;
typedef uint2 uint2_t;
#else
struct alignas(16) E4 {
  float x, y, z, w;
};
struct alignas(8) uint2_t {
  uint32_t x, y;
};
#endif

RN_HD E4 e4_of(V3 p, float w)
{
  E4 e;
  e.x = p.x; e.y = p.y; e.z = p.z; e.w = w;
  return e;
}
RN_HD V3 pos4(const E4& e) { return V3{e.x, e.y, e.z}; }

constexpr int kMaxStageLevels = 4;

// Per-level constants, resolved on the host (rn_plan.h).
struct StageLevel {
  int n;                       // squares per side
  uint32_t kS, kD1, kD2;       // 1 + call index of the level's first generateSquare / loop-1 / loop-2 draw, mod 2^32
  float blahS, blahD, blahDh;  // Terrain.cpp:231 ; :186 unclamped ; :186 with the X extent clamped
};

struct StageParams {
  const E4* gin;        // complete depth-din grid (row pitch gin_pitch); nullptr = analytic initial lattice (din == 2)
  size_t gin_pitch, gin_field_stride;
  E4* gout;             // points at element (win_i0, win_j0) of the depth-dout grid
  size_t gout_pitch, gout_field_stride;
  int din, dout;
  int win_i0, win_j0, win_ni, win_nj;  // the part of the output grid that is stored
  int tile_i0, tile_j0;                // tile index of blockIdx (y, x) = (0, 0)
  uint32_t seed;
  const uint32_t* seeds;               // batch: seed of field blockIdx.z (else nullptr)
  Law law;
  StageLevel lev[kMaxStageLevels];
  int volcano;                         // addVolcano epilogue (final stage only)
#ifdef RN_TIMELINE
  int tl_slot;          // instrumented builds: launch number (profiles/micro/timeline.py)
#endif
  float volcano_gain;                  // amplitude*0.008f, Terrain.cpp:137
};

#if defined(__CUDACC__)
#define RN_CX __host__ __device__ constexpr
#else
#define RN_CX constexpr
#endif
RN_CX int stage_halo(int L) { return ((5 * (1 << (L - 1)) - 2) >> L) << L; }
RN_CX int cdiv_(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }
RN_CX int fdiv_(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
RN_CX int imin_(int a, int b) { return a < b ? a : b; }
RN_CX int imax_(int a, int b) { return a > b ? a : b; }
RN_CX int mod8_(int a) { return ((a % 8) + 8) % 8; }

// Cell ranges of the three phases of a level with half-step h (fine units) for a tile [0,T].
template <int h, int T>
struct LevelRange {
  static constexpr int s = 2 * h;
  static constexpr int S_lo = cdiv_(-5 * h + 2, s), S_hi = fdiv_(T + 3 * h - 2, s);      // centres within 4h-2
  static constexpr int MA_lo = cdiv_(-(3 * h - 2), s), MA_hi = fdiv_(T + 3 * h - 2, s);  // on-lattice coordinate of a mid within 3h-2
  static constexpr int MB_lo = cdiv_(-4 * h + 2, s), MB_hi = fdiv_(T + 2 * h - 2, s);    // half-offset coordinate of a mid within 3h-2
  static constexpr int D_lo = imin_(MA_lo, MB_lo), D_hi = imax_(MA_hi, MB_hi);
  static constexpr int GK_lo = cdiv_(-(2 * h - 2), s), GK_hi = fdiv_(T + 2 * h - 2, s);  // corners within 2h-2
  static constexpr int GC_lo = cdiv_(-3 * h + 2, s), GC_hi = fdiv_(T + h - 2, s);        // centres within 2h-2
  static constexpr int G_lo = imin_(GK_lo, GC_lo), G_hi = imax_(GK_hi, GC_hi);
  static constexpr int NS = S_hi - S_lo + 1, ND = D_hi - D_lo + 1, NG = G_hi - G_lo + 1;
};

// Shared-memory geometry of a stage (offsets in 4-byte words).
//
// LAT, the step-2 lattice the in-place levels work on, is stored as two halves by column parity
// (even columns, then odd columns, each NLAT x PH).  A level with half-step h >= 2 then touches
// unit-stride runs of one half (no bank conflicts for h == 2, 2-way for h == 4), and the last
// level, which reads consecutive columns, alternates between the halves; with the half offset
// congruent to 4 (mod 8) 16-byte slots a quarter warp's LDS.128 covers 8 distinct bank groups.
template <int L, int T>
struct TileGeo {
  static_assert(L >= 1 && L <= kMaxStageLevels, "1..4 fused levels");
  static_assert(T % (1 << L) == 0 && T % 2 == 0, "tile edge must be a multiple of the coarsest fused step");
  static constexpr int H = stage_halo(L);        // halo, fine units
  static constexpr int HL = H / 2;               // halo, LAT units
  static constexpr int NLAT = (T + 2 * H) / 2 + 1;
  static constexpr int PH = (NLAT + 1) / 2;      // row pitch of one LAT half (E4)
  static constexpr int LATOFF = NLAT * PH + mod8_(4 - NLAT * PH);  // E4 offset of the odd-column half, == 4 (mod 8)
  static constexpr int LATSLOTS = LATOFF + NLAT * PH;
  static constexpr int NC = T / 2 + 2;           // last-level class arrays: cells -1 .. T/2
  static constexpr int PC = NC;                  // row pitch (E4 / float)
  static constexpr int NTS = NC + 2;             // TS keeps a second ring (cells -2 .. T/2+1)
  static constexpr int PTS = NTS;
  static constexpr int NR = NC + 1;              // rows of an RNG row table, columns of the jump table
  // E4-slot offsets
  static constexpr int LAT4 = 0;
  static constexpr int CARR4 = LAT4 + LATSLOTS;
  static constexpr int M1ARR4 = CARR4 + NC * PC;
  static constexpr int M2ARR4 = M1ARR4 + NC * PC;
  // word offsets
  static constexpr int LAT = 4 * LAT4, CARR = 4 * CARR4, M1ARR = 4 * M1ARR4, M2ARR = 4 * M2ARR4;
  static constexpr int TL = M1ARR;  // loss plane of the in-place levels (indexed like LAT): dead before M1/M2 are written
  static_assert(LATSLOTS <= 4 * 2 * NC * PC, "in-place loss plane must fit in the M1/M2 arrays");
  static constexpr int TS = M2ARR + 4 * NC * PC;
  static constexpr int TD1 = TS + NTS * PTS;
  static constexpr int TD2 = TD1 + NC * PC;
  static constexpr int JUMP = ((TD2 + NC * PC + 1) / 2) * 2;  // NR x (a, c), 8-byte aligned
  static constexpr int ROWS = JUMP + 2 * NR;                  // L x 3 x NR row start states
  static constexpr int WORDS = ((ROWS + L * 3 * NR + 3) / 4) * 4;
  static constexpr size_t BYTES = (size_t)WORDS * 4;
  // slot of LAT point (row p, column q)
  static RN_HD int lat_slot(int p, int q) { return (q & 1) * LATOFF + p * PH + (q >> 1); }
};

// Addressing of one level.  h == 1: the last level (class arrays); h >= 2: in place on LAT.
template <int L, int T, int h>
struct View {
  using G = TileGeo<L, T>;
  float* sm;
  RN_HD explicit View(float* s) : sm(s) {}
  RN_HD E4* e4(int off) const { return reinterpret_cast<E4*>(sm + off); }
  // LAT slot of the lattice point (a*h + dp, b*h + dq) in cell units + LAT offsets
  RN_HD int lidx(int a, int b, int dp, int dq) const { return G::lat_slot(a * h + G::HL + dp, b * h + G::HL + dq); }
  RN_HD int cidx(int a, int b) const { return (a + 1) * G::PC + (b + 1); }
  RN_HD E4& K(int a, int b) const { return e4(G::LAT)[lidx(a, b, 0, 0)]; }
  RN_HD E4& C(int a, int b) const
  {
    if constexpr (h == 1) return e4(G::CARR)[cidx(a, b)];
    else return e4(G::LAT)[lidx(a, b, h / 2, h / 2)];
  }
  RN_HD E4& M1(int a, int b) const
  {
    if constexpr (h == 1) return e4(G::M1ARR)[cidx(a, b)];
    else return e4(G::LAT)[lidx(a, b, 0, h / 2)];
  }
  RN_HD E4& M2(int a, int b) const
  {
    if constexpr (h == 1) return e4(G::M2ARR)[cidx(a, b)];
    else return e4(G::LAT)[lidx(a, b, h / 2, 0)];
  }
  RN_HD float& TS(int a, int b) const
  {
    if constexpr (h == 1) return sm[G::TS + (a + 2) * G::PTS + (b + 2)];
    else return sm[G::TL + lidx(a, b, h / 2, h / 2)];
  }
  RN_HD float& TD1(int a, int b) const
  {
    if constexpr (h == 1) return sm[G::TD1 + cidx(a, b)];
    else return sm[G::TL + lidx(a, b, 0, h / 2)];
  }
  RN_HD float& TD2(int a, int b) const
  {
    if constexpr (h == 1) return sm[G::TD2 + cidx(a, b)];
    else return sm[G::TL + lidx(a, b, h / 2, 0)];
  }
};

RN_HD uint32_t lcg_state_after32(uint32_t seed, uint32_t k)
{
  Jump sq{kLcgA, kLcgC};
  uint32_t s = seed;
  while (k) {
    if (k & 1u) s = jump_apply(sq, s);
    sq = jump_compose(sq, sq);
    k >>= 1;
  }
  return s;
}

// The LCG jump over 2^B draws, folded at compile time: (A^(2^B), C * (A^(2^B) - 1) / (A - 1)) mod 2^32.
template <int B>
struct LcgPow2 {
  static constexpr uint32_t a = LcgPow2<B - 1>::a * LcgPow2<B - 1>::a;
  static constexpr uint32_t c = (LcgPow2<B - 1>::a + 1u) * LcgPow2<B - 1>::c;
};
template <>
struct LcgPow2<0> {
  static constexpr uint32_t a = kLcgA, c = kLcgC;
};
// State after k draws from `s`: one multiply-add with immediate constants per set bit of k
// (a chain of at most 32 IMADs, no squaring chain, no table loads).
template <int B = 0>
RN_HD uint32_t lcg_skip(uint32_t s, uint32_t k)
{
  if (k & (1u << B)) s = LcgPow2<B>::a * s + LcgPow2<B>::c;
  if constexpr (B < 31) return lcg_skip<B + 1>(s, k);
  else return s;
}

// addVolcano, Terrain.cpp:129-137, for one sample whose squared index distance from the volcano is
// r2 = fx*fx + fy*fy.  d /= maxWidth is a multiplication by the exact power of two inv_w.
// r2 is either 0 (the centre sample) or >= 1, so sqrt / reciprocal see normal numbers.
// FAR: the caller guarantees d >= minRad (the tile lies outside the crater), so the crater
// branch (:133-136) is not compiled in.
// gain_w = gain * W (exact: W is a power of two).
template <bool FAR>
RN_HD float volcano_height_r2(float r2, float inv_w, float gain, float gain_w)
{
  if (FAR) {
    // height = 1/(sqrt(r2)/W) = W * (1/sqrt(r2)) exactly, and scaling the gain by W instead is exact too:
    // one multiplication less, same bits.
    return rcp_rn_normal(sqrt_rn_normal(r2)) * gain_w;
  }
  float d = sqrt_rn_normal(r2) * inv_w;
  float height = rcp_rn_normal(d);
  const float minRad = 0.02f;
  if (!(d >= minRad)) {  // also taken by the centre sample, where the approximations above see 0
    if (r2 == 0.0f) d = 0.0f;
    height = 60000.f * (d * d - minRad * minRad) + (1.0f / minRad);
  }
  return height * gain;
}

// Border form of averageDiamond + generateDiamond (one clamped neighbour is the midpoint
// itself: fresh average as position, untouched 1.0f as shade; Terrain.cpp:149-167,175-182).
RN_HD E4 mid_border(const Law& law, E4 ex0, E4 ex1, E4 ey0, E4 ey1, bool sx0, bool sx1, bool sy0, bool sy1, float blah,
                           uint32_t draw, float* loss_out)
{
  const bool xb = sx0 || sx1;
  if (sx0) ex0.w = 1.0f;
  if (sx1) ex1.w = 1.0f;
  if (sy0) ey0.w = 1.0f;
  if (sy1) ey1.w = 1.0f;
  V3 avg;
  if (xb) avg = vmul(vadd(pos4(ey0), pos4(ey1)), 0.5f);
  else avg = vmul(vadd(pos4(ex0), pos4(ex1)), 0.5f);
  float sh = (ex0.w + ex1.w + ey0.w + ey1.w) * 0.25f;
  const V3 px0 = sx0 ? avg : pos4(ex0), px1 = sx1 ? avg : pos4(ex1);
  const V3 py0 = sy0 ? avg : pos4(ey0), py1 = sy1 ? avg : pos4(ey1);
  const Displaced dp = displace<true>(law, avg, vsub(px0, px1), vsub(py1, py0), blah, draw);
  const float loss = shade_loss(law, dp.offset);
  sh += shade_gain(law, dp.offset);
  sh -= loss;  // the scatter onto the clamped index lands on the midpoint itself
  *loss_out = loss;
  return e4_of(dp.pos, sh);
}

// mid_interior, or mid_border for a missing y neighbour (sy0 / sy1), chosen by selects instead of a branch: the one
// thread of a warp that sits on a lattice column border then shares its warp's instruction stream.
RN_HD E4 mid_select_y(const Law& law, const E4& ex0, const E4& ex1, E4 ey0, E4 ey1, bool sy0, bool sy1, float blah, uint32_t draw,
                      float* loss_out)
{
  const bool yb = sy0 || sy1;
  if (sy0) ey0.w = 1.0f;
  if (sy1) ey1.w = 1.0f;
  const V3 sx = vadd(pos4(ex0), pos4(ex1));
  const V3 avg_i = vmul(vadd(vadd(sx, pos4(ey0)), pos4(ey1)), 0.25f);
  const V3 avg_b = vmul(sx, 0.5f);
  const V3 avg = yb ? avg_b : avg_i;
  float sh = (ex0.w + ex1.w + ey0.w + ey1.w) * 0.25f;
  const V3 py0 = sy0 ? avg : pos4(ey0), py1 = sy1 ? avg : pos4(ey1);
  const Displaced dp = displace<true>(law, avg, vsub(pos4(ex0), pos4(ex1)), vsub(py1, py0), blah, draw);
  const float loss = shade_loss(law, dp.offset);
  sh += shade_gain(law, dp.offset);
  if (yb) sh -= loss;  // the scatter onto the clamped index lands on the midpoint itself
  *loss_out = loss;
  return e4_of(dp.pos, sh);
}

RN_HD E4 mid_interior(const Law& law, const E4& ex0, const E4& ex1, const E4& ey0, const E4& ey1, float blah, uint32_t draw,
                             float* loss_out)
{
  const V3 avg = vmul(vadd(vadd(vadd(pos4(ex0), pos4(ex1)), pos4(ey0)), pos4(ey1)), 0.25f);
  const float sh = (ex0.w + ex1.w + ey0.w + ey1.w) * 0.25f;
  const Displaced dp = displace<true>(law, avg, vsub(pos4(ex0), pos4(ex1)), vsub(pos4(ey1), pos4(ey0)), blah, draw);
  *loss_out = shade_loss(law, dp.offset);
  return e4_of(dp.pos, sh + shade_gain(law, dp.offset));
}

// Two interior midpoints at once (mid_interior twice, bit for bit): the component-wise averages run as packed pairs
// inside each float4, everything from the cross product on as packed pairs across the two samples.
RN_HD F2 e4_xy(const E4& e) { return f2(e.x, e.y); }
RN_HD F2 e4_zw(const E4& e) { return f2(e.z, e.w); }
// ((a + b) + c) + d) * 0.25 on all four components (positions and shade are averaged in the same order)
RN_HD E4 avg4_packed(const E4& a, const E4& b, const E4& c, const E4& d)
{
  const F2 q = f2s(0.25f);
  const F2 xy = f2mul(f2add(f2add(f2add(e4_xy(a), e4_xy(b)), e4_xy(c)), e4_xy(d)), q);
  const F2 zw = f2mul(f2add(f2add(f2add(e4_zw(a), e4_zw(b)), e4_zw(c)), e4_zw(d)), q);
  return E4{f2lo(xy), f2hi(xy), f2lo(zw), f2hi(zw)};
}
RN_HD void mid_interior2(const Law& law, const E4& ax0, const E4& ax1, const E4& ay0, const E4& ay1, const E4& bx0, const E4& bx1,
                         const E4& by0, const E4& by1, float blah, uint32_t draw_a, uint32_t draw_b, E4* out_a, E4* out_b, float* loss_a,
                         float* loss_b)
{
  const F2 q = f2s(0.25f);
  // (x, y, z, shade) averages: ((x0 + x1) + y0) + y1) * 0.25 per component
  const F2 a_xy = f2mul(f2add(f2add(f2add(e4_xy(ax0), e4_xy(ax1)), e4_xy(ay0)), e4_xy(ay1)), q);
  const F2 a_zw = f2mul(f2add(f2add(f2add(e4_zw(ax0), e4_zw(ax1)), e4_zw(ay0)), e4_zw(ay1)), q);
  const F2 b_xy = f2mul(f2add(f2add(f2add(e4_xy(bx0), e4_xy(bx1)), e4_xy(by0)), e4_xy(by1)), q);
  const F2 b_zw = f2mul(f2add(f2add(f2add(e4_zw(bx0), e4_zw(bx1)), e4_zw(by0)), e4_zw(by1)), q);
  const V3P dir1 = v3p(vsub(pos4(ax0), pos4(ax1)), vsub(pos4(bx0), pos4(bx1)));
  const V3P dir2 = v3p(vsub(pos4(ay1), pos4(ay0)), vsub(pos4(by1), pos4(by0)));
  const Displaced2 dp = displace2(law, f2lo(a_xy), f2hi(a_xy), f2lo(b_xy), f2hi(b_xy), dir1, dir2, blah, draw_a, draw_b);
  const F2 rs = f2s(law.radiosity_strength);
  const F2 gain = f2mul(dp.offset, rs);
  const F2 loss = f2mul(f2mul(f2s(0.25f), dp.offset), rs);
  *loss_a = f2lo(loss);
  *loss_b = f2hi(loss);
  *out_a = E4{f2lo(a_xy) + f2lo(dp.shift.x), f2hi(a_xy) + f2lo(dp.shift.y), f2lo(a_zw) + f2lo(dp.shift.z), f2hi(a_zw) + f2lo(gain)};
  *out_b = E4{f2lo(b_xy) + f2hi(dp.shift.x), f2hi(b_xy) + f2hi(dp.shift.y), f2lo(b_zw) + f2hi(dp.shift.z), f2hi(b_zw) + f2hi(gain)};
}


struct TileCtx {
  int I0, J0;        // output-grid coordinates of the tile origin
  uint32_t seed;
  const E4* gin;     // field-adjusted
  E4* gout;
};

// Visits the cells c = tid, tid+NT, ... of an N x N range as (row, column), updating the pair
// incrementally instead of dividing every time.
template <int N, int NT, class F>
RN_HD void for_cells(int tid, F&& f)
{
  constexpr int DA = NT / N, DB = NT % N;
  int ra = tid / N, rb = tid - ra * N;
  for (int c = tid; c < N * N; c += NT) {
    f(ra, rb);
    ra += DA;
    rb += DB;
    if (rb >= N) {
      rb -= N;
      ra++;
    }
  }
}

// EDGE = the tile's dependency cone may reach outside the global lattice: existence tests and the
// border forms of averageDiamond / generateDiamond are compiled in.  Interior tiles (the vast
// majority) use EDGE = false, where every cell of every range exists.
template <int L, int T, int NT, bool EDGE>
struct Tile {
  using G = TileGeo<L, T>;
  static constexpr int NPHASES = 2 + 3 * L;

  static RN_HD bool is_edge(const TileCtx& cx, int Wout)
  {
    return cx.I0 - G::H < 0 || cx.J0 - G::H < 0 || cx.I0 + T + G::H > Wout || cx.J0 + T + G::H > Wout;
  }

  // ------------------------------------------------------------------ phase 0 ----
  // Input lattice, LCG column jumps (A^c, C_c) and, per level and draw table (generateSquare,
  // loop 1, loop 2), the LCG state in front of the first cell of every cell row.
  template <int M>
  static RN_HD uint32_t row_state(const StageParams& sp, const TileCtx& cx, int e)
  {
    constexpr int h = 1 << (L - 1 - M);
    using R = LevelRange<h, T>;
    const StageLevel& lv = sp.lev[M];
    const int which = e / G::NR, r = e - which * G::NR;
    const int A0 = cx.I0 >> (L - M), B0 = cx.J0 >> (L - M);
    const uint32_t n = (uint32_t)lv.n;
    uint32_t k;
    if (which == 0) k = lv.kS + (uint32_t)(A0 + R::S_lo + r) * n + (uint32_t)(B0 + R::S_lo);
    else if (which == 1) k = lv.kD1 + (uint32_t)(A0 + R::D_lo + r) * n + (uint32_t)(B0 + R::D_lo);
    else k = lv.kD2 + (uint32_t)(A0 + R::D_lo + r) * (n + 1u) + (uint32_t)(B0 + R::D_lo);
    return lcg_skip(cx.seed, k);
  }
  template <int M>
  static RN_HD uint32_t row_state_dispatch(const StageParams& sp, const TileCtx& cx, int m, int e)
  {
    if constexpr (M + 1 < L) {
      if (m != M) return row_state_dispatch<M + 1>(sp, cx, m, e);
    }
    return row_state<M>(sp, cx, e);
  }

  static RN_HD void phase_init(const StageParams& sp, const TileCtx& cx, float* sm, int tid)
  {
    // input lattice: LAT indices that are multiples of 2^(L-1) <-> input grid points.  The global
    // loads are issued first so that their latency overlaps the RNG table arithmetic below.
    constexpr int NIN = (T + 2 * G::H) / (1 << L) + 1;
    constexpr int STEP = 1 << (L - 1);
    constexpr int NLOAD = (NIN * NIN + NT - 1) / NT;
    E4* lat = reinterpret_cast<E4*>(sm + G::LAT);
    const int gi0 = (cx.I0 - G::H) >> L, gj0 = (cx.J0 - G::H) >> L;  // arithmetic shift: I0-H is a multiple of 2^L
    const int Win = 1 << sp.din;
    E4 v[NLOAD];
    int slot[NLOAD];
#pragma unroll
    for (int q = 0; q < NLOAD; q++) {
      const int idx = tid + q * NT;
      const int li = idx / NIN, lj = idx - li * NIN;
      const int gi = gi0 + li, gj = gj0 + lj;
      slot[q] = -1;
      if (idx < NIN * NIN && (!EDGE || (gi >= 0 && gj >= 0 && gi <= Win && gj <= Win))) {
        slot[q] = G::lat_slot(li * STEP, lj * STEP);
        if (cx.gin) {
          v[q] = cx.gin[(size_t)gi * sp.gin_pitch + (size_t)gj];
        } else {
          const Elem e = initial_elem(gi, gj, Win);  // Terrain.cpp:46-56
          v[q] = e4_of(pos_of(e), e.shade);
        }
      }
    }
    uint32_t* jump = reinterpret_cast<uint32_t*>(sm + G::JUMP);
    for (int c = tid; c < G::NR; c += NT) {
      const Jump j = jump_of((uint64_t)c);
      jump[2 * c] = j.a;
      jump[2 * c + 1] = j.c;
    }
    uint32_t* rows = reinterpret_cast<uint32_t*>(sm + G::ROWS);
    for (int idx = NT - 1 - tid; idx < L * 3 * G::NR; idx += NT) {  // highest threads first: not the lattice loaders
      const int m = idx / (3 * G::NR);
      rows[idx] = row_state_dispatch<0>(sp, cx, m, idx - m * 3 * G::NR);
    }
#pragma unroll
    for (int q = 0; q < NLOAD; q++)
      if (slot[q] >= 0) lat[slot[q]] = v[q];
  }

  template <int M, int WHICH>
  static RN_HD uint32_t draw_of(const float* sm, int row, int col)
  {
    const uint2_t j = reinterpret_cast<const uint2_t*>(sm + G::JUMP)[col];
    const uint32_t* rows = reinterpret_cast<const uint32_t*>(sm + G::ROWS) + (M * 3 + WHICH) * G::NR;
    return lcg_draw_from_state(j.x * rows[row] + j.y);
  }

  // ------------------------------------------------------------------ phase S ----
  template <int M>
  static RN_HD void phase_S(const StageParams& sp, const TileCtx& cx, float* sm, int tid)
  {
    constexpr int h = 1 << (L - 1 - M);
    using R = LevelRange<h, T>;
    const View<L, T, h> v(sm);
    const StageLevel& lv = sp.lev[M];
    const int A0 = cx.I0 >> (L - M), B0 = cx.J0 >> (L - M);
    const unsigned n = (unsigned)lv.n;
    for_cells<R::NS, NT>(tid, [&](int ra, int rb) {
      const int a = R::S_lo + ra, b = R::S_lo + rb;
      const bool ex = !EDGE || ((unsigned)(A0 + a) < n && (unsigned)(B0 + b) < n);
      float loss = 0.0f;
      if (ex) {
        // averageSquare (Terrain.cpp:216-222) + generateSquare (:224-258)
        const E4 k00 = v.K(a, b), k10 = v.K(a + 1, b), k01 = v.K(a, b + 1), k11 = v.K(a + 1, b + 1);
        const V3 avg = vmul(vadd(vadd(vadd(pos4(k00), pos4(k10)), pos4(k01)), pos4(k11)), 0.25f);
        const float sh = (k00.w + k10.w + k01.w + k11.w) * 0.25f;
        const uint32_t draw = draw_of<M, 0>(sm, ra, rb);
        const Displaced dp = displace<true>(sp.law, avg, vsub(pos4(k11), pos4(k00)), vsub(pos4(k10), pos4(k01)), lv.blahS, draw);
        v.C(a, b) = e4_of(dp.pos, sh + shade_gain(sp.law, dp.offset));
        loss = shade_loss(sp.law, dp.offset);
      }
      v.TS(a, b) = loss;
    });
  }

  // ------------------------------------------------------------------ phase D ----
  template <int M>
  static RN_HD void phase_D(const StageParams& sp, const TileCtx& cx, float* sm, int tid)
  {
    constexpr int h = 1 << (L - 1 - M);
    using R = LevelRange<h, T>;
    const View<L, T, h> v(sm);
    const StageLevel& lv = sp.lev[M];
    const int A0 = cx.I0 >> (L - M), B0 = cx.J0 >> (L - M);
    const int n = lv.n;
    for_cells<R::ND, NT>(tid, [&](int ra, int rb) {
      const int a = R::D_lo + ra, b = R::D_lo + rb;
      const int Ag = A0 + a, Bg = B0 + b;
      bool v1 = a >= R::MA_lo && a <= R::MA_hi && b >= R::MB_lo && b <= R::MB_hi;
      bool v2 = a >= R::MB_lo && a <= R::MB_hi && b >= R::MA_lo && b <= R::MA_hi;
      if (EDGE) {
        v1 = v1 && Ag >= 0 && Ag <= n && Bg >= 0 && Bg < n;
        v2 = v2 && Ag >= 0 && Ag < n && Bg >= 0 && Bg <= n;
      }
      float t1 = 0.0f, t2 = 0.0f;
      if (v1 || v2) {
        // corner shades as generateSquare left them: minus the scatter of the surrounding centres,
        // in call order (x-1,y-1), (x-1,y), (x,y-1), (x,y)   (Terrain.cpp:252-255)
        const float ts_mm = v.TS(a - 1, b - 1), ts_m0 = v.TS(a - 1, b), ts_0m = v.TS(a, b - 1), ts_00 = v.TS(a, b);
        E4 k00 = v.K(a, b);
        k00.w = (((k00.w - ts_mm) - ts_m0) - ts_0m) - ts_00;
        if (v1) {  // loop 1: midpoint (a*s, b*s+h); x neighbours are centres, y neighbours corners
          E4 k01 = v.K(a, b + 1);
          k01.w = (((k01.w - ts_m0) - v.TS(a - 1, b + 1)) - ts_00) - v.TS(a, b + 1);
          const uint32_t draw = draw_of<M, 1>(sm, ra, rb);
          E4 m;
          if (EDGE && (Ag == 0 || Ag == n)) {
            const E4 dummy = k00;
            m = mid_border(sp.law, Ag == 0 ? dummy : v.C(a - 1, b), Ag == n ? dummy : v.C(a, b), k00, k01, Ag == 0, Ag == n, false, false,
                           lv.blahDh, draw, &t1);
          } else {
            m = mid_interior(sp.law, v.C(a - 1, b), v.C(a, b), k00, k01, lv.blahD, draw, &t1);
          }
          v.M1(a, b) = m;
        }
        if (v2) {  // loop 2: midpoint (a*s+h, b*s); x neighbours are corners, y neighbours centres
          E4 k10 = v.K(a + 1, b);
          k10.w = (((k10.w - ts_0m) - ts_00) - v.TS(a + 1, b - 1)) - v.TS(a + 1, b);
          const uint32_t draw = draw_of<M, 2>(sm, ra, rb);
          E4 m;
          if (EDGE && (Bg == 0 || Bg == n)) {
            const E4 dummy = k00;
            m = mid_border(sp.law, k00, k10, Bg == 0 ? dummy : v.C(a, b - 1), Bg == n ? dummy : v.C(a, b), false, false, Bg == 0, Bg == n,
                           lv.blahD, draw, &t2);
          } else {
            m = mid_interior(sp.law, k00, k10, v.C(a, b - 1), v.C(a, b), lv.blahD, draw, &t2);
          }
          v.M2(a, b) = m;
        }
      }
      v.TD1(a, b) = t1;
      v.TD2(a, b) = t2;
    });
  }

  // ------------------------------------------------------------------ phase G ----
  template <int M>
  static RN_HD void phase_G(const StageParams& sp, const TileCtx& cx, float* sm, int tid)
  {
    constexpr int h = 1 << (L - 1 - M);
    using R = LevelRange<h, T>;
    const View<L, T, h> v(sm);
    const StageLevel& lv = sp.lev[M];
    const int A0 = cx.I0 >> (L - M), B0 = cx.J0 >> (L - M);
    const int n = lv.n;
    for_cells<R::NG, NT>(tid, [&](int ra, int rb) {
      const int a = R::G_lo + ra, b = R::G_lo + rb;
      const int Ag = A0 + a, Bg = B0 + b;
      bool gk = a >= R::GK_lo && a <= R::GK_hi && b >= R::GK_lo && b <= R::GK_hi;
      bool gc = a >= R::GC_lo && a <= R::GC_hi && b >= R::GC_lo && b <= R::GC_hi;
      if (EDGE) {
        gk = gk && Ag >= 0 && Ag <= n && Bg >= 0 && Bg <= n;
        gc = gc && Ag >= 0 && Ag < n && Bg >= 0 && Bg < n;
      }
      if (gk) {
        // corner: generateSquare scatter of 4 squares, then generateDiamond loop 1 (same row), loop 2 (same column)
        float sh = v.K(a, b).w;
        sh -= v.TS(a - 1, b - 1);
        sh -= v.TS(a - 1, b);
        sh -= v.TS(a, b - 1);
        sh -= v.TS(a, b);
        sh -= v.TD1(a, b - 1);
        sh -= v.TD1(a, b);
        sh -= v.TD2(a - 1, b);
        sh -= v.TD2(a, b);
        v.K(a, b).w = sh;
      }
      if (gc) {
        // square centre: loop-1 midpoints are its x neighbours, loop-2 midpoints its y neighbours
        float sh = v.C(a, b).w;
        sh -= v.TD1(a, b);
        sh -= v.TD1(a + 1, b);
        sh -= v.TD2(a, b);
        sh -= v.TD2(a, b + 1);
        v.C(a, b).w = sh;
      }
    });
  }

  // ---------------------------------------------------------------- copy-out ----
  // E4 slot of output sample (r, c) of the tile: the four classes of the last level.
  static RN_HD int out_slot(int r, int c)
  {
    const int A = r >> 1, B = c >> 1;
    const int ci = (A + 1) * G::PC + (B + 1);
    if (r & 1) return (c & 1) ? G::CARR4 + ci : G::M2ARR4 + ci;
    return (c & 1) ? G::M1ARR4 + ci : G::LAT4 + G::lat_slot(A + G::HL, B + G::HL);
  }

  static RN_HD void store_point(const StageParams& sp, const TileCtx& cx, const float* sm, int r, int c, int Wout, float inv_w)
  {
    const int gi = cx.I0 + r, gj = cx.J0 + c;
    const int wi = gi - sp.win_i0, wj = gj - sp.win_j0;
    if ((unsigned)wi >= (unsigned)sp.win_ni || (unsigned)wj >= (unsigned)sp.win_nj) return;
    E4 e = reinterpret_cast<const E4*>(sm)[out_slot(r, c)];
    if (sp.volcano) {
      const float fx = (float)(gi - Wout / 2), fy = (float)(gj - Wout / 2);
      e.y = e.y + volcano_height_r2<false>(fx * fx + fy * fy, inv_w, sp.volcano_gain, 0.0f);
    }
    cx.gout[(size_t)wi * sp.gout_pitch + (size_t)wj] = e;
  }

  static RN_HD void phase_out(const StageParams& sp, const TileCtx& cx, float* sm, int tid)
  {
    static_assert(NT % T == 0, "a thread keeps its column in the copy-out");
    const int Wout = 1 << sp.dout;
    const float inv_w = 1.0f / (float)Wout;
    const bool close_i = cx.I0 + T >= Wout, close_j = cx.J0 + T >= Wout;  // the last tile also owns row / column W
    const int wi0 = cx.I0 - sp.win_i0, wj0 = cx.J0 - sp.win_j0;
    const bool inside = wi0 >= 0 && wj0 >= 0 && wi0 + T <= sp.win_ni && wj0 + T <= sp.win_nj;
    if (!EDGE && inside) {
      // Fast path: the whole T x T tile is stored.  A thread owns one column c and the rows
      // r0, r0+RS, ...; RS is even, so its samples all come from the same class array and both
      // the shared-memory slot and the global pointer advance by constants.  Within a warp (32
      // consecutive columns of one row) the lane -> column map is permuted so that each quarter
      // warp reads 8 consecutive slots of ONE array (conflict-free LDS.128) while the warp as a
      // whole still stores 512 contiguous bytes.
      constexpr int RS = NT / T;
      static_assert(RS % 2 == 0 && T % 32 == 0, "row parity must stay fixed per thread; whole warps per row");
      const int lane = tid & 31, r0 = tid / T;
      const int q = lane >> 3, k = lane & 7;
      int c32;
      if (r0 & 1) c32 = (q & 1) * 16 + 2 * k + (q >> 1);                 // M2 runs (even columns), then C runs (odd columns)
      else c32 = (q < 2) ? 4 * k + 2 * q : (q & 1) * 16 + 2 * k + 1;     // K: columns 4k / 4k+2 (one LAT half each); M1 runs
      const int c = (tid % T) - lane + c32;
      const E4* src = reinterpret_cast<const E4*>(sm) + out_slot(r0, c);
      const int sstep = (RS / 2) * ((r0 & 1) || (c & 1) ? G::PC : G::PH);
      E4* dst = cx.gout + (size_t)(wi0 + r0) * sp.gout_pitch + (size_t)(wj0 + c);
      const size_t dstep = (size_t)RS * sp.gout_pitch;
      if (!sp.volcano) {
#pragma unroll 4
        for (int kk = 0; kk < T / RS; kk++) {
          *dst = *src;
          src += sstep;
          dst += dstep;
        }
        return;
      }
      const float fy = (float)(cx.J0 + c - Wout / 2);
      const float fy2 = fy * fy;
      const float gain_w = sp.volcano_gain * (float)Wout;
      int xi = cx.I0 + r0 - Wout / 2;
      // crater test for the whole tile: index distance of its nearest sample from the volcano
      const int di = imax_(imax_(cx.I0 - Wout / 2, Wout / 2 - (cx.I0 + T - 1)), 0);
      const int dj = imax_(imax_(cx.J0 - Wout / 2, Wout / 2 - (cx.J0 + T - 1)), 0);
      const float dmin = sqrtf((float)di * (float)di + (float)dj * (float)dj) * inv_w;
      if (dmin > 0.0205f) {  // every sample has d >= minRad with margin for rounding
#pragma unroll 4
        for (int kk = 0; kk < T / RS; kk++) {
          E4 e = *src;
          const float fx = (float)xi;
          e.y = e.y + volcano_height_r2<true>(fx * fx + fy2, inv_w, sp.volcano_gain, gain_w);
          *dst = e;
          src += sstep;
          dst += dstep;
          xi += RS;
        }
      } else {
        for (int kk = 0; kk < T / RS; kk++) {
          E4 e = *src;
          const float fx = (float)xi;
          e.y = e.y + volcano_height_r2<false>(fx * fx + fy2, inv_w, sp.volcano_gain, gain_w);
          *dst = e;
          src += sstep;
          dst += dstep;
          xi += RS;
        }
      }
      return;
    }
    const int nr = imin_(T, Wout - cx.I0) + (close_i ? 1 : 0);  // rows owned (Wout < T: a single short tile)
    const int nc = imin_(T, Wout - cx.J0) + (close_j ? 1 : 0);
    for (int idx = tid; idx < T * T; idx += NT) {
      const int r = idx / T, c = idx - r * T;
      if (r < nr && c < nc) store_point(sp, cx, sm, r, c, Wout, inv_w);
    }
    if (close_i && nr == T + 1)
      for (int c = tid; c < nc; c += NT) store_point(sp, cx, sm, T, c, Wout, inv_w);
    if (close_j && nc == T + 1)
      for (int r = tid; r < imin_(nr, T); r += NT) store_point(sp, cx, sm, r, T, Wout, inv_w);
  }

  // ---------------------------------------------------------------- dispatch ----
  // phase 0 = init, then per level m: 1+3m = S, 2+3m = D, 3+3m = G, last = copy-out.
  template <int P>
  static RN_HD void phase(const StageParams& sp, const TileCtx& cx, float* sm, int tid)
  {
    if constexpr (P == 0) phase_init(sp, cx, sm, tid);
    else if constexpr (P == NPHASES - 1) phase_out(sp, cx, sm, tid);
    else if constexpr ((P - 1) % 3 == 0) phase_S<(P - 1) / 3>(sp, cx, sm, tid);
    else if constexpr ((P - 1) % 3 == 1) phase_D<(P - 1) / 3>(sp, cx, sm, tid);
    else phase_G<(P - 1) / 3>(sp, cx, sm, tid);
  }
};

}  // namespace rn
