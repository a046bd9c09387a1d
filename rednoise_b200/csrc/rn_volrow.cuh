// rn_volrow.cuh -- the warp-row tile routine of the 3D density-volume EXTENSION: the kernel of every level with n >= 64 cells per
// axis (rn_volume.cu), i.e. of > 99.9 % of the samples of a 513^3 or 2049^3 volume.  Same specification as rn_volume.cuh
// (the CPU specification named in DESIGN.md 5.8, bit for bit), re-cut for the instruction-issue bound the thread-per-cell tile ran into (106 thread
// instructions per sample, 82 % issue: profiles/r02h_vol_level_ncu.txt):
//   * a CTA owns TI x TJ x 64 input cells.  A WARP takes one row of 64 cells along k at a time, a lane two adjacent cells
//     (c = 2l, 2l+1), so the sums, the random() arithmetic and the amplitude law of the two samples are ONE packed-FP32
//     instruction stream (add/mul/fma.rn.f32x2, each half rounded exactly like the scalar form), and what is constant along
//     a row -- row addresses, the i and j terms of the lattice hash, the mean's divisor -- is warp-uniform: no per-thread
//     index arithmetic at all;
//   * the shared-memory image keeps, for every fine (i, j) row of the tile, the even-k samples E(c) (fine k = 2c) and the
//     odd-k samples O(c) (fine k = 2c+1) in two halves: the pair a lane needs is one aligned 64-bit load (E(2l), E(2l+1)) or two
//     32-bit loads (E(2l+1), E(2l+2)), all conflict-free, and the copy-out reads full rows with unit stride;
//   * the dependency cone (edge <- face <- cube <- corner) is recomputed per tile as before: one halo cell on the low side
//     of i and j is just more rows; the halo cell along k (column c = -1: one sample per row) and the closing plane along k
//     (column c = 64 of the volume's last tile) are "column jobs", one thread per row with the generic form of the rule;
//   * the row jobs of a tile (which fine row, where it lives in the image, which entries of the hash tables) do not depend
//     on the tile: they are a table in global memory built at compile time, read by a whole warp at a time;
//   * a neighbour outside the volume reads as +0.0f (x + 0.0f == x for every x the path can produce: no sample is ever
//     -0.0f) and only the divisor changes.  The faces k = 0 and k = W -- one tile in n/64 along k -- are
//     handled by every tile at no cost (a per-lane divisor, a zeroed halo column, the closing column as a column job);
//     tiles on an i or j face of the volume (EDGE, 2 in n/4 per axis) run the same packed code on a zero-filled image
//     and derive the divisor of a row from its position.
// Compiles as plain C++ for tests/host_replay (phases run for every warp and lane in turn).
#pragma once
#include "rn_volume.cuh"

namespace rn {


template <int TI, int TJ>
struct RowGeo {
  static constexpr int TK = kRowTK;
  static constexpr int NI = 2 * TI + 3, NJ = 2 * TJ + 3;  // fine rows -2 .. 2T
  static constexpr int EOFF = 2;    // E(c) at EOFF + c, c = -1 .. 64
  static constexpr int OOFF = 82;   // O(c) at OOFF + c, c = -1 .. 63; OOFF = EOFF + 16 (mod 32): the copy-out's even lanes (E) and odd
                                    // lanes (O) fall on disjoint banks
  static constexpr int PK = 146;    // row pitch in floats (even: 64-bit accesses stay aligned)
  static constexpr int ZERO = NI * NJ * PK;  // one more row of zeros (EDGE: rows beyond the image = beyond the volume)
  static constexpr int WORDS = ZERO + PK;
  static constexpr size_t BYTES = (size_t)WORDS * sizeof(float);
  static constexpr RN_HD int row(int fi, int fj) { return ((fi + 2) * NJ + (fj + 2)) * PK; }
};

RN_HD uint32_t vol_mix_draw(uint32_t h)  // the draw of vol_finish: its last xor-shift cannot reach the top 15 bits
{
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  return h >> 17;
}

// random_from_draw for two draws without the doubling: every intermediate is exactly half of the scalar form's (scaling by
// two is exact), and fma(q, 2, -1) rounds 2q - 1 once, like -1 + (2q).  Exhaustively equal: tests/test_volume3d.py.
RN_HD F2 random_from_draw2h(uint32_t r0, uint32_t r1)
{
  const F2 f = f2((float)(int)r0, (float)(int)r1);
  const F2 rc = f2s(1.0f / 32767.0f);
  F2 q = f2mul(f, rc);
  const F2 e = f2fma(f2neg(q), f2s(32767.0f), f);
  q = f2fma(e, rc, q);
  return f2fma(q, f2s(2.0f), f2s(-1.0f));
}

RN_HD F2 ld2(const float* p)  // 8-byte aligned pair
{
#if defined(__CUDA_ARCH__)
  F2 r;
  r.v = *reinterpret_cast<const unsigned long long*>(p);
  return r;
#else
  return f2(p[0], p[1]);
#endif
}
RN_HD F2 ldu(const float* p) { return f2(p[0], p[1]); }  // any pair
RN_HD void st2(float* p, F2 v)
{
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<unsigned long long*>(p) = v.v;
#else
  p[0] = v.a;
  p[1] = v.b;
#endif
}
RN_HD void vol_copy4(float* dst_smem, const float* src_global)
{
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src_global) : "memory");
#else
  *dst_smem = *src_global;
#endif
}
RN_HD void vol_copy_wait()
{
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// ---- the static job tables ----
// One row job = one fine (i, j) row of one sample class for the 64 cells of the tile along k.  The table lives in constant
// memory (rn_volume.cu): a warp reads its job with a warp-uniform index, i.e. through the constant cache.
enum { J_CUBE = 0, J_FK, J_FI, J_FJ, J_EI, J_EJ, J_EK };
struct RowJob {
  int where;  // G::row(2a, 2b), the image row of the cell row's low corner, | word offset of the FH entry of the job's own fine row << 16
  int ab;     // (a + 1) | (b + 1) << 8 | class << 16
};
template <int TI, int TJ, bool EDGE>
struct RowJobs {
  static constexpr int X = EDGE ? 1 : 0;
  static constexpr int MAX = 7 * (TI + 1) * (TJ + 1);
  RowJob j[MAX];
  int c0[8];  // the jobs of class c: [c0[c], c0[c + 1]); cubes first, then the faces, then the edges
};
template <int TI, int TJ, bool EDGE>
constexpr RowJobs<TI, TJ, EDGE> make_row_jobs()
{
  using G = RowGeo<TI, TJ>;
  constexpr int X = EDGE ? 1 : 0;
  RowJobs<TI, TJ, EDGE> t{};
  int n = 0;
  // fine-row parity (pi, pj) and cell-row ranges [a0, a1) x [b0, b1) of the seven classes
  const int pi[7] = {1, 1, 0, 1, 1, 0, 0}, pj[7] = {1, 1, 1, 0, 0, 1, 0};
  const int a0[7] = {-1, -1, 0, -1, 0, 0, 0}, a1[7] = {TI, TI, TI + X, TI, TI, TI + X, TI + X};
  const int b0[7] = {-1, -1, -1, 0, 0, 0, 0}, b1[7] = {TJ, TJ, TJ, TJ + X, TJ + X, TJ, TJ + X};
  for (int cls = 0; cls < 7; cls++) {
    t.c0[cls] = n;
    for (int a = a0[cls]; a < a1[cls]; a++)
      for (int b = b0[cls]; b < b1[cls]; b++) {
        if (!EDGE && cls == J_FK && a < 0 && b < 0) continue;  // no edge of the tile reads that face row
        t.j[n].where = G::row(2 * a, 2 * b) | ((G::WORDS + (2 * a + pi[cls] + 2) * G::NJ + (2 * b + pj[cls] + 2)) << 16);
        t.j[n].ab = (a + 1) | ((b + 1) << 8) | (cls << 16);
        n++;
      }
  }
  t.c0[7] = n;
  return t;
}

template <int TI, int TJ, bool EDGE>  // EDGE: the tile touches an i or j face of the volume
struct VolRows {
  using G = RowGeo<TI, TJ>;
  using Jobs = RowJobs<TI, TJ, EDGE>;
  static constexpr int X = EDGE ? 1 : 0;  // closing rows (fine index 2T) may be owned
  static constexpr int HASH = G::WORDS;   // FH[NI][NJ]: xs16(seed ^ i term) ^ xs16(j term) of the lattice hash of every fine row of the image
  static constexpr size_t BYTES = (size_t)(G::WORDS + G::NI * G::NJ) * sizeof(float);
  static_assert(G::WORDS + G::NI * G::NJ < 65536, "RowJob::where packs two 16-bit word offsets");

  struct Tile {
    int A0, B0, C0, n;
    bool close_i, close_j;   // EDGE: the closing plane i = W / j = W belongs to the tile
    bool k_first, k_last;    // the tile starts on the face k = 0 (no column c = -1) / owns the closing plane k = W (column c = 64)
  };
  static RN_HD Tile tile(const VolLevel& L, int A0, int B0, int C0)
  {
    Tile t;
    t.A0 = A0; t.B0 = B0; t.C0 = C0; t.n = L.n;
    t.close_i = EDGE && A0 + TI >= L.n;
    t.close_j = EDGE && B0 + TJ >= L.n;
    t.k_first = C0 == 0;
    t.k_last = C0 + G::TK >= L.n;
    return t;
  }
  static RN_HD float rcp_of(int cnt) { return cnt == 6 ? 1.0f / 6.0f : (cnt == 5 ? 1.0f / 5.0f : (cnt == 4 ? 1.0f / 4.0f : (cnt == 3 ? 1.0f / 3.0f : 1.0f / 2.0f))); }
  static RN_HD float at(const float* sm, int fi, int fj, int fk)
  {
    const float* r = sm + G::row(fi, fj);
    return (fk & 1) ? r[G::OOFF + ((fk - 1) >> 1)] : r[G::EOFF + (fk >> 1)];
  }
  static RN_HD float* slot(float* sm, int fi, int fj, int fk)
  {
    float* r = sm + G::row(fi, fj);
    return (fk & 1) ? r + G::OOFF + ((fk - 1) >> 1) : r + G::EOFF + (fk >> 1);
  }

  // ---- the generic form of the rule for ONE sample at fine tile coordinates (fi, fj, fk) of parities (PI, PJ, PK) (column
  //      jobs): neighbours in the specification's order, existence tested on the global coordinates.  KPOS says where along k the
  //      caller put the sample: 1 = the halo column fk = -1 of a tile that does not start on the face k = 0 (both neighbours along
  //      k exist), 2 = the closing plane fk = 2 TK = the face k = W (no neighbour at fk + 1); in a lean tile every neighbour along i
  //      and j exists, so there all the tests -- and with them the divisor -- are compile-time constants.
  template <int PI, int PJ, int PK_, int KPOS>
  static RN_HD void generic_sample(const VolLevel& L, const Tile& t, float* sm, int fi, int fj, int fk)
  {
    const int gi = 2 * t.A0 + fi, gj = 2 * t.B0 + fj, gk = 2 * t.C0 + fk, N2 = 2 * t.n;
    if (EDGE && (gi < 0 || gi > N2 || gj < 0 || gj > N2 || gk < 0 || gk > N2)) return;
    auto exists = [&](int axis, int g, int sgn) -> bool {  // the neighbour at g + sgn along `axis`
      if (!EDGE && axis < 2) return true;
      if (!EDGE && KPOS == 1) return true;
      if (!EDGE && KPOS == 2) return sgn < 0;
      return !(g + sgn < 0 || g + sgn > N2);
    };
    constexpr int odd = PI + PJ + PK_;
    float sum = 0.0f;
    int cnt = 0;
    if constexpr (odd == 3) {
      for (int q = 0; q < 8; q++) {
        const float v = at(sm, fi + ((q & 4) ? 1 : -1), fj + ((q & 2) ? 1 : -1), fk + ((q & 1) ? 1 : -1));
        sum = q ? sum + v : v;
      }
      cnt = 8;
    } else if constexpr (odd == 2) {
      constexpr int a = !PI ? 0 : (!PJ ? 1 : 2);  // the face's normal; the in-face axes ascending, the later one fastest
      for (int q = 0; q < 4; q++) {
        const int s0 = (q & 2) ? 1 : -1, s1 = (q & 1) ? 1 : -1;
        const float v = a == 0 ? at(sm, fi, fj + s0, fk + s1) : (a == 1 ? at(sm, fi + s0, fj, fk + s1) : at(sm, fi + s0, fj + s1, fk));
        sum = q ? sum + v : v;
      }
      cnt = 4;
      const int ga = a == 0 ? gi : (a == 1 ? gj : gk);
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        if (!exists(a, ga, sgn)) continue;
        sum += a == 0 ? at(sm, fi + sgn, fj, fk) : (a == 1 ? at(sm, fi, fj + sgn, fk) : at(sm, fi, fj, fk + sgn));
        cnt++;
      }
    } else {  // an edge along axis a: its two ends, then the faces along the other two axes (ascending), - before +
      constexpr int a = PI ? 0 : (PJ ? 1 : 2);
      sum = a == 0 ? at(sm, fi - 1, fj, fk) : (a == 1 ? at(sm, fi, fj - 1, fk) : at(sm, fi, fj, fk - 1));
      sum += a == 0 ? at(sm, fi + 1, fj, fk) : (a == 1 ? at(sm, fi, fj + 1, fk) : at(sm, fi, fj, fk + 1));
      cnt = 2;
      for (int b = 0; b < 3; b++) {
        if (b == a) continue;
        const int gb = b == 0 ? gi : (b == 1 ? gj : gk);
        for (int sgn = -1; sgn <= 1; sgn += 2) {
          if (!exists(b, gb, sgn)) continue;
          sum += b == 0 ? at(sm, fi + sgn, fj, fk) : (b == 1 ? at(sm, fi, fj + sgn, fk) : at(sm, fi, fj, fk + sgn));
          cnt++;
        }
      }
    }
    const float rc = cnt == 8 ? 0.125f : rcp_of(cnt);
    *slot(sm, fi, fj, fk) = vol_finish(L, sum, rc, (uint32_t)gi * L.ci, (uint32_t)gj * L.cj, (uint32_t)gk * L.ck);
  }
  // column job at fine k = -1: the sample, or +0.0f on the face k = 0 (where the packed row code reads it as a neighbour that
  // does not exist)
  template <int PI, int PJ>
  static RN_HD void halo_column(const VolLevel& L, const Tile& t, float* sm, int fi, int fj)
  {
    if (t.k_first) *slot(sm, fi, fj, -1) = 0.0f;
    else generic_sample<PI, PJ, 1, 1>(L, t, sm, fi, fj, -1);
  }

  // ------------------------------------------------------------------ phase 0: the input lattice, the hash tables ----
  // EDGE tiles: zero() for every thread and a barrier first
  static RN_HD void zero(float* sm, int tid, int nt)
  {
    for (int q = tid; q < G::WORDS; q += nt) sm[q] = 0.0f;
  }
  static RN_HD uint32_t xs16(uint32_t x) { return x ^ (x >> 16); }  // the hash's first xor-shift: linear over xor, so it is applied to
                                                                    // the i, j and k terms separately, ahead of time
  static RN_HD void load(const VolLevel& L, const Tile& t, float* sm, int warp, int lane, int nw, int tid)
  {
    uint32_t* hash = reinterpret_cast<uint32_t*>(sm + HASH);
    for (int q = tid; q < G::NI * G::NJ; q += nw * 32) {
      const int fi = q / G::NJ, fj = q - fi * G::NJ;
      hash[q] = xs16(L.seed ^ ((uint32_t)(2 * t.A0 + fi - 2) * L.ci)) ^ xs16((uint32_t)(2 * t.B0 + fj - 2) * L.cj);
    }
    // lattice rows (a, b), a = -1 .. TI, b = -1 .. TJ, round-robin over the warps; columns c = -1 .. 64 of a row: lane - 1, lane + 31
    // and (lanes 0, 1) 63, 64.  Row addresses advance from one base pointer (never dereferenced where the row does not exist).
    const size_t m = (size_t)t.n + 1;
    const float* src0 = L.gin + (((size_t)(t.A0 - L.in_lo) * m + (size_t)t.B0) * m + (size_t)t.C0 + (size_t)lane);
    const int sa = (int)(m * m), sb = (int)m;  // 32-bit row offsets: (TI + 1) * (n + 1)^2 < 2^31 for every level (n <= 1024)
    float* dst0 = sm + G::EOFF + lane;
    const bool c_lo = !(t.k_first && lane == 0);
    for (int r = warp; r < (TI + 2) * (TJ + 2); r += nw) {
      const int a = r / (TJ + 2) - 1, b = r - (a + 1) * (TJ + 2) - 1;
      if (EDGE && (t.A0 + a < 0 || t.B0 + b < 0)) continue;
      const float* src = src0 + (a * sa + b * sb);
      float* dst = dst0 + G::row(2 * a, 2 * b);
      if (c_lo) vol_copy4(dst - 1, src - 1);            // c = lane - 1: -1 .. 30
      vol_copy4(dst + 31, src + 31);                    // c = lane + 31: 31 .. 62
      if (lane < 2) vol_copy4(dst + 63, src + 63);      // c = 63, 64
    }
    vol_copy_wait();
  }

  struct LaneK {
    int l2;                       // 2 * lane
    uint32_t fE0, fE1, fO0, fO1;  // xs16 of the k term of the lattice hash for E(2l), E(2l+1), O(2l), O(2l+1)
    bool k_lo;                    // the lane's first E sample lies on the face k = 0 of the volume
  };
  static RN_HD LaneK lane_k(const VolLevel& L, const Tile& t, int lane)
  {
    LaneK k;
    k.l2 = 2 * lane;
    const uint32_t hE0 = (uint32_t)(2 * (t.C0 + k.l2)) * L.ck, hO0 = hE0 + L.ck, hE1 = hO0 + L.ck, hO1 = hE1 + L.ck;
    k.fE0 = xs16(hE0);
    k.fO0 = xs16(hO0);
    k.fE1 = xs16(hE1);
    k.fO1 = xs16(hO1);
    k.k_lo = t.k_first && lane == 0;
    return k;
  }
  // the rest of the hash after its first xor-shift (vol_mix_draw)
  static RN_HD uint32_t draw_of(uint32_t h)
  {
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    return h >> 17;
  }
  static RN_HD F2 finish(const VolLevel& L, F2 sum, F2 rc, uint32_t f0, uint32_t f1)
  {
    const F2 r = random_from_draw2h(draw_of(f0), draw_of(f1));
    const F2 q = f2mul(f2mul(f2s(L.amplitude), r), f2s(L.scale));
    return f2fma(f2mul(sum, rc), f2s(L.one), q);  // p + q rounded once; written as an FMA so that ptxas cannot contract the products
  }

  // What a row routine needs to know about its job.  cnt = the neighbours of a sample of the row that exist, both neighbours
  // along k counted (lean tiles: all of them).
  struct Row {
    int base;
    uint32_t fh;    // xs16(seed ^ i term ^ j term) of the row's own fine (i, j)
    int cnt;
    bool hi_i, hi_j;  // EDGE: the rows 2a+1 / 2b+1 lie beyond the image (= beyond the volume): zeros
  };
  static RN_HD const float* rp(const float* sm, const Row& r, int di, int dj)
  {
    if (EDGE && ((di == 1 && r.hi_i) || (dj == 1 && r.hi_j))) return sm + G::ZERO;
    return sm + r.base + (di * G::NJ + dj) * G::PK;
  }
  static RN_HD float* rw(float* sm, const Row& r, int di, int dj) { return sm + r.base + (di * G::NJ + dj) * G::PK; }
  // divisor pairs: the lane's first E-class sample loses one neighbour along k on the face k = 0
  static RN_HD F2 rc_e(const Row& r, const LaneK& k)
  {
    if (EDGE) return f2(rcp_of(r.cnt - (k.k_lo ? 1 : 0)), rcp_of(r.cnt));
    return f2(k.k_lo ? 1.0f / 5.0f : 1.0f / 6.0f, 1.0f / 6.0f);
  }
  static RN_HD F2 rc_o(const Row& r) { return EDGE ? f2s(rcp_of(r.cnt)) : f2s(1.0f / 6.0f); }

  // table entry -> Row; false: the row does not exist or nothing reads it (EDGE only)
  static RN_HD bool row_of(const Tile& t, const float* sm, const RowJob& j, Row* r)
  {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(sm);
    r->base = j.where & 0xffff;
    r->fh = w[j.where >> 16];
    r->cnt = 6;
    r->hi_i = false;
    r->hi_j = false;
    if (!EDGE) return true;
    const int a = (j.ab & 0xff) - 1, b = ((j.ab >> 8) & 0xff) - 1, c = j.ab >> 16;
    const int ga = t.A0 + a, gb = t.B0 + b, N2 = 2 * t.n;
    r->hi_i = a == TI;
    r->hi_j = b == TJ;
    if (ga < 0 || gb < 0) return false;
    if ((a == TI && !t.close_i) || (b == TJ && !t.close_j)) return false;
    if (c == J_FK && a < 0 && b < 0) return false;
    const int ti = (ga > 0 ? 1 : 0) + (2 * ga < N2 ? 1 : 0), tj = (gb > 0 ? 1 : 0) + (2 * gb < N2 ? 1 : 0);  // neighbours at fine i -+ 1 / j -+ 1 of an even row
    r->cnt = c == J_FI || c == J_EJ ? 4 + ti : (c == J_FJ || c == J_EI ? 4 + tj : (c == J_EK ? 2 + ti + tj : 6));
    return true;
  }
  // the jobs of class CLS that fall to this warp: the round-robin runs on across the class boundaries (job q goes to warp
  // q mod nw, nw a power of two), so that every phase stays balanced without a per-job switch on the class; two jobs at a time
  template <int CLS>
  static RN_HD void run_jobs(const VolLevel& L, const Tile& t, float* sm, const Jobs& jobs, const LaneK& k, int warp, int nw)
  {
    const int b = jobs.c0[CLS], e = jobs.c0[CLS + 1];
    int q = b + ((warp - b) & (nw - 1));
    for (; q + nw < e; q += 2 * nw) {
      Row r0, r1;
      const bool ok0 = row_of(t, sm, jobs.j[q], &r0), ok1 = row_of(t, sm, jobs.j[q + nw], &r1);
      F2 s0 = f2s(0.0f), s1 = f2s(0.0f);
      if (ok0) s0 = row_sum<CLS>(sm, k, r0);
      if (ok1) s1 = row_sum<CLS>(sm, k, r1);
      if (ok0) row_put<CLS>(L, sm, k, r0, s0);
      if (ok1) row_put<CLS>(L, sm, k, r1, s1);
    }
    if (q < e) {
      Row r;
      if (row_of(t, sm, jobs.j[q], &r)) row_put<CLS>(L, sm, k, r, row_sum<CLS>(sm, k, r));
    }
  }

  // ------------------------------------------------------------------ the seven row routines ----
  // Each in two halves, so that a warp can have the shared-memory loads of TWO jobs in flight before it stores either result
  // (the compiler cannot move a load across a store to the same image): row_sum = the neighbours in the specification's order,
  // row_put = mean + displacement, stored to the row's own slots.
  //   J_CUBE cube centres of cell row (a, b): fine row (2a+1, 2b+1), O(c)      J_FK face centres, normal k: (2a+1, 2b+1), E(c)
  //   J_FI   face centres, normal i: (2a, 2b+1), O(c)                           J_FJ face centres, normal j: (2a+1, 2b), O(c)
  //   J_EI   edge midpoints along i: (2a+1, 2b), E(c)    J_EJ along j: (2a, 2b+1), E(c)    J_EK along k: (2a, 2b), O(c)
  template <int CLS>
  static RN_HD F2 row_sum(const float* sm, const LaneK& k, const Row& r)
  {
    if constexpr (CLS == J_CUBE) {
      const float* r00 = rp(sm, r, 0, 0) + G::EOFF + k.l2;
      const float* r01 = rp(sm, r, 0, 2) + G::EOFF + k.l2;
      const float* r10 = rp(sm, r, 2, 0) + G::EOFF + k.l2;
      const float* r11 = rp(sm, r, 2, 2) + G::EOFF + k.l2;
      F2 s = ld2(r00);
      s = f2add(s, ldu(r00 + 1));
      s = f2add(s, ld2(r01));
      s = f2add(s, ldu(r01 + 1));
      s = f2add(s, ld2(r10));
      s = f2add(s, ldu(r10 + 1));
      s = f2add(s, ld2(r11));
      return f2add(s, ldu(r11 + 1));
    } else if constexpr (CLS == J_FK) {
      const float* own = rp(sm, r, 1, 1);
      F2 s = ld2(rp(sm, r, 0, 0) + G::EOFF + k.l2);
      s = f2add(s, ld2(rp(sm, r, 0, 2) + G::EOFF + k.l2));
      s = f2add(s, ld2(rp(sm, r, 2, 0) + G::EOFF + k.l2));
      s = f2add(s, ld2(rp(sm, r, 2, 2) + G::EOFF + k.l2));
      s = f2add(s, ldu(own + G::OOFF + k.l2 - 1));
      return f2add(s, ld2(own + G::OOFF + k.l2));
    } else if constexpr (CLS == J_FI) {
      const float* r0 = rp(sm, r, 0, 0) + G::EOFF + k.l2;
      const float* r1 = rp(sm, r, 0, 2) + G::EOFF + k.l2;
      F2 s = ld2(r0);
      s = f2add(s, ldu(r0 + 1));
      s = f2add(s, ld2(r1));
      s = f2add(s, ldu(r1 + 1));
      s = f2add(s, ld2(rp(sm, r, -1, 1) + G::OOFF + k.l2));
      return f2add(s, ld2(rp(sm, r, 1, 1) + G::OOFF + k.l2));
    } else if constexpr (CLS == J_FJ) {
      const float* r0 = rp(sm, r, 0, 0) + G::EOFF + k.l2;
      const float* r1 = rp(sm, r, 2, 0) + G::EOFF + k.l2;
      F2 s = ld2(r0);
      s = f2add(s, ldu(r0 + 1));
      s = f2add(s, ld2(r1));
      s = f2add(s, ldu(r1 + 1));
      s = f2add(s, ld2(rp(sm, r, 1, -1) + G::OOFF + k.l2));
      return f2add(s, ld2(rp(sm, r, 1, 1) + G::OOFF + k.l2));
    } else if constexpr (CLS == J_EI) {
      const float* own = rp(sm, r, 1, 0);
      F2 s = ld2(rp(sm, r, 0, 0) + G::EOFF + k.l2);
      s = f2add(s, ld2(rp(sm, r, 2, 0) + G::EOFF + k.l2));
      s = f2add(s, ld2(rp(sm, r, 1, -1) + G::EOFF + k.l2));
      s = f2add(s, ld2(rp(sm, r, 1, 1) + G::EOFF + k.l2));
      s = f2add(s, ldu(own + G::OOFF + k.l2 - 1));
      return f2add(s, ld2(own + G::OOFF + k.l2));
    } else if constexpr (CLS == J_EJ) {
      const float* own = rp(sm, r, 0, 1);
      F2 s = ld2(rp(sm, r, 0, 0) + G::EOFF + k.l2);
      s = f2add(s, ld2(rp(sm, r, 0, 2) + G::EOFF + k.l2));
      s = f2add(s, ld2(rp(sm, r, -1, 1) + G::EOFF + k.l2));
      s = f2add(s, ld2(rp(sm, r, 1, 1) + G::EOFF + k.l2));
      s = f2add(s, ldu(own + G::OOFF + k.l2 - 1));
      return f2add(s, ld2(own + G::OOFF + k.l2));
    } else {
      const float* own = rp(sm, r, 0, 0);
      F2 s = ld2(own + G::EOFF + k.l2);
      s = f2add(s, ldu(own + G::EOFF + k.l2 + 1));
      s = f2add(s, ld2(rp(sm, r, -1, 0) + G::OOFF + k.l2));
      s = f2add(s, ld2(rp(sm, r, 1, 0) + G::OOFF + k.l2));
      s = f2add(s, ld2(rp(sm, r, 0, -1) + G::OOFF + k.l2));
      return f2add(s, ld2(rp(sm, r, 0, 1) + G::OOFF + k.l2));
    }
  }
  template <int CLS>
  static RN_HD void row_put(const VolLevel& L, float* sm, const LaneK& k, const Row& r, F2 s)
  {
    constexpr int di = (CLS == J_CUBE || CLS == J_FK || CLS == J_FJ || CLS == J_EI) ? 1 : 0;  // the row's own fine (i, j) relative to (2a, 2b)
    constexpr int dj = (CLS == J_CUBE || CLS == J_FK || CLS == J_FI || CLS == J_EJ) ? 1 : 0;
    constexpr bool e_class = CLS == J_FK || CLS == J_EI || CLS == J_EJ;                        // even fine k: E(c)
    float* own = rw(sm, r, di, dj);
    if constexpr (e_class) st2(own + G::EOFF + k.l2, finish(L, s, rc_e(r, k), r.fh ^ k.fE0, r.fh ^ k.fE1));
    else st2(own + G::OOFF + k.l2, finish(L, s, CLS == J_CUBE ? f2s(0.125f) : rc_o(r), r.fh ^ k.fO0, r.fh ^ k.fO1));
  }

  // ------------------------------------------------------------------ phases 1 .. 3 ----
  // Row jobs go round-robin over the warps; the column jobs of a phase one per thread (tid).
  static constexpr int NC = (TI + 1) * (TJ + 1);
  static constexpr int NFI = (TI + X) * (TJ + 1), NFJ = (TI + 1) * (TJ + X);
  static constexpr int NEI = TI * (TJ + X), NEJ = (TI + X) * TJ;

  // (the column jobs go to the LAST threads of the block: with 8 warps the first warps carry one row job more)
  static RN_HD void cubes(const VolLevel& L, const Tile& t, float* sm, const Jobs& jobs, int warp, int lane, int nw, int tid)
  {
    const LaneK k = lane_k(L, t, lane);
    run_jobs<J_CUBE>(L, t, sm, jobs, k, warp, nw);
    const int ct = nw * 32 - 1 - tid;
    if (ct < NC) halo_column<1, 1>(L, t, sm, 2 * (ct / (TJ + 1) - 1) + 1, 2 * (ct % (TJ + 1) - 1) + 1);
  }
  static RN_HD void faces(const VolLevel& L, const Tile& t, float* sm, const Jobs& jobs, int warp, int lane, int nw, int tid)
  {
    const LaneK k = lane_k(L, t, lane);
    run_jobs<J_FK>(L, t, sm, jobs, k, warp, nw);
    run_jobs<J_FI>(L, t, sm, jobs, k, warp, nw);
    run_jobs<J_FJ>(L, t, sm, jobs, k, warp, nw);
    // column jobs: O(-1) of the normal-i and normal-j faces, E(64) of the normal-k faces on the closing plane
    // (one class per warp, from the last warp down: a warp runs one branch only)
    static_assert(NFI <= 32 && NFJ <= 32 && NC <= 32, "a column-job class fits one warp");
    const int ct = nw * 32 - 1 - tid, cw = ct >> 5, q = ct & 31;
    if (cw == 0 && q < NFI) {
      const int a = q / (TJ + 1), b = q % (TJ + 1) - 1;
      if (!(EDGE && a == TI && !t.close_i)) halo_column<0, 1>(L, t, sm, 2 * a, 2 * b + 1);
    } else if (cw == 1 && q < NFJ) {
      const int a = q / (TJ + X) - 1, b = q % (TJ + X);
      if (!(EDGE && b == TJ && !t.close_j)) halo_column<1, 0>(L, t, sm, 2 * a + 1, 2 * b);
    } else if (cw == 2 && q < NC) {
      if (t.k_last) generic_sample<1, 1, 0, 2>(L, t, sm, 2 * (q / (TJ + 1) - 1) + 1, 2 * (q % (TJ + 1) - 1) + 1, 2 * G::TK);
    }
  }
  static RN_HD void edges(const VolLevel& L, const Tile& t, float* sm, const Jobs& jobs, int warp, int lane, int nw, int tid)
  {
    const LaneK k = lane_k(L, t, lane);
    run_jobs<J_EI>(L, t, sm, jobs, k, warp, nw);
    run_jobs<J_EJ>(L, t, sm, jobs, k, warp, nw);
    run_jobs<J_EK>(L, t, sm, jobs, k, warp, nw);
    if (t.k_last) {  // E(64) of the edges along i and along j on the closing plane
      static_assert(NEI <= 32 && NEJ <= 32, "a column-job class fits one warp");
      const int ct = nw * 32 - 1 - tid, cw = ct >> 5, q = ct & 31;
      if (cw == 0 && q < NEI) {
        const int a = q / (TJ + X), b = q % (TJ + X);
        if (!(EDGE && b == TJ && !t.close_j)) generic_sample<1, 0, 0, 2>(L, t, sm, 2 * a + 1, 2 * b, 2 * G::TK);
      } else if (cw == 1 && q < NEJ) {
        const int a = q / TJ, b = q % TJ;
        if (!(EDGE && a == TI && !t.close_i)) generic_sample<0, 1, 0, 2>(L, t, sm, 2 * a, 2 * b + 1, 2 * G::TK);
      }
    }
  }

  // ------------------------------------------------------------------ phase 4: the finished tile, full rows along k ----
  static RN_HD void store(const VolLevel& L, const Tile& t, const float* sm, int warp, int lane, int nw)
  {
    const size_t M = 2 * (size_t)t.n + 1;
    constexpr int SJ = 2 * TJ + X;
    const int src0 = (lane & 1) ? G::OOFF + (lane >> 1) : G::EOFF + (lane >> 1);
    if (!EDGE && nw % (2 * TJ) == 0) {  // a warp keeps its fine j and walks down i (every (nw / 2 TJ)-th row): running pointers
      const int per = nw / (2 * TJ), fj = warp % (2 * TJ), part = warp / (2 * TJ);
      const int lo = (L.out_lo - 2 * t.A0 > 0 ? L.out_lo - 2 * t.A0 : 0) + part;
      const int hi = L.out_hi - 2 * t.A0 < 2 * TI - 1 ? L.out_hi - 2 * t.A0 : 2 * TI - 1;
      const float* src = sm + G::row(lo, fj) + src0;
      const size_t plane = M * M;
      float* dst = L.gout + ((size_t)(2 * t.A0 + lo - L.out_lo) * plane + (size_t)(2 * t.B0 + fj) * M + (size_t)(2 * t.C0) + (size_t)lane);
      for (int fi = lo; fi <= hi; fi += per) {
#pragma unroll
        for (int q = 0; q < 4; q++) dst[32 * q] = src[16 * q];
        src += per * G::NJ * G::PK;
        dst += per * plane;
      }
      if (t.k_last && lo + per * lane <= hi) {  // the closing column k = W: lane q stores it for the warp's q-th row
        const int fi = lo + per * lane;
        const float* cs = sm + G::row(fi, fj) + G::EOFF + G::TK;
        L.gout[(size_t)(2 * t.A0 + fi - L.out_lo) * plane + (size_t)(2 * t.B0 + fj) * M + (size_t)(2 * t.C0) + (size_t)(2 * G::TK)] = *cs;
      }
      return;
    }
    for (int r = warp; r < (2 * TI + X) * SJ; r += nw) {
      const int fi = r / SJ, fj = r % SJ;
      if (EDGE && ((fi == 2 * TI && !t.close_i) || (fj == 2 * TJ && !t.close_j))) continue;
      const int gi = 2 * t.A0 + fi, gj = 2 * t.B0 + fj;
      if (gi < L.out_lo || gi > L.out_hi) continue;
      const float* src = sm + G::row(fi, fj);
      float* dst = L.gout + ((size_t)(gi - L.out_lo) * M + (size_t)gj) * M + (size_t)(2 * t.C0) + lane;
#pragma unroll
      for (int q = 0; q < 4; q++) dst[32 * q] = src[src0 + 16 * q];
      if (t.k_last && lane == 0) dst[2 * G::TK] = src[G::EOFF + G::TK];
    }
  }
};

}  // namespace rn
