// rn_volume.cuh -- the tile routine of the 3D density-volume EXTENSION (see rn_volume.cu for the design) and its host-side
// planning.  Compiles as plain C++ too: tests/host_replay runs the same tile code with one "thread" per tile (NT = 1, the
// phases then follow one another exactly as the barriers order them on the GPU).
#pragma once
#include "rn_math.cuh"

namespace rn {

#if defined(__CUDA_ARCH__)
#define RN_VOL_TID ((int)threadIdx.x)
#define RN_VOL_SYNC() __syncthreads()
#else
#define RN_VOL_TID 0
#define RN_VOL_SYNC() ((void)0)
#endif

constexpr uint32_t kHashI = 0x9E3779B1u, kHashJ = 0x85EBCA77u, kHashK = 0xC2B2AE3Du;

struct VolLevel {
  const float* gin;    // plane in_lo of the level-l grid (side m = n + 1)
  float* gout;         // plane out_lo of the level-(l+1) grid (side M = 2n + 1)
  int n;               // cells per axis of the input grid
  int in_lo;           // first stored input plane (lattice index along i)
  int out_lo, out_hi;  // output planes to store (fine index along i), inclusive
  int tile_i0;         // tile index along i of blockIdx.z == 0
  uint32_t seed;
  uint32_t ci, cj, ck; // hash multipliers of one fine step of this level: (coordinate << sh) * C == coordinate * (C << sh)
  float amplitude, scale;  // scale = s / W: ((amplitude * r) * s) / W == (amplitude * r) * scale exactly (powers of two)
  float one;               // 1.0f the compiler cannot see (rn_volrow.cuh: a sum of packed products written as an FMA)
};

template <int TI, int TJ, int TK>
struct VolGeo {
  // fine coordinates -2 .. 2T of the tile, stored at +2; along k split by parity
  static constexpr int NI = 2 * TI + 3, NJ = 2 * TJ + 3;
  static constexpr int HK = TK + 2;                       // even half: fine k = 2c (c = -1 .. TK) at c + 1
  static constexpr int PK0 = HK + TK + 1;                 // odd half: fine k = 2c + 1 (c = -1 .. TK-1) at HK + c + 1
  static constexpr int PK = (PK0 % 2) ? PK0 : PK0 + 1;    // odd row pitch
  static constexpr int PJ = NJ * PK;
  static constexpr int WORDS = NI * PJ;
  static constexpr size_t BYTES = (size_t)WORDS * sizeof(float);
  static RN_HD int row(int fi, int fj) { return (fi + 2) * PJ + (fj + 2) * PK; }  // fine (i, j) -> row start
  static RN_HD int E(int c) { return c + 1; }        // fine k = 2c
  static RN_HD int O(int c) { return HK + c + 1; }   // fine k = 2c + 1
};

// value of the new sample at global fine coordinates (I, J, K): mean + amplitude law (Rock.cpp:185) with the lattice hash's draw
RN_HD float vol_finish(const VolLevel& L, float sum, float rc, uint32_t hi, uint32_t hj, uint32_t hk)
{
  uint32_t h = L.seed ^ hi ^ hj ^ hk;
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  const float r = random_from_draw(h >> 17);
  return sum * rc + (L.amplitude * r) * L.scale;
}

template <int TI, int TJ, int TK, int NT, bool EDGE>
RN_HD void vol_tile(const VolLevel& L, float* sm, int A0, int B0, int C0)
{
  using G = VolGeo<TI, TJ, TK>;
  const int tid = RN_VOL_TID;
  const int n = L.n;
  const size_t m = (size_t)n + 1, M = 2 * (size_t)n + 1;
  const bool close_i = EDGE && A0 + TI >= n, close_j = EDGE && B0 + TJ >= n, close_k = EDGE && C0 + TK >= n;
  constexpr float r2 = 1.0f / 2.0f, r3 = 1.0f / 3.0f, r4 = 1.0f / 4.0f, r5 = 1.0f / 5.0f, r6 = 1.0f / 6.0f;
  // ---- phase 0: the input lattice, cells -1 .. T (T + 2 points per axis)
  {
    constexpr int LJ = TJ + 2, LK = TK + 2, NL = (TI + 2) * LJ * LK;
    for (int idx = tid; idx < NL; idx += NT) {
      const int c = idx % LK - 1, b = (idx / LK) % LJ - 1, a = idx / (LK * LJ) - 1;
      const int ga = A0 + a, gb = B0 + b, gc = C0 + c;
      if (!EDGE || (ga >= 0 && ga <= n && gb >= 0 && gb <= n && gc >= 0 && gc <= n))
        sm[G::row(2 * a, 2 * b) + G::E(c)] = L.gin[((size_t)(ga - L.in_lo) * m + (size_t)gb) * m + (size_t)gc];
    }
  }
  RN_VOL_SYNC();
  // ---- phase 1: cube centres of cells -1 .. T-1
  {
    constexpr int CJ = TJ + 1, CK = TK + 1, NC = (TI + 1) * CJ * CK;
    for (int idx = tid; idx < NC; idx += NT) {
      const int c = idx % CK - 1, b = (idx / CK) % CJ - 1, a = idx / (CK * CJ) - 1;
      const int ga = A0 + a, gb = B0 + b, gc = C0 + c;
      if (EDGE && !(ga >= 0 && ga < n && gb >= 0 && gb < n && gc >= 0 && gc < n)) continue;
      const float* p = sm + G::row(2 * a, 2 * b);
      float sum = p[G::E(c)];
      sum += p[G::E(c + 1)];
      sum += p[2 * G::PK + G::E(c)];
      sum += p[2 * G::PK + G::E(c + 1)];
      sum += p[2 * G::PJ + G::E(c)];
      sum += p[2 * G::PJ + G::E(c + 1)];
      sum += p[2 * G::PJ + 2 * G::PK + G::E(c)];
      sum += p[2 * G::PJ + 2 * G::PK + G::E(c + 1)];
      sm[G::row(2 * a + 1, 2 * b + 1) + G::O(c)] =
          vol_finish(L, sum, 0.125f, (uint32_t)(2 * ga + 1) * L.ci, (uint32_t)(2 * gb + 1) * L.cj, (uint32_t)(2 * gc + 1) * L.ck);
    }
  }
  RN_VOL_SYNC();
  // ---- phase 2: the three low faces of cells -1 .. T-1 (EDGE: .. T, the closing planes); a face needs the cube centres on
  //      both sides, so its index along its normal starts at 0
  {
    constexpr int X = EDGE ? 2 : 1;
    constexpr int FJ = TJ + X, FK = TK + X, NF = (TI + X) * FJ * FK;
    for (int idx = tid; idx < NF; idx += NT) {
      const int c = idx % FK - 1, b = (idx / FK) % FJ - 1, a = idx / (FK * FJ) - 1;
      const int ga = A0 + a, gb = B0 + b, gc = C0 + c;
      const uint32_t hi0 = (uint32_t)(2 * ga) * L.ci, hj0 = (uint32_t)(2 * gb) * L.cj, hk0 = (uint32_t)(2 * gc) * L.ck;
      const bool in_a = a < TI && (!EDGE || (ga >= 0 && ga < n)), in_b = b < TJ && (!EDGE || (gb >= 0 && gb < n)),
                 in_c = c < TK && (!EDGE || (gc >= 0 && gc < n));  // the cell's odd coordinate along the axis exists and is computed here
      const bool on_a = a >= 0 && (a < TI || close_i) && (!EDGE || (ga >= 0 && ga <= n));
      const bool on_b = b >= 0 && (b < TJ || close_j) && (!EDGE || (gb >= 0 && gb <= n));
      const bool on_c = c >= 0 && (c < TK || close_k) && (!EDGE || (gc >= 0 && gc <= n));
      const float* p = sm + G::row(2 * a, 2 * b);
      if (on_a && in_b && in_c) {  // normal i: (2a, 2b+1, 2c+1); corners (j-,k-), (j-,k+), (j+,k-), (j+,k+); cubes i-, i+
        float sum = p[G::E(c)];
        sum += p[G::E(c + 1)];
        sum += p[2 * G::PK + G::E(c)];
        sum += p[2 * G::PK + G::E(c + 1)];
        float rc = r6;
        if (EDGE) {
          int cnt = 4;
          if (ga >= 1) { sum += p[-G::PJ + G::PK + G::O(c)]; cnt++; }
          if (ga <= n - 1) { sum += p[G::PJ + G::PK + G::O(c)]; cnt++; }
          rc = cnt == 6 ? r6 : r5;
        } else {
          sum += p[-G::PJ + G::PK + G::O(c)];
          sum += p[G::PJ + G::PK + G::O(c)];
        }
        sm[G::row(2 * a, 2 * b + 1) + G::O(c)] = vol_finish(L, sum, rc, hi0, hj0 + L.cj, hk0 + L.ck);
      }
      if (in_a && on_b && in_c) {  // normal j: (2a+1, 2b, 2c+1); corners (i-,k-), (i-,k+), (i+,k-), (i+,k+); cubes j-, j+
        float sum = p[G::E(c)];
        sum += p[G::E(c + 1)];
        sum += p[2 * G::PJ + G::E(c)];
        sum += p[2 * G::PJ + G::E(c + 1)];
        float rc = r6;
        if (EDGE) {
          int cnt = 4;
          if (gb >= 1) { sum += p[G::PJ - G::PK + G::O(c)]; cnt++; }
          if (gb <= n - 1) { sum += p[G::PJ + G::PK + G::O(c)]; cnt++; }
          rc = cnt == 6 ? r6 : r5;
        } else {
          sum += p[G::PJ - G::PK + G::O(c)];
          sum += p[G::PJ + G::PK + G::O(c)];
        }
        sm[G::row(2 * a + 1, 2 * b) + G::O(c)] = vol_finish(L, sum, rc, hi0 + L.ci, hj0, hk0 + L.ck);
      }
      if (in_a && in_b && on_c) {  // normal k: (2a+1, 2b+1, 2c); corners (i-,j-), (i-,j+), (i+,j-), (i+,j+); cubes k-, k+
        float sum = p[G::E(c)];
        sum += p[2 * G::PK + G::E(c)];
        sum += p[2 * G::PJ + G::E(c)];
        sum += p[2 * G::PJ + 2 * G::PK + G::E(c)];
        float rc = r6;
        if (EDGE) {
          int cnt = 4;
          if (gc >= 1) { sum += p[G::PJ + G::PK + G::O(c - 1)]; cnt++; }
          if (gc <= n - 1) { sum += p[G::PJ + G::PK + G::O(c)]; cnt++; }
          rc = cnt == 6 ? r6 : r5;
        } else {
          sum += p[G::PJ + G::PK + G::O(c - 1)];
          sum += p[G::PJ + G::PK + G::O(c)];
        }
        sm[G::row(2 * a + 1, 2 * b + 1) + G::E(c)] = vol_finish(L, sum, rc, hi0 + L.ci, hj0 + L.cj, hk0);
      }
    }
  }
  RN_VOL_SYNC();
  // ---- phase 3: the three low edges of cells 0 .. T-1 (EDGE: .. T)
  {
    constexpr int X = EDGE ? 1 : 0;
    constexpr int EJ = TJ + X, EK = TK + X, NE = (TI + X) * EJ * EK;
    for (int idx = tid; idx < NE; idx += NT) {
      const int c = idx % EK, b = (idx / EK) % EJ, a = idx / (EK * EJ);
      const int ga = A0 + a, gb = B0 + b, gc = C0 + c;
      const uint32_t hi0 = (uint32_t)(2 * ga) * L.ci, hj0 = (uint32_t)(2 * gb) * L.cj, hk0 = (uint32_t)(2 * gc) * L.ck;
      const bool in_a = a < TI && (!EDGE || ga < n), in_b = b < TJ && (!EDGE || gb < n), in_c = c < TK && (!EDGE || gc < n);
      const bool on_a = (a < TI || close_i) && (!EDGE || ga <= n), on_b = (b < TJ || close_j) && (!EDGE || gb <= n),
                 on_c = (c < TK || close_k) && (!EDGE || gc <= n);
      const float* p = sm + G::row(2 * a, 2 * b);
      // the count of an edge's neighbours: 2 ends + up to 4 faces
      if (in_a && on_b && on_c) {  // along i: (2a+1, 2b, 2c); ends i-, i+; faces j-, j+, k-, k+
        float sum = p[G::E(c)];
        sum += p[2 * G::PJ + G::E(c)];
        float rc = r6;
        if (EDGE) {
          int cnt = 2;
          if (gb >= 1) { sum += p[G::PJ - G::PK + G::E(c)]; cnt++; }
          if (gb <= n - 1) { sum += p[G::PJ + G::PK + G::E(c)]; cnt++; }
          if (gc >= 1) { sum += p[G::PJ + G::O(c - 1)]; cnt++; }
          if (gc <= n - 1) { sum += p[G::PJ + G::O(c)]; cnt++; }
          rc = cnt == 6 ? r6 : (cnt == 5 ? r5 : (cnt == 4 ? r4 : (cnt == 3 ? r3 : r2)));
        } else {
          sum += p[G::PJ - G::PK + G::E(c)];
          sum += p[G::PJ + G::PK + G::E(c)];
          sum += p[G::PJ + G::O(c - 1)];
          sum += p[G::PJ + G::O(c)];
        }
        sm[G::row(2 * a + 1, 2 * b) + G::E(c)] = vol_finish(L, sum, rc, hi0 + L.ci, hj0, hk0);
      }
      if (on_a && in_b && on_c) {  // along j: (2a, 2b+1, 2c); ends j-, j+; faces i-, i+, k-, k+
        float sum = p[G::E(c)];
        sum += p[2 * G::PK + G::E(c)];
        float rc = r6;
        if (EDGE) {
          int cnt = 2;
          if (ga >= 1) { sum += p[-G::PJ + G::PK + G::E(c)]; cnt++; }
          if (ga <= n - 1) { sum += p[G::PJ + G::PK + G::E(c)]; cnt++; }
          if (gc >= 1) { sum += p[G::PK + G::O(c - 1)]; cnt++; }
          if (gc <= n - 1) { sum += p[G::PK + G::O(c)]; cnt++; }
          rc = cnt == 6 ? r6 : (cnt == 5 ? r5 : (cnt == 4 ? r4 : (cnt == 3 ? r3 : r2)));
        } else {
          sum += p[-G::PJ + G::PK + G::E(c)];
          sum += p[G::PJ + G::PK + G::E(c)];
          sum += p[G::PK + G::O(c - 1)];
          sum += p[G::PK + G::O(c)];
        }
        sm[G::row(2 * a, 2 * b + 1) + G::E(c)] = vol_finish(L, sum, rc, hi0, hj0 + L.cj, hk0);
      }
      if (on_a && on_b && in_c) {  // along k: (2a, 2b, 2c+1); ends k-, k+; faces i-, i+, j-, j+
        float sum = p[G::E(c)];
        sum += p[G::E(c + 1)];
        float rc = r6;
        if (EDGE) {
          int cnt = 2;
          if (ga >= 1) { sum += p[-G::PJ + G::O(c)]; cnt++; }
          if (ga <= n - 1) { sum += p[G::PJ + G::O(c)]; cnt++; }
          if (gb >= 1) { sum += p[-G::PK + G::O(c)]; cnt++; }
          if (gb <= n - 1) { sum += p[G::PK + G::O(c)]; cnt++; }
          rc = cnt == 6 ? r6 : (cnt == 5 ? r5 : (cnt == 4 ? r4 : (cnt == 3 ? r3 : r2)));
        } else {
          sum += p[-G::PJ + G::O(c)];
          sum += p[G::PJ + G::O(c)];
          sum += p[-G::PK + G::O(c)];
          sum += p[G::PK + G::O(c)];
        }
        sm[G::row(2 * a, 2 * b) + G::O(c)] = vol_finish(L, sum, rc, hi0, hj0, hk0 + L.ck);
      }
    }
  }
  RN_VOL_SYNC();
  // ---- phase 4: the finished tile, full rows along k
  {
    constexpr int X = EDGE ? 1 : 0;
    constexpr int SJ = 2 * TJ + X, SK = 2 * TK + X, NS = (2 * TI + X) * SJ * SK;
    for (int idx = tid; idx < NS; idx += NT) {
      const int fk = idx % SK, fj = (idx / SK) % SJ, fi = idx / (SK * SJ);
      const int gi = 2 * A0 + fi, gj = 2 * B0 + fj, gk = 2 * C0 + fk;
      if (EDGE && ((fi == 2 * TI && !close_i) || (fj == 2 * TJ && !close_j) || (fk == 2 * TK && !close_k) || gi > 2 * n || gj > 2 * n || gk > 2 * n)) continue;
      if (gi < L.out_lo || gi > L.out_hi) continue;
      const float v = sm[G::row(fi, fj) + ((fk & 1) ? G::O(fk >> 1) : G::E(fk >> 1))];
      L.gout[((size_t)(gi - L.out_lo) * M + (size_t)gj) * M + (size_t)gk] = v;
    }
  }
}

constexpr int kVolTI = 8, kVolTJ = 8, kVolTK = 16, kVolThreads = 256;
// the warp-row kernel (rn_volrow.cuh): levels with at least kRowMinN cells per axis
#ifndef RN_VOL_ROW_WARPS
#define RN_VOL_ROW_WARPS 8   // warps per CTA of the warp-row kernel (a power of two)
#endif
constexpr int kRowTI = 4, kRowTJ = 4, kRowTK = 64, kRowWarps = RN_VOL_ROW_WARPS;
constexpr int kRowMinN = 64;
// levels with kRowMinN <= n < kRowBigN run the warp-row kernel on 2 x 2 x 64-cell tiles: the n = 64 level has only 256 of the
// 4 x 4 tiles (less than one wave), so its duration is one CTA's latency, and a 2 x 2 tile has a third of the row jobs
constexpr int kRowSmallT = 2, kRowBigN = 128;
inline int vol_row_tile(int n) { return n >= kRowBigN ? kRowTI : kRowSmallT; }

// tile extent along i of the level whose input grid has n cells per axis
// the coarse levels (8 <= n < kRowMinN) are latency-bound -- a handful of CTAs, each walking its tile phase by phase: 4 x 4 x 4-cell tiles
// with 128 threads give 8x more, 8x shorter CTAs than the 8 x 8 x 16 tile (RN_VOL_COARSE_TILE 8, the first form)
#ifndef RN_VOL_COARSE_TILE
#define RN_VOL_COARSE_TILE 4
#endif
inline int vol_tile_i(int n) { return n >= kRowMinN ? vol_row_tile(n) : (n <= 4 ? n : (RN_VOL_COARSE_TILE == 4 ? 4 : (n >= 16 ? kVolTI : n))); }


// Plane ranges: lo[l], hi[l] = the planes (lattice index along i) of the level-l grid (side 2^l + 1) that the slab
// [i0, i0+ni) of the depth-`depth` volume needs; l = 1 .. depth (lo[depth] = i0).
inline void volume_ranges(int depth, int i0, int ni, int* lo, int* hi)
{
  lo[depth] = i0;
  hi[depth] = i0 + ni - 1;
  for (int l = depth - 1; l >= 1; l--) {
    const int n = 1 << l, T = vol_tile_i(n);
    const int olo = lo[l + 1], ohi = hi[l + 1];
    const int t_lo = (olo < 2 * n ? olo : 2 * n - 1) / (2 * T);  // the closing plane 2n belongs to the last tile
    int t_hi = (ohi < 2 * n ? ohi : 2 * n - 1) / (2 * T);
    int a = t_lo * T - 1, b = (t_hi + 1) * T;
    lo[l] = a < 0 ? 0 : a;
    hi[l] = b > n ? n : b;
  }
}


// The three coarsest levels (n = 2, 4, 8: 5^3, 9^3, 17^3 samples) cost a launch each and nothing else; when the slab's cone
// holds their grids completely (every slab of a volume up to 513^3, thick slabs beyond) ONE CTA runs them back to back, each level
// as a single whole-volume tile (rn_volume.cu: k_vol_coarse).
constexpr int kVolFusedLevels = 3;
inline bool vol_coarse_fused(int depth, const int* lo, const int* hi)
{
  if (depth < kVolFusedLevels + 1) return false;
  for (int l = 1; l <= kVolFusedLevels; l++)
    if (lo[l] != 0 || hi[l] != (1 << l)) return false;
  return true;
}

// Parameters of the launch that refines the level-l grid into the level-(l+1) grid (lo/hi from volume_ranges).
inline VolLevel vol_level_make(int depth, int l, const int* lo, const int* hi, const float* gin, float* gout, uint32_t seed, float amplitude,
                               int* tiles_i)
{
  const int W = 1 << depth;
  const int n = 1 << l, T = vol_tile_i(n);
  VolLevel L;
  L.gin = gin;
  L.gout = gout;
  L.n = n;
  L.in_lo = lo[l];
  L.out_lo = lo[l + 1];
  L.out_hi = hi[l + 1];
  const int t_lo = (L.out_lo < 2 * n ? L.out_lo : 2 * n - 1) / (2 * T);  // the closing plane 2n belongs to the last tile
  const int t_hi = (L.out_hi < 2 * n ? L.out_hi : 2 * n - 1) / (2 * T);
  L.tile_i0 = t_lo;
  L.seed = seed;
  const int sh = depth - (l + 1);  // one fine step of level l+1 = 2^sh lattice units of the final volume
  L.ci = kHashI << sh;
  L.cj = kHashJ << sh;
  L.ck = kHashK << sh;
  L.amplitude = amplitude;
  L.scale = (float)(W >> l) / (float)W;  // s / W with s = W / n
  L.one = 1.0f;
  *tiles_i = t_hi - t_lo + 1;
  return L;
}

}  // namespace rn
