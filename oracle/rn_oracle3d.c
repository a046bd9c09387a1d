/* oracle/rn_oracle3d.c -- TEST INFRASTRUCTURE ONLY (see rn_oracle.h).
 *
 * CPU specification of the 3D density-volume EXTENSION (BASELINE.json configs "512^3" and "2048^3").
 * The reference has NO volume generator: PARITY UNPINNED BY THE REFERENCE.  This
 * file is the definition; it follows the reference's style -- midpoint displacement with the
 * amplitude law of Rock::generateBaseRecursive (renderEngine/source/Rock.cpp:181-188:
 * offset = amplitude * random() * extent / maxWidth, random() of include/basics.h:16-19) -- as the
 * natural 3D generalisation of Terrain.cpp:60-82: per level (step s, h = s/2) three passes,
 *   1. cube centres   (i,j,k all odd multiples of h): mean of the 8 cube corners            + offset
 *   2. face centres   (two odd, one even)           : mean of 4 face corners + the 1-2 adjacent cube centres + offset
 *   3. edge midpoints (one odd, two even)           : mean of 2 edge ends + the 2-4 adjacent face centres    + offset
 * on a (W+1)^3 lattice of float densities, W = 2^depth, levels s = W/2 .. 2 (depth-1 levels), the 27
 * points at multiples of W/2 starting at 0.  Neighbours outside [0,W] are left out of the mean.
 *
 * The draw of a point is an INTEGER LATTICE HASH of (seed, i, j, k): 2048^3 needs 8.6 G draws, more than the 2^32
 * period of the CRT LCG the 2D path reproduces, and an index-addressed generator is what makes slabs independent.
 * Round 2 changed the definition of that hash (this file IS the specification, so digests changed with it): round 1
 * mixed the 64-bit linear index with the splitmix64 finaliser, whose two 64-bit multiplications were 40 of the
 * ~100 GPU instructions per sample; now the three coordinates are multiplied by odd 32-bit constants, xor-ed with the
 * seed and passed through a 32-bit avalanche finaliser (two multiplications, three xor-shifts); the top 15 bits are
 * the draw, in the range of the CRT's rand().
 *
 * Storage: planes are allocated on demand, so a thin slab of a 2049^3 volume (whose coarse levels reach far beyond
 * the slab) can be computed without 34 GB of memory.
 */
#include "rn_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static uint32_t vol_draw(uint32_t seed, uint32_t i, uint32_t j, uint32_t k)
{
  uint32_t h = seed ^ (i * 0x9E3779B1u) ^ (j * 0x85EBCA77u) ^ (k * 0xC2B2AE3Du);
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h >> 17;
}

uint32_t rno_volume_draw(uint32_t seed, uint32_t i, uint32_t j, uint32_t k) { return vol_draw(seed, i, j, k); }

typedef struct { float** plane; int W; size_t S; uint32_t seed; float amplitude; int failed; } vol;

/* element (i,j,k); the plane is created (zero-filled) when first touched */
static float* vat(vol* v, int i, int j, int k)
{
  static float sink;
  if (!v->plane[i])
  {
    v->plane[i] = (float*)calloc(v->S * v->S, sizeof(float));
    if (!v->plane[i]) { v->failed = 1; return &sink; }
  }
  return v->plane[i] + (size_t)j * v->S + (size_t)k;
}
#define VAT(v, i, j, k) (*vat((v), (i), (j), (k)))

static float vol_offset(const vol* v, int i, int j, int k, int s)
{
  const float r = rno_random_from_draw((int)vol_draw(v->seed, (uint32_t)i, (uint32_t)j, (uint32_t)k)); /* basics.h:16-19 */
  return ((v->amplitude * r) * (float)s) / (float)v->W;                                                  /* Rock.cpp:185 */
}

static int inside(const vol* v, int c) { return c >= 0 && c <= v->W; }

/* The slab form: the level's edge midpoints are produced for planes i in [lo, hi], its face centres one half-step
 * further out (an edge midpoint reads face centres at i +- h) and its cube centres two (a face centre reads cube
 * centres at i +- h).  The full volume is lo = 0, hi = W. */
static void vol_level(vol* v, int s, int lo, int hi)
{
  const int h = s / 2, W = v->W;
  int ilo, ihi;
  const float rc[7] = {0.0f, 1.0f, 1.0f / 2.0f, 1.0f / 3.0f, 1.0f / 4.0f, 1.0f / 5.0f, 1.0f / 6.0f};
  /* pass 1: cube centres */
  ilo = lo - 2 * h; ihi = hi + 2 * h;
  for (int i = h; i < W; i += s)
  {
    if (i < ilo || i > ihi) continue;
    for (int j = h; j < W; j += s)
      for (int k = h; k < W; k += s)
      {
        float sum = VAT(v, i - h, j - h, k - h);
        sum += VAT(v, i - h, j - h, k + h);
        sum += VAT(v, i - h, j + h, k - h);
        sum += VAT(v, i - h, j + h, k + h);
        sum += VAT(v, i + h, j - h, k - h);
        sum += VAT(v, i + h, j - h, k + h);
        sum += VAT(v, i + h, j + h, k - h);
        sum += VAT(v, i + h, j + h, k + h);
        VAT(v, i, j, k) = sum * 0.125f + vol_offset(v, i, j, k, s);
      }
  }
  /* pass 2: face centres; a = the axis whose coordinate lies on the s lattice */
  ilo = lo - h; ihi = hi + h;
  for (int a = 0; a < 3; a++)
    for (int i = (a == 0 ? 0 : h); i <= W; i += s)
    {
      if (i < ilo || i > ihi) continue;
      for (int j = (a == 1 ? 0 : h); j <= W; j += s)
        for (int k = (a == 2 ? 0 : h); k <= W; k += s)
        {
          int c[3] = {i, j, k};
          const int b0 = (a + 1) % 3 < (a + 2) % 3 ? (a + 1) % 3 : (a + 2) % 3; /* the two in-face axes, ascending */
          const int b1 = (a + 1) % 3 < (a + 2) % 3 ? (a + 2) % 3 : (a + 1) % 3;
          float sum = 0.0f;
          int cnt = 0;
          for (int q = 0; q < 4; q++) /* corners: (-,-), (-,+), (+,-), (+,+) in (b0, b1) */
          {
            int p[3] = {c[0], c[1], c[2]};
            p[b0] += (q & 2) ? h : -h;
            p[b1] += (q & 1) ? h : -h;
            sum = (cnt == 0) ? VAT(v, p[0], p[1], p[2]) : sum + VAT(v, p[0], p[1], p[2]);
            cnt++;
          }
          for (int sgn = -1; sgn <= 1; sgn += 2) /* adjacent cube centres along a: -, + */
          {
            int p[3] = {c[0], c[1], c[2]};
            p[a] += sgn * h;
            if (inside(v, p[a])) { sum += VAT(v, p[0], p[1], p[2]); cnt++; }
          }
          VAT(v, i, j, k) = sum * rc[cnt] + vol_offset(v, i, j, k, s);
        }
    }
  /* pass 3: edge midpoints; a = the axis along the edge (the coordinate that is an odd multiple of h) */
  ilo = lo; ihi = hi;
  for (int a = 0; a < 3; a++)
    for (int i = (a == 0 ? h : 0); i <= W; i += s)
    {
      if (i < ilo || i > ihi) continue;
      for (int j = (a == 1 ? h : 0); j <= W; j += s)
        for (int k = (a == 2 ? h : 0); k <= W; k += s)
        {
          int c[3] = {i, j, k};
          const int b0 = (a + 1) % 3 < (a + 2) % 3 ? (a + 1) % 3 : (a + 2) % 3;
          const int b1 = (a + 1) % 3 < (a + 2) % 3 ? (a + 2) % 3 : (a + 1) % 3;
          int p[3] = {c[0], c[1], c[2]};
          p[a] -= h;
          float sum = VAT(v, p[0], p[1], p[2]); /* edge ends: -, + */
          p[a] += s;
          sum += VAT(v, p[0], p[1], p[2]);
          int cnt = 2;
          const int axes[2] = {b0, b1};
          for (int t = 0; t < 2; t++) /* adjacent face centres: along b0 (-, +), then along b1 (-, +) */
            for (int sgn = -1; sgn <= 1; sgn += 2)
            {
              int f[3] = {c[0], c[1], c[2]};
              f[axes[t]] += sgn * h;
              if (inside(v, f[axes[t]])) { sum += VAT(v, f[0], f[1], f[2]); cnt++; }
            }
          VAT(v, i, j, k) = sum * rc[cnt] + vol_offset(v, i, j, k, s);
        }
    }
}

/* What a slab needs of the level with half-step h beyond its own planes: every finer level (half-step hh) reads
 * 3*hh planes further out (edge midpoints <- face centres <- cube centres <- corners). */
static int vol_halo(int h)
{
  int e = 0;
  for (int hh = 1; hh < h; hh *= 2) e += 3 * hh;
  return e;
}

/* Planes [i0, i0+ni) of the (W+1)^3 volume into `slab` (ni x (W+1) x (W+1) floats, dense). */
int rno_volume3d(uint32_t seed, int depth, float amplitude, int i0, int ni, float* slab)
{
  if (!slab || depth < 2 || depth > 11) return RN_ERR_BAD_PARAM;
  vol v;
  v.W = 1 << depth;
  v.S = (size_t)v.W + 1;
  v.seed = seed;
  v.amplitude = amplitude;
  v.failed = 0;
  if (i0 < 0 || ni <= 0 || i0 + ni > v.W + 1) return RN_ERR_BAD_PARAM;
  v.plane = (float**)calloc(v.S, sizeof(float*));
  if (!v.plane) return RN_ERR_OOM;
  for (int i = 0; i <= v.W; i += v.W / 2)
    for (int j = 0; j <= v.W; j += v.W / 2)
      for (int k = 0; k <= v.W; k += v.W / 2) VAT(&v, i, j, k) = 0.0f;
  for (int s = v.W / 2; s >= 2; s /= 2)
  {
    const int e = vol_halo(s / 2);
    vol_level(&v, s, i0 - e, i0 + ni - 1 + e);
  }
  for (int i = 0; i < ni; i++)
  {
    if (v.plane[i0 + i]) memcpy(slab + (size_t)i * v.S * v.S, v.plane[i0 + i], v.S * v.S * sizeof(float));
    else v.failed = 1; /* every plane of the slab is written by the last level */
  }
  for (int i = 0; i <= v.W; i++) free(v.plane[i]);
  free(v.plane);
  return v.failed ? RN_ERR_OOM : RN_OK;
}
