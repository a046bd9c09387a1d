/* oracle/rn_oracle.c -- TEST INFRASTRUCTURE ONLY (see rn_oracle.h).
 *
 * Plain-C restatement of the reference's red-noise field synthesis
 * (diamond-square midpoint displacement driven by the CRT rand()).  Every
 * function cites the reference file:line it follows; paths are relative to
 * /root/reference/renderEngine/.  All arithmetic is IEEE binary32, evaluated in
 * the operand order of the cited lines, with no FMA contraction (build with
 * -ffp-contract=off, never -ffast-math).
 */
#include "rn_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ RNG --- */
/* MSVC 2008 CRT rand(): the reference's only entropy source (include/basics.h:18,
 * consumed at source/Terrain.cpp:188,233).  Published constants; reproduces the
 * known answers srand(1) -> 41 18467 6334 26500 19169 15724 11478 29358. */
#define RNO_A 214013u
#define RNO_C 2531011u

uint32_t rno_lcg_next_state(uint32_t s) { return s * RNO_A + RNO_C; }

/* state_k = A^k*seed + C*(A^k-1)/(A-1) (mod 2^32), by binary composition of the
 * affine map x -> A*x + C. */
uint32_t rno_lcg_state_after(uint32_t seed, uint64_t k)
{
  uint32_t a = RNO_A, c = RNO_C; /* map for 2^b steps */
  uint32_t s = seed;
  while (k)
  {
    if (k & 1u) s = a * s + c;
    c = (a + 1u) * c; /* (a,c) o (a,c) = (a*a, a*c + c) */
    a = a * a;
    k >>= 1;
  }
  return s;
}

uint32_t rno_lcg_draw(uint32_t seed, uint64_t k)
{
  return (rno_lcg_state_after(seed, k + 1) >> 16) & 0x7fffu;
}

/* include/basics.h:16-19: -1.0f + (2.0f * (float)rand() / (float)RAND_MAX), RAND_MAX = 0x7fff */
float rno_random_from_draw(int r)
{
  return -1.0f + (2.0f * (float)r / (float)0x7fff);
}

/* Sum over levels of 3n^2 + 2n, n = 4, 8, ..., W/2 (Terrain.cpp:60-82: one rand()
 * per generateSquare (n^2) and per generateDiamond (2 n (n+1))). */
uint64_t rno_draw_count(int depth)
{
  uint64_t total = 0;
  for (int l = 0; l < depth - 2; l++)
  {
    uint64_t n = 4ull << l;
    total += 3 * n * n + 2 * n;
  }
  return total;
}

/* Call index of the rand() that displaces point (i,j). */
uint64_t rno_call_index(int depth, int i, int j)
{
  if (depth < 3 || depth > 30) return UINT64_MAX;
  const int W = 1 << depth;
  if (i < 0 || j < 0 || i > W || j > W) return UINT64_MAX;
  uint64_t base = 0;
  for (int l = 0; l < depth - 2; l++)
  {
    const int s = W >> (l + 2), h = s >> 1;
    const uint64_t n = (uint64_t)(W / s);
    const int io = (i % s) != 0, jo = (j % s) != 0; /* off the level's corner lattice */
    if ((i % h) == 0 && (j % h) == 0 && (io || jo))
    {
      if (io && jo) return base + (uint64_t)(i / s) * n + (uint64_t)(j / s);                          /* square centre */
      if (!io) return base + n * n + (uint64_t)(i / s) * n + (uint64_t)(j / s);                       /* loop 1: (i*s, j*s+h) */
      return base + n * n + n * (n + 1) + (uint64_t)(i / s) * (n + 1) + (uint64_t)(j / s);            /* loop 2: (i*s+h, j*s) */
    }
    base += 3 * n * n + 2 * n;
  }
  return UINT64_MAX; /* one of the 25 points at multiples of W/4: never displaced */
}

/* ------------------------------------------------------------ helpers --- */
typedef struct { float x, y, z; } v3;

static inline v3 v3_of(const rn_element* e) { v3 r = {e->x, e->y, e->z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { v3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; } /* Vector3.h:61-64 */
static inline v3 v3_sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; } /* Vector3.h:75-78 */
static inline v3 v3_mul(v3 a, float f) { v3 r = {a.x * f, a.y * f, a.z * f}; return r; }    /* Vector3.h:31-34 */
/* Vector3.h:130-135 */
static inline v3 v3_cross(v3 a, v3 b)
{
  v3 r = {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
  return r;
}
/* Vector3.h:121-125 (static normalise) with magnitudeSquared :101-104 and operator/ :45-49 */
static inline v3 v3_normalise(v3 v)
{
  float f = v.x * v.x + v.y * v.y + v.z * v.z;
  if (f < 1e-10f) { v3 r = {0.0f, 0.0f, 1.0f}; return r; }
  float div = 1.0f / sqrtf(f);
  return v3_mul(v, div);
}

typedef struct
{
  const rn_params* p;
  rn_element* g;
  int W;
  size_t pitch;
  float band_hi, band_lo; /* 3.0f/featureScale, 0.5f/featureScale (Terrain.cpp:191,235) */
} field;

#define G(f, i, j) ((f)->g[(size_t)(i) * (f)->pitch + (size_t)(j)])

/* persistence^(level+2); the reference's (float)step/(float)maxWidth when persistence == 0.5 */
static float level_gain(const rn_params* p, int level)
{
  float g = 1.0f;
  for (int t = 0; t < level + 2; t++) g *= p->persistence;
  return g;
}

static int check_params(const rn_params* p)
{
  if (!p) return RN_ERR_BAD_PARAM;
  if (p->depth < RN_MIN_DEPTH || p->depth > RN_MAX_DEPTH) return RN_ERR_BAD_PARAM;
  if (p->lacunarity != 2) return RN_ERR_UNSUPPORTED;
  if (!(p->persistence > 0.0f) || !(p->persistence <= 1.0f)) return RN_ERR_BAD_PARAM;
  return RN_OK;
}

/* Terrain.cpp:44-56 */
static void field_init(field* f, const rn_params* p, rn_element* grid)
{
  f->p = p;
  f->g = grid;
  f->W = 1 << p->depth;
  f->pitch = (size_t)f->W + 1;
  f->band_hi = 3.0f / p->feature_scale;
  f->band_lo = 0.5f / p->feature_scale;
  const int W = f->W;
  for (int i = 0; i <= W; i++)
  {
    float x = 2.0f * (float)(i - W / 2) / (float)W;
    for (int j = 0; j <= W; j++)
    {
      float z = 2.0f * (float)(j - W / 2) / (float)W;
      rn_element* e = &G(f, i, j);
      e->x = x;
      e->y = z * 0.04f;
      e->z = z;
      e->shade = 1.0f;
    }
  }
}

/* The common tail of generateSquare / generateDiamond (Terrain.cpp:185-206 and
 * :230-250): scale the unit normal, draw, band / bulgy branch, amplitude law.
 * Returns the final offset; *normal is updated in place. */
static inline float displace_law(const field* f, v3* normal, float blah, int r, float my, float mx)
{
  const rn_params* p = f->p;
  *normal = v3_mul(*normal, blah);
  float offset = rno_random_from_draw(r);
  if (p->bulgy_rocks && my < f->band_hi && my > f->band_lo)
  {
    if (blah < 0.05f)
    {
      offset *= 0.75f;
      offset += 0.5f;
    }
  }
  else
  {
    normal->x *= p->sideways_scale;
    normal->z *= p->sideways_scale;
  }
  offset *= p->amplitude * (1.0f + mx * p->amplitude_variation);
  return offset;
}

/* ------------------------------------------------ sequential restatement --- */
/* Terrain.cpp:216-222 */
static void seq_average_square(field* f, int x0, int y0, int x1, int y1)
{
  const int mx = (x0 + x1) / 2, my = (y0 + y1) / 2;
  const rn_element *a = &G(f, x0, y0), *b = &G(f, x1, y0), *c = &G(f, x0, y1), *d = &G(f, x1, y1);
  rn_element* m = &G(f, mx, my);
  v3 s = v3_mul(v3_add(v3_add(v3_add(v3_of(a), v3_of(b)), v3_of(c)), v3_of(d)), 0.25f);
  m->x = s.x; m->y = s.y; m->z = s.z;
  m->shade = (a->shade + b->shade + c->shade + d->shade) * 0.25f;
}

/* Terrain.cpp:224-258 */
static void seq_generate_square(field* f, int x0, int y0, int x1, int y1, float gain, uint32_t* state)
{
  const rn_params* p = f->p;
  const int mx = (x0 + x1) / 2, my = (y0 + y1) / 2;
  rn_element* m = &G(f, mx, my);
  v3 dir1 = v3_sub(v3_of(&G(f, x1, y1)), v3_of(&G(f, x0, y0)));
  v3 dir2 = v3_sub(v3_of(&G(f, x1, y0)), v3_of(&G(f, x0, y1)));
  v3 normal = v3_normalise(v3_cross(dir1, dir2));
  float blah = (1.414f * 2.0f) * gain;
  *state = rno_lcg_next_state(*state);
  int r = (int)((*state >> 16) & 0x7fffu);
  float offset = displace_law(f, &normal, blah, r, m->y, m->x);
  m->shade += offset * p->radiosity_strength;
  G(f, x0, y0).shade -= 0.25f * offset * p->radiosity_strength;
  G(f, x1, y1).shade -= 0.25f * offset * p->radiosity_strength;
  G(f, x0, y1).shade -= 0.25f * offset * p->radiosity_strength;
  G(f, x1, y0).shade -= 0.25f * offset * p->radiosity_strength;
  v3 d = v3_mul(normal, offset);
  m->x += d.x; m->y += d.y; m->z += d.z;
}

/* Terrain.cpp:144-168 */
static void seq_average_diamond(field* f, int x0, int y0, int x1, int y1)
{
  const int W = f->W;
  const int mx = (x0 + x1) / 2, my = (y0 + y1) / 2;
  rn_element* m = &G(f, mx, my);
  v3 s;
  if (x0 < 0 || x1 > W)
  {
    if (x0 < 0) x0 = 0;
    if (x1 > W) x1 = W;
    s = v3_mul(v3_add(v3_of(&G(f, mx, y0)), v3_of(&G(f, mx, y1))), 0.5f);
  }
  else if (y0 < 0 || y1 > W)
  {
    if (y0 < 0) y0 = 0;
    if (y1 > W) y1 = W;
    s = v3_mul(v3_add(v3_of(&G(f, x0, my)), v3_of(&G(f, x1, my))), 0.5f);
  }
  else
    s = v3_mul(v3_add(v3_add(v3_add(v3_of(&G(f, x0, my)), v3_of(&G(f, x1, my))), v3_of(&G(f, mx, y0))), v3_of(&G(f, mx, y1))), 0.25f);
  m->x = s.x; m->y = s.y; m->z = s.z;
  /* always the 4-term form with the clamped indices; on a border one term is m's own stale shade */
  m->shade = (G(f, x0, my).shade + G(f, x1, my).shade + G(f, mx, y0).shade + G(f, mx, y1).shade) * 0.25f;
}

/* Terrain.cpp:171-214.  `gain` stands in for (float)step/(float)maxWidth; the clamped
 * X extent halves it on an X border (Terrain.cpp:186). */
static void seq_generate_diamond(field* f, int x0, int y0, int x1, int y1, float gain, uint32_t* state)
{
  const rn_params* p = f->p;
  const int W = f->W;
  const int step = x1 - x0;
  const int mx = (x0 + x1) / 2, my = (y0 + y1) / 2;
  if (x0 < 0) x0 = 0;
  if (x1 > W) x1 = W;
  if (y0 < 0) y0 = 0;
  if (y1 > W) y1 = W;
  rn_element* m = &G(f, mx, my);
  v3 dir1 = v3_sub(v3_of(&G(f, x0, my)), v3_of(&G(f, x1, my)));
  v3 dir2 = v3_sub(v3_of(&G(f, mx, y1)), v3_of(&G(f, mx, y0)));
  v3 normal = v3_normalise(v3_cross(dir1, dir2));
  float frac = ((float)x1 - (float)x0) / (float)step; /* 1 or 0.5, exact */
  float blah = 2.0f * (gain * frac);
  *state = rno_lcg_next_state(*state);
  int r = (int)((*state >> 16) & 0x7fffu);
  float offset = displace_law(f, &normal, blah, r, m->y, m->x);
  m->shade += offset * p->radiosity_strength;
  G(f, mx, y0).shade -= 0.25f * offset * p->radiosity_strength;
  G(f, mx, y1).shade -= 0.25f * offset * p->radiosity_strength;
  G(f, x0, my).shade -= 0.25f * offset * p->radiosity_strength;
  G(f, x1, my).shade -= 0.25f * offset * p->radiosity_strength;
  v3 d = v3_mul(normal, offset);
  m->x += d.x; m->y += d.y; m->z += d.z;
}

/* Terrain.cpp:58-82 */
int rno_terrain2d_seq(const rn_params* p, rn_element* grid)
{
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!grid) return RN_ERR_BAD_PARAM;
  field f;
  field_init(&f, p, grid);
  const int W = f.W;
  uint32_t state = p->seed; /* srand(seed) */
  int step = W / 4;
  int level = 0;
  do
  {
    const float gain = level_gain(p, level);
    for (int x = 0; x < W; x += step)
      for (int y = 0; y < W; y += step) seq_average_square(&f, x, y, x + step, y + step);
    for (int x = 0; x < W; x += step)
      for (int y = 0; y < W; y += step) seq_generate_square(&f, x, y, x + step, y + step, gain, &state);
    for (int x = -step / 2; x < W; x += step)
      for (int y = 0; y < W; y += step) seq_average_diamond(&f, x, y, x + step, y + step);
    for (int x = 0; x < W; x += step)
      for (int y = -step / 2; y < W; y += step) seq_average_diamond(&f, x, y, x + step, y + step);
    for (int x = -step / 2; x < W; x += step)
      for (int y = 0; y < W; y += step) seq_generate_diamond(&f, x, y, x + step, y + step, gain, &state);
    for (int x = 0; x < W; x += step)
      for (int y = -step / 2; y < W; y += step) seq_generate_diamond(&f, x, y, x + step, y + step, gain, &state);
    step /= 2;
    level++;
  } while (step > 1);
  return RN_OK;
}

/* --------------------------------------------- index-addressed restatement --- */
/* Same arithmetic, but every sample is a pure function of (seed, call index,
 * neighbours): the formulation the CUDA kernels use.  Shade is accumulated by an
 * ordered gather instead of the reference's scatter. */
static inline int in_range(int v, int W) { return v >= 0 && v <= W; }

int rno_terrain2d_indexed(const rn_params* p, rn_element* grid, int reverse)
{
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!grid) return RN_ERR_BAD_PARAM;
  field f;
  field_init(&f, p, grid);
  const int W = f.W;
  const float rs = p->radiosity_strength;
  float* off = (float*)malloc(sizeof(float) * f.pitch * f.pitch); /* offset of the point created at (i,j) */
  if (!off) return RN_ERR_OOM;
#define OFF(i, j) off[(size_t)(i) * f.pitch + (size_t)(j)]
#define T_OF(i, j) ((0.25f * OFF(i, j)) * rs)
  uint64_t base = 0;
  for (int level = 0; level < p->depth - 2; level++)
  {
    const int s = W >> (level + 2), h = s >> 1, n = W / s;
    const float gain = level_gain(p, level);
    const uint64_t un = (uint64_t)n;
    /* S stage: averageSquare + generateSquare fused per centre */
    for (int q = 0; q < n * n; q++)
    {
      const int idx = reverse ? n * n - 1 - q : q;
      const int i = idx / n, j = idx % n;
      const int x0 = i * s, y0 = j * s, x1 = x0 + s, y1 = y0 + s;
      const rn_element *a = &G(&f, x0, y0), *b = &G(&f, x1, y0), *c = &G(&f, x0, y1), *d = &G(&f, x1, y1);
      rn_element* m = &G(&f, x0 + h, y0 + h);
      v3 avg = v3_mul(v3_add(v3_add(v3_add(v3_of(a), v3_of(b)), v3_of(c)), v3_of(d)), 0.25f);
      float sh = (a->shade + b->shade + c->shade + d->shade) * 0.25f;
      v3 normal = v3_normalise(v3_cross(v3_sub(v3_of(d), v3_of(a)), v3_sub(v3_of(b), v3_of(c))));
      int r = (int)rno_lcg_draw(p->seed, base + (uint64_t)i * un + (uint64_t)j);
      float o = displace_law(&f, &normal, (1.414f * 2.0f) * gain, r, avg.y, avg.x);
      v3 dd = v3_mul(normal, o);
      m->x = avg.x + dd.x; m->y = avg.y + dd.y; m->z = avg.z + dd.z;
      m->shade = sh + o * rs;
      OFF(x0 + h, y0 + h) = o;
    }
    /* corners gather the S-stage scatter in square order (x outer, y inner) */
    for (int q = 0; q < (n + 1) * (n + 1); q++)
    {
      const int idx = reverse ? (n + 1) * (n + 1) - 1 - q : q;
      const int px = (idx / (n + 1)) * s, py = (idx % (n + 1)) * s;
      float sh = G(&f, px, py).shade;
      if (px - s >= 0 && py - s >= 0) sh -= T_OF(px - h, py - h);
      if (px - s >= 0 && py < W) sh -= T_OF(px - h, py + h);
      if (px < W && py - s >= 0) sh -= T_OF(px + h, py - h);
      if (px < W && py < W) sh -= T_OF(px + h, py + h);
      G(&f, px, py).shade = sh;
    }
    /* D stage: averageDiamond + generateDiamond fused per edge midpoint; both loops */
    const int nd = 2 * n * (n + 1);
    for (int q = 0; q < nd; q++)
    {
      const int idx = reverse ? nd - 1 - q : q;
      int mx, my, i, j;
      uint64_t k;
      if (idx < n * (n + 1))
      { /* loop 1: (i*s, j*s+h), i in [0,n], j in [0,n) */
        i = idx / n; j = idx % n;
        mx = i * s; my = j * s + h;
        k = base + un * un + (uint64_t)i * un + (uint64_t)j;
      }
      else
      { /* loop 2: (i*s+h, j*s), i in [0,n), j in [0,n] */
        const int e = idx - n * (n + 1);
        i = e / (n + 1); j = e % (n + 1);
        mx = i * s + h; my = j * s;
        k = base + un * un + un * (un + 1) + (uint64_t)i * (un + 1) + (uint64_t)j;
      }
      int x0 = mx - h, x1 = mx + h, y0 = my - h, y1 = my + h;
      rn_element* m = &G(&f, mx, my);
      v3 avg;
      const int xb = (x0 < 0 || x1 > W), yb = (y0 < 0 || y1 > W);
      if (x0 < 0) x0 = 0;
      if (x1 > W) x1 = W;
      if (y0 < 0) y0 = 0;
      if (y1 > W) y1 = W;
      if (xb) avg = v3_mul(v3_add(v3_of(&G(&f, mx, y0)), v3_of(&G(&f, mx, y1))), 0.5f);
      else if (yb) avg = v3_mul(v3_add(v3_of(&G(&f, x0, my)), v3_of(&G(&f, x1, my))), 0.5f);
      else avg = v3_mul(v3_add(v3_add(v3_add(v3_of(&G(&f, x0, my)), v3_of(&G(&f, x1, my))), v3_of(&G(&f, mx, y0))), v3_of(&G(&f, mx, y1))), 0.25f);
      /* m's own (stale) shade is 1.0f: it has never been written before this level */
      float sh = (G(&f, x0, my).shade + G(&f, x1, my).shade + G(&f, mx, y0).shade + G(&f, mx, y1).shade) * 0.25f;
      m->x = avg.x; m->y = avg.y; m->z = avg.z; /* generateDiamond reads m.pos through the clamped index */
      v3 dir1 = v3_sub(v3_of(&G(&f, x0, my)), v3_of(&G(&f, x1, my)));
      v3 dir2 = v3_sub(v3_of(&G(&f, mx, y1)), v3_of(&G(&f, mx, y0)));
      v3 normal = v3_normalise(v3_cross(dir1, dir2));
      float frac = ((float)x1 - (float)x0) / (float)s;
      int r = (int)rno_lcg_draw(p->seed, k);
      float o = displace_law(&f, &normal, 2.0f * (gain * frac), r, avg.y, avg.x);
      sh += o * rs;
      if (xb || yb) sh -= (0.25f * o) * rs; /* the clamped neighbour is m itself */
      v3 dd = v3_mul(normal, o);
      m->x = avg.x + dd.x; m->y = avg.y + dd.y; m->z = avg.z + dd.z;
      m->shade = sh;
      OFF(mx, my) = o;
    }
    /* corners and centres gather the D-stage scatter: loop 1 sources, then loop 2 */
    for (int q = 0; q < (2 * n + 1) * (2 * n + 1); q++)
    {
      const int idx = reverse ? (2 * n + 1) * (2 * n + 1) - 1 - q : q;
      const int a = idx / (2 * n + 1), b = idx % (2 * n + 1);
      if ((a & 1) != (b & 1)) continue; /* an edge midpoint: receives nothing more at this level */
      const int px = a * h, py = b * h;
      float sh = G(&f, px, py).shade;
      if (!(a & 1))
      { /* corner: loop-1 midpoints share its x, loop-2 midpoints share its y */
        if (in_range(py - h, W)) sh -= T_OF(px, py - h);
        if (in_range(py + h, W)) sh -= T_OF(px, py + h);
        if (in_range(px - h, W)) sh -= T_OF(px - h, py);
        if (in_range(px + h, W)) sh -= T_OF(px + h, py);
      }
      else
      { /* square centre: loop-1 midpoints are its x neighbours, loop-2 its y neighbours */
        sh -= T_OF(px - h, py);
        sh -= T_OF(px + h, py);
        sh -= T_OF(px, py - h);
        sh -= T_OF(px, py + h);
      }
      G(&f, px, py).shade = sh;
    }
    base += 3 * un * un + 2 * un;
  }
#undef OFF
#undef T_OF
  free(off);
  return RN_OK;
}

/* ---------------------------------------------------------------- volcano --- */
/* Terrain.cpp:121-141 and :89 */
void rno_add_volcano(const rn_params* p, rn_element* grid, float* volcano_pos)
{
  const int W = 1 << p->depth;
  const size_t pitch = (size_t)W + 1;
  const int vx = W / 2, vz = W / 2;
  for (int x = 0; x <= W; x++)
    for (int y = 0; y <= W; y++)
    {
      float fx = (float)(x - vx), fy = (float)(y - vz);
      float d = sqrtf(fx * fx + fy * fy);
      d /= (float)W;
      float height = 1.0f / d;
      float minRad = 0.02f;
      if (d < minRad) height = 60000.f * (d * d - minRad * minRad) + (1 / minRad);
      height *= p->amplitude * 0.008f;
      grid[(size_t)x * pitch + (size_t)y].y += height;
    }
  if (volcano_pos)
  {
    const rn_element* e = &grid[(size_t)vx * pitch + (size_t)vz];
    volcano_pos[0] = e->x * p->size;
    volcano_pos[1] = e->y * p->size;
    volcano_pos[2] = e->z * p->size;
  }
}

int rno_terrain2d(const rn_params* p, rn_element* grid, float* volcano_pos)
{
  int st = rno_terrain2d_seq(p, grid);
  if (st != RN_OK) return st;
  if (p->apply_volcano) rno_add_volcano(p, grid, volcano_pos);
  return RN_OK;
}

typedef struct
{
  const rn_params* p;
  const uint32_t* seeds;
  rn_element* grids;
  int ntiles, nthreads, tid, bad;
} batch_job;

static void* batch_worker(void* arg)
{
  batch_job* jb = (batch_job*)arg;
  const size_t side = ((size_t)1 << jb->p->depth) + 1;
  for (int t = jb->tid; t < jb->ntiles; t += jb->nthreads)
  {
    rn_params q = *jb->p;
    q.seed = jb->seeds[t];
    if (rno_terrain2d(&q, jb->grids + (size_t)t * side * side, 0) != RN_OK) jb->bad = 1;
  }
  return 0;
}

/* The restatement keeps no global state (unlike the reference's rand()), so
 * independent seeds run on independent host threads. */
int rno_terrain2d_batch(const rn_params* p, const uint32_t* seeds, int ntiles, rn_element* grids, int threads)
{
  int st = check_params(p);
  if (st != RN_OK) return st;
  if (!seeds || !grids || ntiles < 0) return RN_ERR_BAD_PARAM;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads > ntiles) threads = ntiles > 0 ? ntiles : 1;
  pthread_t th[256];
  batch_job jobs[256];
  for (int t = 0; t < threads; t++)
  {
    batch_job jb = {p, seeds, grids, ntiles, threads, t, 0};
    jobs[t] = jb;
    if (t > 0 && pthread_create(&th[t], 0, batch_worker, &jobs[t]) != 0) return RN_ERR_OOM;
  }
  batch_worker(&jobs[0]);
  int bad = jobs[0].bad;
  for (int t = 1; t < threads; t++)
  {
    pthread_join(th[t], 0);
    bad |= jobs[t].bad;
  }
  return bad ? RN_ERR_BAD_PARAM : RN_OK;
}

/* FNV-1a 64 over raw bytes (the digest used by tests/golden/). */
uint64_t rno_fnv64(const void* data, size_t nbytes)
{
  const unsigned char* b = (const unsigned char*)data;
  uint64_t hsh = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < nbytes; i++)
  {
    hsh ^= b[i];
    hsh *= 0x100000001b3ull;
  }
  return hsh;
}
