// rn_levelsync.cu -- algorithm 1: level-synchronous diamond-square in global memory.
//
// The straightforward index-addressed form of Terrain.cpp:60-82: four launches per
// level (square stage, corner gather, diamond stage, corner+centre gather), one
// thread per sample, the whole field resident in HBM.  It is the simplest correct
// CUDA statement of the path and is kept as (a) the cross-check for the fused tile
// kernels (rn_fused.cu) and (b) the generator of the small coarse grids.
//
// Shade is accumulated by ordered gather, never by atomics, so
// the result is bit-identical to the reference's sequential scatter.
#include "rn_kernels.h"

namespace rn {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ Elem ld(const Elem* g, size_t pitch, int i, int j)
{
  const float4 v = *reinterpret_cast<const float4*>(g + (size_t)i * pitch + (size_t)j);
  return Elem{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ void st(Elem* g, size_t pitch, int i, int j, V3 p, float shade)
{
  *reinterpret_cast<float4*>(g + (size_t)i * pitch + (size_t)j) = make_float4(p.x, p.y, p.z, shade);
}

// Terrain.cpp:46-56 restricted to the 25 lattice points at multiples of W/4 -- the only
// samples whose initial value is ever read (every other sample is written before use).
__global__ void k_seed_lattice(Elem* g, size_t pitch, int W)
{
  const int t = threadIdx.x;
  if (t >= 25) return;
  const int i = (t / 5) * (W / 4), j = (t % 5) * (W / 4);
  const Elem e = initial_elem(i, j, W);
  st(g, pitch, i, j, pos_of(e), e.shade);
}

// averageSquare + generateSquare (Terrain.cpp:216-258), one thread per square centre.
__global__ void k_square(Elem* g, float* off, size_t pitch, LevelInfo li, Law law, uint32_t seed)
{
  const int n = li.n;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)n * (unsigned)n) return;
  const int i = idx / n, j = idx % n;
  const int x0 = i * li.s, y0 = j * li.s, x1 = x0 + li.s, y1 = y0 + li.s;
  const Elem a = ld(g, pitch, x0, y0), b = ld(g, pitch, x1, y0), c = ld(g, pitch, x0, y1), d = ld(g, pitch, x1, y1);
  const V3 avg = vmul(vadd(vadd(vadd(pos_of(a), pos_of(b)), pos_of(c)), pos_of(d)), 0.25f);
  const float sh = (a.shade + b.shade + c.shade + d.shade) * 0.25f;
  const uint32_t k = li.base + (uint32_t)i * (uint32_t)n + (uint32_t)j;
  const uint32_t draw = lcg_draw_from_state(lcg_state_after(seed, (uint64_t)k + 1));
  const Displaced dp = displace(law, avg, vsub(pos_of(d), pos_of(a)), vsub(pos_of(b), pos_of(c)), (1.414f * 2.0f) * li.gain, draw);
  st(g, pitch, x0 + li.h, y0 + li.h, dp.pos, sh + shade_gain(law, dp.offset));
  off[(size_t)(x0 + li.h) * pitch + (size_t)(y0 + li.h)] = dp.offset;
}

// Corners receive the square stage's scatter (Terrain.cpp:252-255) in call order.
__global__ void k_square_gather(Elem* g, const float* off, size_t pitch, LevelInfo li, Law law, int W)
{
  const int n1 = li.n + 1;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)n1 * (unsigned)n1) return;
  const int px = (idx / n1) * li.s, py = (idx % n1) * li.s, h = li.h;
  float* shp = &g[(size_t)px * pitch + (size_t)py].shade;
  float sh = *shp;
  const bool xl = px > 0, xh = px < W, yl = py > 0, yh = py < W;
  if (xl && yl) sh -= shade_loss(law, off[(size_t)(px - h) * pitch + (size_t)(py - h)]);
  if (xl && yh) sh -= shade_loss(law, off[(size_t)(px - h) * pitch + (size_t)(py + h)]);
  if (xh && yl) sh -= shade_loss(law, off[(size_t)(px + h) * pitch + (size_t)(py - h)]);
  if (xh && yh) sh -= shade_loss(law, off[(size_t)(px + h) * pitch + (size_t)(py + h)]);
  *shp = sh;
}

// averageDiamond + generateDiamond (Terrain.cpp:144-214), one thread per edge midpoint,
// both loops of Terrain.cpp:69-80 in one launch.
__global__ void k_diamond(Elem* g, float* off, size_t pitch, LevelInfo li, Law law, uint32_t seed, int W)
{
  const unsigned n = li.n;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned half = n * (n + 1);
  if (idx >= 2 * half) return;
  int mx, my;
  uint32_t k;
  if (idx < half) {  // loop 1 (Terrain.cpp:75-77): midpoint (i*s, j*s+h)
    const unsigned i = idx / n, j = idx % n;
    mx = i * li.s;
    my = j * li.s + li.h;
    k = li.base + n * n + i * n + j;
  } else {  // loop 2 (Terrain.cpp:78-80): midpoint (i*s+h, j*s)
    const unsigned e = idx - half, i = e / (n + 1), j = e % (n + 1);
    mx = i * li.s + li.h;
    my = j * li.s;
    k = li.base + n * n + half + i * (n + 1) + j;
  }
  int x0 = mx - li.h, x1 = mx + li.h, y0 = my - li.h, y1 = my + li.h;
  const bool xb = (x0 < 0) || (x1 > W), yb = (y0 < 0) || (y1 > W);
  x0 = max(x0, 0); x1 = min(x1, W); y0 = max(y0, 0); y1 = min(y1, W);
  // On a border one clamped neighbour is the midpoint itself: its position is the
  // fresh average (averageDiamond ran first) and its shade is the untouched 1.0f.
  Elem ex0, ex1, ey0, ey1;
  const Elem self_stale{0.f, 0.f, 0.f, 1.0f};
  ex0 = (x0 == mx) ? self_stale : ld(g, pitch, x0, my);
  ex1 = (x1 == mx) ? self_stale : ld(g, pitch, x1, my);
  ey0 = (y0 == my) ? self_stale : ld(g, pitch, mx, y0);
  ey1 = (y1 == my) ? self_stale : ld(g, pitch, mx, y1);
  V3 avg;
  if (xb) avg = vmul(vadd(pos_of(ey0), pos_of(ey1)), 0.5f);
  else if (yb) avg = vmul(vadd(pos_of(ex0), pos_of(ex1)), 0.5f);
  else avg = vmul(vadd(vadd(vadd(pos_of(ex0), pos_of(ex1)), pos_of(ey0)), pos_of(ey1)), 0.25f);
  float sh = (ex0.shade + ex1.shade + ey0.shade + ey1.shade) * 0.25f;
  const V3 px0 = (x0 == mx) ? avg : pos_of(ex0), px1 = (x1 == mx) ? avg : pos_of(ex1);
  const V3 py0 = (y0 == my) ? avg : pos_of(ey0), py1 = (y1 == my) ? avg : pos_of(ey1);
  const float frac = xb ? 0.5f : 1.0f;  // ((float)maxX-(float)minX)/step after clamping (Terrain.cpp:186)
  const uint32_t draw = lcg_draw_from_state(lcg_state_after(seed, (uint64_t)k + 1));
  const Displaced dp = displace(law, avg, vsub(px0, px1), vsub(py1, py0), 2.0f * (li.gain * frac), draw);
  sh += shade_gain(law, dp.offset);
  if (xb || yb) sh -= shade_loss(law, dp.offset);  // the scatter onto the clamped index hits the midpoint itself
  st(g, pitch, mx, my, dp.pos, sh);
  off[(size_t)mx * pitch + (size_t)my] = dp.offset;
}

// Corners and square centres receive the diamond stage's scatter (Terrain.cpp:208-211):
// all loop-1 sources first, then loop-2 sources, each in call order.
__global__ void k_diamond_gather(Elem* g, const float* off, size_t pitch, LevelInfo li, Law law, int W)
{
  const int m = 2 * li.n + 1;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)m * (unsigned)m) return;
  const int a = idx / m, b = idx % m;
  if ((a & 1) != (b & 1)) return;  // an edge midpoint: nothing more lands on it at this level
  const int h = li.h, px = a * h, py = b * h;
  float* shp = &g[(size_t)px * pitch + (size_t)py].shade;
  float sh = *shp;
  if (!(a & 1)) {  // corner: loop-1 midpoints share its row index, loop-2 midpoints its column index
    if (py - h >= 0) sh -= shade_loss(law, off[(size_t)px * pitch + (size_t)(py - h)]);
    if (py + h <= W) sh -= shade_loss(law, off[(size_t)px * pitch + (size_t)(py + h)]);
    if (px - h >= 0) sh -= shade_loss(law, off[(size_t)(px - h) * pitch + (size_t)py]);
    if (px + h <= W) sh -= shade_loss(law, off[(size_t)(px + h) * pitch + (size_t)py]);
  } else {  // square centre
    sh -= shade_loss(law, off[(size_t)(px - h) * pitch + (size_t)py]);
    sh -= shade_loss(law, off[(size_t)(px + h) * pitch + (size_t)py]);
    sh -= shade_loss(law, off[(size_t)px * pitch + (size_t)(py - h)]);
    sh -= shade_loss(law, off[(size_t)px * pitch + (size_t)(py + h)]);
  }
  *shp = sh;
}

inline unsigned blocks_for(uint64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

// Runs levels [0, nlevels) of a depth-`depth` field on `g` (pitch in elements);
// `off` is a scratch plane with the same pitch and (W+1) rows.
cudaError_t levelsync_generate(Elem* g, float* off, size_t pitch, int depth, int nlevels, const Law& law, float persistence,
                               uint32_t seed, cudaStream_t stream, uint64_t* launches)
{
  const int W = 1 << depth;
  k_seed_lattice<<<1, 32, 0, stream>>>(g, pitch, W);
  uint64_t nl = 1;
  for (int level = 0; level < nlevels; level++) {
    const LevelInfo li = level_info(depth, level, persistence);
    const uint64_t n = (uint64_t)li.n;
    k_square<<<blocks_for(n * n), kThreads, 0, stream>>>(g, off, pitch, li, law, seed);
    k_square_gather<<<blocks_for((n + 1) * (n + 1)), kThreads, 0, stream>>>(g, off, pitch, li, law, W);
    k_diamond<<<blocks_for(2 * n * (n + 1)), kThreads, 0, stream>>>(g, off, pitch, li, law, seed, W);
    k_diamond_gather<<<blocks_for((2 * n + 1) * (2 * n + 1)), kThreads, 0, stream>>>(g, off, pitch, li, law, W);
    nl += 4;
  }
  if (launches) *launches += nl;
  return cudaGetLastError();
}

}  // namespace rn
