// rn_volume.cu -- 3D density-volume EXTENSION (BASELINE.json configs 512^3 / 2048^3).
//
// The reference has no volume generator; the definition is the CPU specification described in DESIGN.md section 5.8 (3D
// midpoint displacement in the reference's style: cube centres, face centres, edge midpoints per level,
// amplitude law of Rock.cpp:185, random() of basics.h:16-19).  This file is the level-synchronous CUDA form:
// one launch per (level, pass, axis), one thread per new sample, k (the contiguous index) across the lanes.
// A sample's draw is a pure function of (seed, lattice index), and a slab of planes needs only a bounded
// neighbourhood of every coarser level, so z-slabs (here: i-slabs) of one volume are independent -- the
// multi-GPU form recomputes that halo instead of exchanging it.  Bit-exact against the oracle.
#include "rn_kernels.h"

namespace rn {

namespace {

__device__ __forceinline__ uint32_t vol_draw(uint32_t seed, uint64_t idx)
{
  uint64_t z = idx + 0x9E3779B97F4A7C15ull * ((uint64_t)seed + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 33) & 0x7fffu;
}

struct VolLevel {
  float* d;
  int W, s, h;
  uint32_t seed;
  float amplitude, inv_w;
  int ia0;  // first index along i of this launch (the slab restriction)
};

// PASS 1: cube centres (AXIS unused).  PASS 2: face centres, AXIS = the coordinate on the s lattice.
// PASS 3: edge midpoints, AXIS = the coordinate that is an odd multiple of h.
template <int PASS, int AXIS>
__global__ void k_volume(VolLevel L, int cnt_k)
{
  const int kk = blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= cnt_k) return;
  const int h = L.h, s = L.s;
  // odd[axis]: the coordinate is an odd multiple of h
  const bool odd0 = PASS == 1 || (PASS == 2 && AXIS != 0) || (PASS == 3 && AXIS == 0);
  const bool odd1 = PASS == 1 || (PASS == 2 && AXIS != 1) || (PASS == 3 && AXIS == 1);
  const bool odd2 = PASS == 1 || (PASS == 2 && AXIS != 2) || (PASS == 3 && AXIS == 2);
  const int i = (L.ia0 + (int)blockIdx.z) * s + (odd0 ? h : 0);
  const int j = (int)blockIdx.y * s + (odd1 ? h : 0);
  const int k = kk * s + (odd2 ? h : 0);
  // linear index of the sample and the three half-step strides: every neighbour is base +- a stride
  const size_t S = (size_t)L.W + 1;
  const ptrdiff_t st[3] = {(ptrdiff_t)(S * S) * h, (ptrdiff_t)S * h, (ptrdiff_t)h};
  const size_t base = ((size_t)i * S + (size_t)j) * S + (size_t)k;
  const float* p = L.d + base;
  float sum;
  float rcnt;
  if (PASS == 1) {
    sum = p[-st[0] - st[1] - st[2]];
    sum += p[-st[0] - st[1] + st[2]];
    sum += p[-st[0] + st[1] - st[2]];
    sum += p[-st[0] + st[1] + st[2]];
    sum += p[st[0] - st[1] - st[2]];
    sum += p[st[0] - st[1] + st[2]];
    sum += p[st[0] + st[1] - st[2]];
    sum += p[st[0] + st[1] + st[2]];
    rcnt = 0.125f;
  } else {
    constexpr int B0 = AXIS == 0 ? 1 : 0;               // the other two axes, ascending
    constexpr int B1 = AXIS == 2 ? 1 : 2;
    const int c[3] = {i, j, k};
    int cnt;
    if (PASS == 2) {
      // 4 face corners (-,-), (-,+), (+,-), (+,+) in (B0, B1), then the cube centres along AXIS: -, +
      sum = p[-st[B0] - st[B1]];
      sum += p[-st[B0] + st[B1]];
      sum += p[st[B0] - st[B1]];
      sum += p[st[B0] + st[B1]];
      cnt = 4;
      if (c[AXIS] - h >= 0) { sum += p[-st[AXIS]]; cnt++; }
      if (c[AXIS] + h <= L.W) { sum += p[st[AXIS]]; cnt++; }
    } else {
      // 2 edge ends along AXIS: -, +; then the face centres along B0 (-, +), then along B1 (-, +)
      sum = p[-st[AXIS]];
      sum += p[st[AXIS]];
      cnt = 2;
      if (c[B0] - h >= 0) { sum += p[-st[B0]]; cnt++; }
      if (c[B0] + h <= L.W) { sum += p[st[B0]]; cnt++; }
      if (c[B1] - h >= 0) { sum += p[-st[B1]]; cnt++; }
      if (c[B1] + h <= L.W) { sum += p[st[B1]]; cnt++; }
    }
    rcnt = cnt == 6 ? 1.0f / 6.0f : (cnt == 5 ? 1.0f / 5.0f : (cnt == 4 ? 1.0f / 4.0f : (cnt == 3 ? 1.0f / 3.0f : 1.0f / 2.0f)));
  }
  const float r = random_from_draw(vol_draw(L.seed, (uint64_t)base));
  L.d[base] = sum * rcnt + ((L.amplitude * r) * (float)L.s) * L.inv_w;
}

__global__ void k_volume_init(float* d, int W)
{
  const int t = threadIdx.x;
  if (t >= 27) return;
  const size_t S = (size_t)W + 1;
  const int i = (t / 9) * (W / 2), j = ((t / 3) % 3) * (W / 2), k = (t % 3) * (W / 2);
  d[((size_t)i * S + (size_t)j) * S + (size_t)k] = 0.0f;
}

// indices a along an axis whose coordinate a*s + off lies in [lo, hi] and in [0, W]
inline void index_range(int s, int off, int cnt, int lo, int hi, int* a0, int* a1)
{
  int x0 = lo - off, x1 = hi - off;
  *a0 = x0 <= 0 ? 0 : (x0 + s - 1) / s;
  *a1 = x1 < 0 ? -1 : x1 / s;
  if (*a1 > cnt - 1) *a1 = cnt - 1;
}

template <int PASS, int AXIS>
cudaError_t launch_pass(VolLevel L, int lo, int hi, cudaStream_t stream, uint64_t* launches)
{
  const int n = L.W / L.s;
  const bool odd0 = PASS == 1 || (PASS == 2 && AXIS != 0) || (PASS == 3 && AXIS == 0);
  const bool odd1 = PASS == 1 || (PASS == 2 && AXIS != 1) || (PASS == 3 && AXIS == 1);
  const bool odd2 = PASS == 1 || (PASS == 2 && AXIS != 2) || (PASS == 3 && AXIS == 2);
  const int cnt_i = odd0 ? n : n + 1, cnt_j = odd1 ? n : n + 1, cnt_k = odd2 ? n : n + 1;
  int a0, a1;
  index_range(L.s, odd0 ? L.h : 0, cnt_i, lo, hi, &a0, &a1);
  if (a1 < a0) return cudaSuccess;
  L.ia0 = a0;
  const int threads = cnt_k >= 128 ? 128 : 32;
  const dim3 grid((unsigned)((cnt_k + threads - 1) / threads), (unsigned)cnt_j, (unsigned)(a1 - a0 + 1));
  k_volume<PASS, AXIS><<<grid, threads, 0, stream>>>(L, cnt_k);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace

// Fills planes [i0, i0+ni) of the (W+1)^3 volume `d` (device memory, full-volume layout).
cudaError_t volume_generate(float* d, int depth, uint32_t seed, float amplitude, int i0, int ni, cudaStream_t stream, uint64_t* launches)
{
  const int W = 1 << depth;
  k_volume_init<<<1, 32, 0, stream>>>(d, W);
  if (launches) *launches += 1;
  for (int s = W / 2; s >= 2; s /= 2) {
    const int h = s / 2;
    int e = 0;
    for (int hh = 1; hh < h; hh *= 2) e += 3 * hh;  // reach of the finer levels
    const int lo = i0 - e, hi = i0 + ni - 1 + e;
    VolLevel L;
    L.d = d; L.W = W; L.s = s; L.h = h; L.seed = seed; L.amplitude = amplitude; L.inv_w = 1.0f / (float)W; L.ia0 = 0;
    cudaError_t err;
    if ((err = launch_pass<1, 0>(L, lo - 2 * h, hi + 2 * h, stream, launches)) != cudaSuccess) return err;
    if ((err = launch_pass<2, 0>(L, lo - h, hi + h, stream, launches)) != cudaSuccess) return err;
    if ((err = launch_pass<2, 1>(L, lo - h, hi + h, stream, launches)) != cudaSuccess) return err;
    if ((err = launch_pass<2, 2>(L, lo - h, hi + h, stream, launches)) != cudaSuccess) return err;
    if ((err = launch_pass<3, 0>(L, lo, hi, stream, launches)) != cudaSuccess) return err;
    if ((err = launch_pass<3, 1>(L, lo, hi, stream, launches)) != cudaSuccess) return err;
    if ((err = launch_pass<3, 2>(L, lo, hi, stream, launches)) != cudaSuccess) return err;
  }
  return cudaSuccess;
}

}  // namespace rn
