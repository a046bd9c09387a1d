// rn_util.cu -- small kernels around the synthesis: addVolcano epilogue and the raw
// LCG draw stream (the integer part of the path, exposed for bit-exact tests).
#include "rn_kernels.h"

namespace rn {

namespace {

// Terrain.cpp:121-141, one thread per sample of the window; j is the fast index so a warp
// touches 512 contiguous bytes of one row.
__global__ void k_volcano(Elem* g, size_t pitch, int W, int i0, int j0, int ni, int nj, float amplitude)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= nj || i >= ni) return;
  float* y = &g[(size_t)i * pitch + (size_t)j].y;
  *y = *y + volcano_height(i0 + i, j0 + j, W, amplitude);
}

// One thread per 8 consecutive draws: one skip-ahead, then 8 sequential LCG steps,
// packed into one 128-bit store.
__global__ void k_lcg_draws(uint16_t* dst, uint32_t seed, uint64_t first, uint64_t n)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t k0 = t * 8;
  if (k0 >= n) return;
  uint32_t s = lcg_state_after(seed, first + k0);
  uint16_t v[8];
#pragma unroll
  for (int q = 0; q < 8; q++) {
    s = s * kLcgA + kLcgC;
    v[q] = (uint16_t)lcg_draw_from_state(s);
  }
  if (k0 + 8 <= n && (reinterpret_cast<uintptr_t>(dst + k0) & 15u) == 0) {
    uint4 pk;
    pk.x = v[0] | ((uint32_t)v[1] << 16);
    pk.y = v[2] | ((uint32_t)v[3] << 16);
    pk.z = v[4] | ((uint32_t)v[5] << 16);
    pk.w = v[6] | ((uint32_t)v[7] << 16);
    *reinterpret_cast<uint4*>(dst + k0) = pk;
  } else {
    for (int q = 0; q < 8 && k0 + q < n; q++) dst[k0 + q] = v[q];
  }
}

// Exhaustive check of the normal-range sqrt / reciprocal sequences (rn_math.cuh) against the
// IEEE intrinsics, over every binary32 bit pattern with biased exponent in [27, 227]
// (2^-100 .. 2^100), both signs skipped for sqrt.  counts[0] = sqrt mismatches, counts[1] = rcp
// mismatches, counts[2] = mismatches of the composed 1/sqrt used by normalise().
__global__ void k_selftest_math(unsigned long long* counts)
{
  const uint64_t first = (uint64_t)27 << 23, last = (uint64_t)228 << 23;
  unsigned long long bad_s = 0, bad_r = 0, bad_c = 0;
  for (uint64_t b = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < last; b += (uint64_t)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((uint32_t)b);
    const float s_ref = __fsqrt_rn(x), s_fast = sqrt_rn_normal(x);
    if (__float_as_uint(s_ref) != __float_as_uint(s_fast)) bad_s++;
    const float r_ref = __frcp_rn(x), r_fast = rcp_rn_normal(x);
    if (__float_as_uint(r_ref) != __float_as_uint(r_fast)) bad_r++;
    const float c_ref = __frcp_rn(s_ref), c_fast = rcp_rn_normal(s_fast);
    if (__float_as_uint(c_ref) != __float_as_uint(c_fast)) bad_c++;
  }
  if (bad_s) atomicAdd(&counts[0], bad_s);
  if (bad_r) atomicAdd(&counts[1], bad_r);
  if (bad_c) atomicAdd(&counts[2], bad_c);
}

}  // namespace

cudaError_t selftest_math(unsigned long long* dev_counts3, cudaStream_t stream, uint64_t* launches)
{
  cudaError_t e = cudaMemsetAsync(dev_counts3, 0, 3 * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  k_selftest_math<<<148 * 8, 256, 0, stream>>>(dev_counts3);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t volcano_apply(Elem* g, size_t pitch, int depth, int i0, int j0, int ni, int nj, float amplitude, cudaStream_t stream,
                          uint64_t* launches)
{
  if (ni <= 0 || nj <= 0) return cudaSuccess;
  // blockIdx.y is limited to 65535 rows: RN_MAX_DEPTH = 14 gives at most 16385 rows.
  dim3 grid((unsigned)((nj + 255) / 256), (unsigned)ni);
  k_volcano<<<grid, 256, 0, stream>>>(g, pitch, 1 << depth, i0, j0, ni, nj, amplitude);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t lcg_draws(uint16_t* dst, uint32_t seed, uint64_t first, uint64_t n, cudaStream_t stream, uint64_t* launches)
{
  if (n == 0) return cudaSuccess;
  const uint64_t threads = (n + 7) / 8;
  k_lcg_draws<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(dst, seed, first, n);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace rn
