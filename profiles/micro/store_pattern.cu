// Micro-benchmark (evidence for DESIGN.md): HBM write bandwidth of 16-byte-per-lane stores
//   A: fully coalesced (lane l writes element base+l)                   -- 512 contiguous bytes per warp instruction
//   B: half-sector interleave: one instruction writes the even elements of a 64-element run, a second
//      instruction (same thread, back to back) the odd elements         -- 16 B at 32 B stride per instruction
//   C: like B but the odd elements are written by a second pass of the CTA after __syncthreads (~us later)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void kA(float4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_float4(i, 1, 2, 3);
}
__global__ void kB(float4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * i + 1 < n; i += (size_t)gridDim.x * blockDim.x) {
    p[2 * i] = make_float4(i, 1, 2, 3);
    p[2 * i + 1] = make_float4(i, 4, 5, 6);
  }
}
// tile-like: each CTA owns a 64x64 tile of a 4096-wide image; pass 1 writes even columns of all rows, barrier, pass 2 odd columns
__global__ void kC(float4* p, int W) {
  const int tx = blockIdx.x % (W / 64), ty = blockIdx.x / (W / 64);
  float4* base = p + (size_t)ty * 64 * W + tx * 64;
  for (int pass = 0; pass < 2; pass++) {
    for (int idx = threadIdx.x; idx < 64 * 32; idx += blockDim.x) {
      const int r = idx / 32, c = (idx % 32) * 2 + pass;
      base[(size_t)r * W + c] = make_float4(idx, pass, 2, 3);
    }
    __syncthreads();
  }
}
__global__ void kD(float4* p, int W) {  // tile-like, fully coalesced rows
  const int tx = blockIdx.x % (W / 64), ty = blockIdx.x / (W / 64);
  float4* base = p + (size_t)ty * 64 * W + tx * 64;
  for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) base[(size_t)(idx / 64) * W + (idx % 64)] = make_float4(idx, 0, 2, 3);
}
int main() {
  const int W = 4096;
  const size_t n = (size_t)W * W;
  float4* p;
  cudaMalloc(&p, n * 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 4; which++) {
    float best = 1e9f;
    for (int rep = 0; rep < 12; rep++) {
      cudaEventRecord(e0);
      if (which == 0) kA<<<148 * 8, 256>>>(p, n);
      if (which == 1) kB<<<148 * 8, 256>>>(p, n);
      if (which == 2) kC<<<(W / 64) * (W / 64), 512>>>(p, W);
      if (which == 3) kD<<<(W / 64) * (W / 64), 512>>>(p, W);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    printf("pattern %c: %.1f us  %.0f GB/s\n", "ABCD"[which], best * 1e3, n * 16 / (best * 1e-3) / 1e9);
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
