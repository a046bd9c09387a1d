// Microbenchmark: do the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2 = add/mul/fma.rn.f32x2) free
// issue slots?  Each kernel performs the same number of FP32 lane operations and the same integer filler work
// per iteration; the scalar form issues 16 FP + 8 INT instructions, the packed form 8 FP2 + 8 INT.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_issue f32x2_issue.cu && ./f32x2_issue
#include <cuda_runtime.h>
#include <cstdio>

typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x; }
__device__ __forceinline__ float hi(u64 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return y; }

template <int INT_FILL>
__global__ void k_scalar(float* out, int iters, float c0, float c1)
{
  float a[8];
  unsigned q[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; q[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = __fmul_rn(__fadd_rn(a[i], c0), c1); }
    if (INT_FILL) {
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = (q[i] ^ (q[i] >> 3)) + 0x9e37u;   // 2 integer instructions... LOP3 + IADD (SHF folded)
    }
  }
  float s = 0; unsigned t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { s += a[i]; t += q[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}

template <int INT_FILL>
__global__ void k_packed(float* out, int iters, float c0, float c1)
{
  u64 a[4];
  unsigned q[8];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = pack(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = threadIdx.x + i;
  const u64 k0 = pack(c0, c0), k1 = pack(c1, c1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = mul2(add2(a[i], k0), k1);
    if (INT_FILL) {
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = (q[i] ^ (q[i] >> 3)) + 0x9e37u;
    }
  }
  float s = 0; unsigned t = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += lo(a[i]) + hi(a[i]);
#pragma unroll
  for (int i = 0; i < 8; i++) t += q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}

template <class K>
float run(K kern, float* out, int iters)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148 * 8, 256>>>(out, iters, 1.0001f, 0.9999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<148 * 8, 256>>>(out, iters, 1.0001f, 0.9999f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main()
{
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  const double laneops = 148.0 * 8 * 256 * iters * 16;
  float a = run(k_scalar<0>, out, iters), b = run(k_packed<0>, out, iters), c = run(k_scalar<1>, out, iters), d = run(k_packed<1>, out, iters);
  printf("FP only      : scalar %.3f ms (%.1f Tlane-op/s)  packed %.3f ms (%.1f Tlane-op/s)\n", a, laneops / a * 1e-9, b, laneops / b * 1e-9);
  printf("FP + integer : scalar %.3f ms  packed %.3f ms  (ratio %.2f)\n", c, d, d / c);
  return 0;
}
