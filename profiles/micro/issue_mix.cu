// Micro-benchmark (evidence for DESIGN.md section 5.0): at what rate does one SM sub-partition issue the arithmetic of k_march?
// Register-only loops, 12 one-warp CTAs per SM (3 warps per scheduler, like the output kernel), no memory, no branches inside the
// unrolled body.  Reports warp instructions per cycle and scheduler for: packed FP32 with register operands (FMUL2, FFMA2), scalar
// FADD, integer ALU work, and a mix in the proportions of the kernel's dynamic opcode histogram (profiles/r02_march_lines.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o issue_mix issue_mix.cu && ./issue_mix
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
#define MUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define ADD2(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define FADD(d, a, b) asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define IADD(d, a, b) asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define LOP(d, a, b) asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))

template <int MODE>
__global__ void __launch_bounds__(32) k(float* out, int iters, float c0, float c1, long long* cyc)
{
  u64 p[8], q[8];
  float f[8], g[8];
  int n[8], m[8];
  for (int i = 0; i < 8; i++) {
    float a = threadIdx.x * 0.001f + i, b = 1.0f + 1e-6f * i;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a), "f"(a + 0.5f));
    asm("mov.b64 %0, {%1, %2};" : "=l"(q[i]) : "f"(b), "f"(b));
    f[i] = a; g[i] = c0 * 1e-7f * (i + 1); n[i] = threadIdx.x + i; m[i] = i * 7 + 1;
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (MODE == 0) {  // FMUL2 reg, reg: 16 per round
#pragma unroll
        for (int i = 0; i < 8; i++) { MUL2(p[i], p[i], q[i]); MUL2(p[i], p[i], q[(i + 1) & 7]); }
      } else if (MODE == 1) {  // FFMA2 three registers
#pragma unroll
        for (int i = 0; i < 8; i++) { FMA2(p[i], p[i], q[i], q[(i + 3) & 7]); FMA2(p[i], p[i], q[(i + 1) & 7], q[(i + 5) & 7]); }
      } else if (MODE == 2) {  // scalar FADD
#pragma unroll
        for (int i = 0; i < 8; i++) { FADD(f[i], f[i], g[i]); FADD(f[i], f[i], g[(i + 1) & 7]); }
      } else if (MODE == 3) {  // integer ALU
#pragma unroll
        for (int i = 0; i < 8; i++) { IADD(n[i], n[i], m[i]); LOP(n[i], n[i], m[(i + 1) & 7]); }
      } else if (MODE == 4) {  // the kernel's arithmetic mix per 16 instructions: 4 FADD, 3 FMUL2, 2 FADD2, 2 FFMA2, 5 integer / move class
        FADD(f[0], f[0], g[0]); MUL2(p[0], p[0], q[0]); IADD(n[0], n[0], m[0]); ADD2(p[1], p[1], q[1]);
        FADD(f[1], f[1], g[1]); FMA2(p[2], p[2], q[2], q[3]); LOP(n[1], n[1], m[1]); MUL2(p[3], p[3], q[4]);
        FADD(f[2], f[2], g[2]); IADD(n[2], n[2], m[2]); ADD2(p[4], p[4], q[5]); FMA2(p[5], p[5], q[6], q[7]);
        FADD(f[3], f[3], g[3]); LOP(n[3], n[3], m[3]); MUL2(p[6], p[6], q[0]); IADD(n[4], n[4], m[4]);
      } else {  // MODE 5: packed and integer alternating (can the two pipes overlap?)
#pragma unroll
        for (int i = 0; i < 8; i++) { MUL2(p[i], p[i], q[i]); IADD(n[i], n[i], m[i]); }
      }
    }
  }
  long long t1 = clock64();
  float s = 0; int t = 0;
  for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi + f[i]; t += n[i]; }
  out[blockIdx.x * 32 + threadIdx.x] = s + (float)t + c1;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm)
{
  const int grid = 148 * warps_per_sm, iters = 4000;
  float* out; long long* cyc;
  cudaMalloc(&out, grid * 32 * 4); cudaMalloc(&cyc, grid * 8);
  k<MODE><<<grid, 32>>>(out, iters, 1.0f, 0.0f, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 32>>>(out, iters, 1.0f, 0.0f, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[grid];
  cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; i++) mean += h[i]; mean /= grid;
  const double inst = (double)iters * 4 * 16;  // per warp
  printf("%-52s %2d warps/SM: %.3f warp inst / cycle / scheduler by clock64 (per warp %.2f cycles per inst); %.3f by wall time at 1.965 GHz\n", name,
         warps_per_sm, inst * warps_per_sm / 4.0 / mean, mean / inst, inst * grid / (ms * 1e-3 * 1.965e9 * 148 * 4));
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main()
{
  for (int w : {4, 12, 16}) {
    run<0>("FMUL2 reg,reg", w); run<1>("FFMA2 reg,reg,reg", w); run<2>("FADD reg,reg", w); run<3>("IADD3/LOP3", w);
    run<4>("k_march mix (4 FADD 3 FMUL2 2 FADD2 2 FFMA2 5 int)", w); run<5>("FMUL2 + IADD alternating", w);
  }
  return 0;
}
