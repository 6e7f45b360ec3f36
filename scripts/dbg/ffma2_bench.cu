// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) vs MUFU on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float r;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
template <int MODE> __global__ void k(float *out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  u64 pa, pb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
  u64 p[8];
  unsigned xr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) xr[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma1(acc[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pa, pb);
    } else if (MODE == 2) {  // 8 MUFU.EX2 + 8 FFMA
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
        acc[i] = y;
      }
    } else if (MODE == 4) {  // 8 FFMA2 + 8 scalar FFMA, independent chains: does scalar work ride for free?
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p[i] = fma2(p[i], pa, pb);
        acc[i] = fma1(acc[i], a, b);
      }
    } else if (MODE == 5) {  // 8 FFMA2 + 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p[i] = fma2(p[i], pa, pb);
        acc[2 * i] = fma1(acc[2 * i], a, b);
        acc[2 * i + 1] = fma1(acc[2 * i + 1], a, b);
      }
    } else if (MODE == 6) {  // 8 FFMA2 + 4 MUFU
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pa, pb);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
        acc[i] = y;
      }
    } else if (MODE == 7) {  // 8 FFMA2 + 8 integer LOP3/IADD (ALU pipe)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p[i] = fma2(p[i], pa, pb);
        unsigned u = __float_as_uint(acc[i]);
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(u) : "r"(it));
        acc[i] = __uint_as_float(u);
      }
    } else if (MODE >= 8 && MODE <= 12) {
      // 8: 16 FFMA + 8 XOR   9: 8 FFMA2 + 4 XOR   10: 8 FFMA2 + 16 XOR   11: 16 XOR   12: 16 FFMA + 16 XOR
      constexpr int NX = (MODE == 8) ? 8 : (MODE == 9) ? 4 : 16;
      if (MODE == 8 || MODE == 12) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma1(acc[i], a, b);
      } else if (MODE != 11) {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pa, pb);
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(xr[i]) : "r"(it));
      }
    } else if (MODE >= 13 && MODE <= 16) {  // 8 MUFU of one kind: 13 rsqrt, 14 sqrt, 15 rcp, 16 rsqrt+ex2 alternating
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y;
        if (MODE == 13 || (MODE == 16 && (i & 1)))
          asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
        else if (MODE == 14)
          asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
        else if (MODE == 15)
          asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
        else
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
        acc[i] = y;
      }
    } else if (MODE == 3) {  // mix: 1 MUFU per 4 FFMA2
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pa, pb);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
        acc[i] = y;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
    s += lo + hi;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) s += (float)xr[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double ops_per_iter_per_thread) {
  int sms = 148, blocks = sms * 4, threads = 256, iters = 20000;
  float *out;
  cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ops = ops_per_iter_per_thread * iters * (double)blocks * threads;
  printf("%-28s %8.3f ms  %8.2f Gop/s  = %6.1f op/clk/SM @1.965GHz\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("FFMA (16 per iter)", 16);
  run<1>("FFMA2 (8 per iter, 16 fma)", 16);
  run<2>("MUFU.EX2 (8 per iter)", 8);
  run<3>("8 FFMA2 + 2 MUFU", 18);
  run<4>("8 FFMA2 + 8 FFMA", 24);
  run<5>("8 FFMA2 + 16 FFMA", 32);
  run<6>("8 FFMA2 + 4 MUFU", 20);
  run<7>("8 FFMA2 + 8 XOR", 24);
  run<8>("16 FFMA + 8 XOR", 24);
  run<9>("8 FFMA2 + 4 XOR", 20);
  run<10>("8 FFMA2 + 16 XOR", 32);
  run<11>("16 XOR", 16);
  run<12>("16 FFMA + 16 XOR", 32);
  run<13>("MUFU.RSQ (8 per iter)", 8);
  run<14>("MUFU.SQRT (8 per iter)", 8);
  run<15>("MUFU.RCP (8 per iter)", 8);
  run<16>("MUFU.RSQ/EX2 alternating (8)", 8);
  return 0;
}
