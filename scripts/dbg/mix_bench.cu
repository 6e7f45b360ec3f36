// Microbenchmark (sm_100a): what the FP32 pipe really sustains for packed FP32x2 instructions whose operands are
// all DIFFERENT registers (ffma2_bench.cu streams one register against two loop constants), and the throughput
// bound of the instruction mix of one cross-pair trip of sfw_score_crowd (36 FFMA2 + 36 FMUL2 + 16 FADD2 +
// 8 FADD + 16 MUFU + ~40 ALU).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_bench mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define MUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define ADD2(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FMA1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define EX2(d, a) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a))
#define XOR(d, a) asm volatile("xor.b32 %0, %0, %1;" : "+r"(d) : "r"(a))

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
  u64 p[8], q[8], r[8];
  float f[8];
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = threadIdx.x * 0.001f + i;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(v), "f"(v + a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(q[i]) : "f"(v * b), "f"(v - a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(r[i]) : "f"(v + b), "f"(v * a));
    f[i] = v;
    x[i] = threadIdx.x + i;
  }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { // FFMA2, one streaming operand + two loop constants (what ffma2_bench measures)
#pragma unroll
      for (int i = 0; i < 8; ++i) FMA2(p[i], p[i], q[0], r[0]);
    } else if (MODE == 1) { // FFMA2, three different registers
#pragma unroll
      for (int i = 0; i < 8; ++i) FMA2(p[i], q[i], r[i], p[i]);
    } else if (MODE == 2) { // FFMA2, three different registers, none of them the destination
#pragma unroll
      for (int i = 0; i < 8; ++i) FMA2(p[i], q[i], r[i], p[(i + 3) & 7]);
    } else if (MODE == 3) { // FMUL2, two different registers
#pragma unroll
      for (int i = 0; i < 8; ++i) MUL2(p[i], p[i], q[i]);
    } else if (MODE == 4) { // FADD2, two different registers
#pragma unroll
      for (int i = 0; i < 8; ++i) ADD2(p[i], p[i], q[i]);
    } else if (MODE == 5) { // scalar FFMA, three different registers
      float *pf = reinterpret_cast<float *>(p), *qf = reinterpret_cast<float *>(q), *rf = reinterpret_cast<float *>(r);
#pragma unroll
      for (int i = 0; i < 16; ++i) FMA1(pf[i], qf[i], rf[i], pf[i]);
    } else if (MODE == 6 || MODE == 7 || MODE == 8) {
      // the trip's mix, scaled by 1/4: 9 FFMA2 + 9 FMUL2 + 4 FADD2 + 2 FADD + 4 MUFU + 10 ALU, independent chains
      // 7: without the ALU work; 8: without the MUFU
#pragma unroll
      for (int i = 0; i < 8; ++i) FMA2(p[i], q[i], r[i], p[i]);
      FMA2(q[0], p[0], r[1], q[0]);
#pragma unroll
      for (int i = 0; i < 8; ++i) MUL2(r[i], r[i], q[(i + 1) & 7]);
      MUL2(q[1], q[1], p[2]);
#pragma unroll
      for (int i = 0; i < 4; ++i) ADD2(q[2 + i], q[2 + i], p[i]);
      FMA1(f[4], f[4], a, b);
      FMA1(f[5], f[5], a, b);
      if (MODE != 8) {
#pragma unroll
        for (int i = 0; i < 4; ++i) EX2(f[i], f[i]);
      }
      if (MODE != 7) {
#pragma unroll
        for (int i = 0; i < 8; ++i) XOR(x[i], it);
        XOR(x[0], x[1]);
        XOR(x[2], x[3]);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
    s += lo + hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q[i]));
    s += lo + hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i]));
    s += lo + hi + f[i] + (float)x[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double inst_per_iter, int warps_per_smsp) {
  int sms = 148, threads = 256, iters = 20000;
  int blocks = sms * warps_per_smsp / 2; // 8 warps per block = 2 per SMSP
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
  // cycles one SMSP spends per loop iteration of ONE warp = elapsed cycles / (iters * warps per SMSP)
  const double clk = 1.92e9;
  double cyc = ms * 1e-3 * clk / iters / warps_per_smsp;
  printf("%-44s warps/SMSP %d  %8.3f ms  %7.2f cyc per warp-iteration  = %5.2f cyc per instruction (@1.92 GHz)\n", name,
         warps_per_smsp, ms, cyc, cyc / inst_per_iter);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8}) {
    run<0>("FFMA2 stream + 2 constants (8)", 8, w);
    run<1>("FFMA2 3 different regs, dst = src c (8)", 8, w);
    run<2>("FFMA2 3 different regs, dst elsewhere (8)", 8, w);
    run<3>("FMUL2 2 different regs (8)", 8, w);
    run<4>("FADD2 2 different regs (8)", 8, w);
    run<5>("FFMA scalar 3 different regs (16)", 16, w);
    run<6>("trip mix / 4 (9+9+4 packed, 2 FFMA, 4 MUFU, 10 ALU)", 38, w);
    run<7>("trip mix / 4 without the ALU work", 28, w);
    run<8>("trip mix / 4 without the MUFU", 34, w);
  }
  return 0;
}
