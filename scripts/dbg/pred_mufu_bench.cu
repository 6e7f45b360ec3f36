// Microbenchmark (sm_100a): does a predicated-OFF MUFU still hold the XU pipe?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pred_mufu_bench pred_mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float *out, int iters, float thr) {
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 0.001f + i * 0.01f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) // unpredicated
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      else // predicated on a run-time test that is false (MODE 1) / true (MODE 2) for every lane
        asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; @p ex2.approx.ftz.f32 %0, %0;}" : "+f"(f[i]) : "f"(MODE == 1 ? thr : -thr));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name) {
  int blocks = 148 * 4, threads = 256, iters = 20000;
  float *out;
  cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 100, 1e30f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 1e30f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * 1.92e9 / iters / 8 /* warps per SMSP */ / 8 /* MUFU per iteration */;
  printf("%-46s %8.3f ms  %6.2f cycles per warp-MUFU per sub-partition\n", name, ms, cyc);
}
int main() {
  run<0>("ex2 unpredicated");
  run<1>("setp + @p ex2, predicate false in every lane");
  run<2>("setp + @p ex2, predicate true in every lane");
  return 0;
}
