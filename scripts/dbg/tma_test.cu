#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int use_g, int fence, int x, int y, uint32_t bytes, uint8_t* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const CUtensorMap* tp = use_g ? gmap : &tmap;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem)), "l"(tp), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  void *p = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const uint32_t pitch = (argc > 2 ? atoi(argv[2]) : 256), rows = 256; const uint32_t bw = (argc > 3 ? atoi(argv[3]) : 64), bh = (argc > 4 ? atoi(argv[4]) : 32);
  std::vector<uint8_t> h(pitch * rows);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i % pitch) + (i / pitch));
  uint8_t *d, *out; cudaMalloc(&d, h.size()); cudaMalloc(&out, 65536);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap tm; CUresult r; uint32_t bytes = 0;
  int use_g = 0;
  int promo = argc > 5 ? atoi(argv[5]) : 0; int fence = argc > 6 ? atoi(argv[6]) : 0;
  if (variant == 0) { // uint8 64x32
    cuuint64_t dims[2] = {pitch, rows}; cuuint64_t strides[1] = {pitch}; cuuint32_t box[2] = {bw, bh}; cuuint32_t es[2] = {1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); bytes = bw * bh;
  } else if (variant == 1) { // float32 16x32 (same bytes)
    cuuint64_t dims[2] = {pitch / 4, rows}; cuuint64_t strides[1] = {pitch}; cuuint32_t box[2] = {16, 32}; cuuint32_t es[2] = {1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); bytes = 64 * 32;
  } else if (variant == 2) { // uint8, descriptor in global memory
    cuuint64_t dims[2] = {pitch, rows}; cuuint64_t strides[1] = {pitch}; cuuint32_t box[2] = {64, 32}; cuuint32_t es[2] = {1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); bytes = 64 * 32; use_g = 1;
  } else { // uint8 128B swizzle, box 128 x 8
    cuuint64_t dims[2] = {pitch, rows}; cuuint64_t strides[1] = {pitch}; cuuint32_t box[2] = {128, 8}; cuuint32_t es[2] = {1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); bytes = 128 * 8;
  }
  CUtensorMap* gm; cudaMalloc(&gm, sizeof(tm)); cudaMemcpy(gm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  printf("variant %d encode r=%d\n", variant, (int)r);
  int cx = argc > 7 ? atoi(argv[7]) : 64; int cy = argc > 8 ? atoi(argv[8]) : 32;
  k<<<1, 64, 16384>>>(tm, gm, use_g, fence, cx, cy, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d run: %s\n", variant, cudaGetErrorString(e));
  if (e == cudaSuccess) { std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost); printf("o[0..3]=%d %d %d %d (expect %d..)\n", o[0], o[1], o[2], o[3], (cx+cy)&255); }
  return 0;
}
