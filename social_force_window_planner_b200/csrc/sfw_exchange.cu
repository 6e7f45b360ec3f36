// sfw_exchange.cu — multi-GPU winner exchange fused into the scorer's epilogue.
//
// The only cross-rank step of the path is "every rank learns every scene's winning (v, w)" (DESIGN.md 5).
// Instead of a collective AFTER the kernel, the block that reduces a scene's winner stores the 32-byte record
// straight into every rank's gather buffer over NVLink (export_best, sfw_forces.cuh) and bumps an arrival
// counter with system-scope release semantics; a rank that needs the gathered records enqueues a one-warp wait
// kernel on its stream (acquire on its own counters) and reads its LOCAL buffer.  One process per GPU: the
// buffers are shared through cudaIpc handles, which the caller moves between the processes (the Python side
// uses torch.distributed.all_gather_object for that, once, at set-up).
//
// Buffer: [2 epochs][world][max_scenes] SfwBest, then arrived[SFW_MAX_RANKS].  Two epochs suffice: a rank
// cannot start tick t + 2 before it has seen every peer's record of tick t + 1, and a peer's fetch of tick t is
// stream-ordered before its tick t + 1 kernel.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>

#include "sfw_ctx.h"

namespace {

__global__ void sfw_exchange_wait_kernel(const unsigned int *arrived, uint32_t world, unsigned int target) {
  const uint32_t q = threadIdx.x;
  if (q < world) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(arrived + q) : "memory");
      if ((int)(v - target) < 0)
        __nanosleep(200);
    } while ((int)(v - target) < 0); // wrap-safe
  }
}

inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

} // namespace

extern "C" {

int sfw_exchange_export(sfw_ctx *c, uint32_t max_scenes, void *handle_out) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!handle_out || !max_scenes)
    return sfw_fail(c, SFW_ERR_ARG, "sfw_exchange_export: null handle / zero scenes");
  if (c->xchg.exported)
    return sfw_fail(c, SFW_ERR_STATE, "sfw_exchange_export: already exported");
  SFW_CK(c, cudaSetDevice(c->device));
  c->xchg.max_scenes = max_scenes;
  c->xchg.off_arrived = up256(2ull * SFW_MAX_RANKS * max_scenes * sizeof(SfwBest));
  c->xchg.bytes = c->xchg.off_arrived + 256;
  SFW_CK(c, cudaMalloc((void **)&c->xchg.local, c->xchg.bytes));
  SFW_CK(c, cudaMemset(c->xchg.local, 0, c->xchg.bytes));
  SFW_CK(c, cudaMallocHost((void **)&c->xchg.host, c->xchg.off_arrived));
  cudaIpcMemHandle_t h;
  SFW_CK(c, cudaIpcGetMemHandle(&h, c->xchg.local));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, sizeof(h));
  c->xchg.exported = true;
  return SFW_OK;
}

int sfw_exchange_connect(sfw_ctx *c, uint32_t rank, uint32_t world, const void *handles) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->xchg.exported)
    return sfw_fail(c, SFW_ERR_STATE, "sfw_exchange_connect before sfw_exchange_export");
  if (!handles || !world || world > SFW_MAX_RANKS || rank >= world)
    return sfw_fail(c, SFW_ERR_ARG, "sfw_exchange_connect: rank %u / world %u (max %d)", rank, world, SFW_MAX_RANKS);
  SFW_CK(c, cudaSetDevice(c->device));
  for (uint32_t q = 0; q < world; ++q) {
    if (q == rank) {
      c->xchg.peer[q] = c->xchg.local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t *)handles + 64 * (size_t)q, sizeof(h));
    SFW_CK(c, cudaIpcOpenMemHandle(&c->xchg.peer[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  c->xchg.rank = rank;
  c->xchg.world = world;
  c->xchg.connected = true;
  return SFW_OK;
}

int sfw_exchange_sync(sfw_ctx *c) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->xchg.connected || !c->xchg.epoch)
    return sfw_fail(c, SFW_ERR_STATE, "sfw_exchange_sync: no exported launch to wait for");
  SFW_CK(c, cudaSetDevice(c->device));
  sfw_exchange_wait_kernel<<<1, 32, 0, c->stream>>>(
      reinterpret_cast<const unsigned int *>(c->xchg.local + c->xchg.off_arrived), c->xchg.world,
      (unsigned int)c->xchg.expected);
  SFW_CK(c, cudaGetLastError());
  c->launches += 1;
  return SFW_OK;
}

int sfw_exchange_fetch(sfw_ctx *c, SfwBest *all_best_out) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->xchg.connected || !c->xchg.epoch || !all_best_out)
    return sfw_fail(c, SFW_ERR_STATE, "sfw_exchange_fetch: nothing exchanged yet");
  SFW_CK(c, cudaSetDevice(c->device));
  const uint32_t n = c->B.n_scenes, ms = c->xchg.max_scenes, world = c->xchg.world;
  const size_t slot = (size_t)((c->xchg.epoch - 1) & 1u);
  const size_t off = slot * world * ms * sizeof(SfwBest), bytes = (size_t)world * ms * sizeof(SfwBest);
  SFW_CK(c, cudaMemcpyAsync(c->xchg.host, c->xchg.local + off, bytes, cudaMemcpyDeviceToHost, c->stream));
  SFW_CK(c, cudaStreamSynchronize(c->stream));
  for (uint32_t q = 0; q < world; ++q)
    memcpy(all_best_out + (size_t)q * n, c->xchg.host + (size_t)q * ms * sizeof(SfwBest), (size_t)n * sizeof(SfwBest));
  return SFW_OK;
}

const void *sfw_exchange_device_buffer(sfw_ctx *c) {
  if (!c || !c->xchg.connected || !c->xchg.epoch)
    return nullptr;
  const size_t slot = (size_t)((c->xchg.epoch - 1) & 1u);
  return c->xchg.local + slot * c->xchg.world * c->xchg.max_scenes * sizeof(SfwBest);
}

} // extern "C"
