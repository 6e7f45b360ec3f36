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
// Two uses:
//   * scene batch (BASELINE configs[3]): rank q scores its own scenes_per_rank[q] scenes; after the wait every rank
//     holds all of them (sfw_exchange_fetch: rank order = global scene order of a block partition);
//   * row slabs of ONE batch (configs[1]/[4] on several GPUs): every rank stages the same scenes and scores its
//     linvel rows; the wait kernel then MERGES the world records of every scene with the reference's tie-break
//     order on the device (sfw_exchange_merge), so every rank ends up with the full grid's winner.
//
// Buffer: [2 epochs][world][max_scenes] SfwBest (sized for SFW_MAX_RANKS), then arrived[SFW_MAX_RANKS], then
// merged[max_scenes].
// Two epochs suffice: a rank cannot start tick t + 2 before it has seen every peer's record of tick t + 1, and a
// peer's fetch of tick t is stream-ordered before its tick t + 1 kernel.
//
// Failure behaviour: the wait is BOUNDED (sfw_exchange_set_timeout, default 10 s).  If a peer's records do not
// arrive — it died, skipped a tick, or staged fewer scenes than announced — the wait kernel gives up, writes the
// peer's rank into a mapped pinned status word and the next fetch / merge returns SFW_ERR_STATE naming the peer.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>

#include "sfw_ctx.h"
#include "sfw_forces.cuh"

namespace {

struct WaitArgs {
  unsigned int target[SFW_MAX_RANKS];
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Lane q waits for rank q's arrival counter to reach its target (wrap-safe), at most timeout_ns; a lane that gives
// up reports 1 + q through `status` (mapped pinned host memory).  Returns false in every lane if any lane gave up.
__device__ __forceinline__ bool wait_for_peers(const unsigned int *arrived, uint32_t world, const WaitArgs &a,
                                               unsigned long long timeout_ns, unsigned int *status) {
  const uint32_t q = threadIdx.x & 31u;
  bool ok = true;
  if (q < world) {
    const unsigned long long t0 = global_ns();
    for (;;) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(arrived + q) : "memory");
      if ((int)(v - a.target[q]) >= 0)
        break;
      __nanosleep(200);
      if (global_ns() - t0 > timeout_ns) {
        ok = false;
        // a plain store: `status` is mapped pinned HOST memory, where device atomics need PCIe atomics support
        // (compute-sanitizer rejects them); whichever lane's rank lands is a rank that did not deliver
        *reinterpret_cast<volatile unsigned int *>(status) = 1u + q;
        __threadfence_system();
        break;
      }
    }
  }
  return __all_sync(0xffffffffu, ok);
}

__global__ void sfw_exchange_wait_kernel(const unsigned int *arrived, uint32_t world, const WaitArgs a,
                                         unsigned long long timeout_ns, unsigned int *status) {
  wait_for_peers(arrived, world, a, timeout_ns, status);
}

// Row-slab mode: after the wait, merge the `world` slab winners of every scene (total order of the reference's
// sequential best-update, better() in sfw_forces.cuh) into merged[scene].  One block; thread per scene.
__global__ void sfw_exchange_merge_kernel(const unsigned int *arrived, uint32_t world, const WaitArgs a,
                                          unsigned long long timeout_ns, unsigned int *status,
                                          const SfwBest *gathered /* [world][max_scenes] of this epoch */,
                                          uint32_t max_scenes, uint32_t n_scenes, const double *linvels,
                                          const double *angvels, uint32_t n_w, SfwBest *merged) {
  __shared__ int s_ok;
  if (threadIdx.x < 32u) {
    const bool ok = wait_for_peers(arrived, world, a, timeout_ns, status);
    if (threadIdx.x == 0u)
      s_ok = ok ? 1 : 0;
  }
  __syncthreads();
  if (!s_ok)
    return;
  for (uint32_t s = threadIdx.x; s < n_scenes; s += blockDim.x) {
    SfwBest best;
    memset(&best, 0, sizeof(best));
    float bc = -1.f;
    uint32_t bi = 0u;
    for (uint32_t q = 0; q < world; ++q) {
      const SfwBest r = gathered[(size_t)q * max_scenes + s];
      const float c = r.valid ? r.cost : -1.f;
      if (better(c, r.index, bc, bi, linvels, angvels, n_w)) {
        bc = c;
        bi = r.index;
        best = r;
      }
    }
    merged[s] = best;
  }
}

// A rank whose row slab is empty still owes its peers one (invalid) record per scene.
__global__ void sfw_exchange_export_invalid_kernel(const __grid_constant__ SfwExchangeDev X, uint32_t n_scenes) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scenes)
    return;
  SfwBest r;
  memset(&r, 0, sizeof(r));
  export_best(X, s, r);
}

inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

int check_status(sfw_ctx *c, const char *what) {
  if (c->xchg.status && *c->xchg.status) {
    const unsigned int q = *c->xchg.status - 1u;
    *c->xchg.status = 0u;
    return sfw_fail(c, SFW_ERR_STATE,
                    "%s: rank %u did not deliver its winner records within %.1f s (peer dead, a skipped tick, or "
                    "fewer scenes staged than announced with sfw_exchange_expect)",
                    what, q, c->xchg.timeout_s);
  }
  return SFW_OK;
}

WaitArgs wait_args(const sfw_ctx *c) {
  WaitArgs a;
  for (uint32_t q = 0; q < SFW_MAX_RANKS; ++q)
    a.target[q] = (unsigned int)c->xchg.expected[q];
  return a;
}

} // namespace

// called by sfw_run (sfw_abi.cu) for an empty row slab
cudaError_t sfw_launch_export_invalid(const SfwExchangeDev &X, uint32_t n_scenes, cudaStream_t stream) {
  sfw_exchange_export_invalid_kernel<<<(n_scenes + 127u) / 128u, 128, 0, stream>>>(X, n_scenes);
  return cudaGetLastError();
}

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
  c->xchg.off_merged = c->xchg.off_arrived + 256;
  c->xchg.bytes = c->xchg.off_merged + up256((size_t)max_scenes * sizeof(SfwBest));
  SFW_CK(c, cudaMalloc((void **)&c->xchg.local, c->xchg.bytes));
  SFW_CK(c, cudaMemset(c->xchg.local, 0, c->xchg.bytes));
  SFW_CK(c, cudaMallocHost((void **)&c->xchg.host, (size_t)SFW_MAX_RANKS * max_scenes * sizeof(SfwBest)));
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
  c->xchg.counts_set = false;
  return SFW_OK;
}

int sfw_exchange_connect_local(sfw_ctx *const *ctxs, uint32_t world) {
  if (!ctxs || !world || world > SFW_MAX_RANKS)
    return SFW_ERR_ARG;
  for (uint32_t r = 0; r < world; ++r)
    if (!ctxs[r])
      return SFW_ERR_ARG;
  for (uint32_t r = 0; r < world; ++r) {
    sfw_ctx *c = ctxs[r];
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->xchg.exported)
      return sfw_fail(c, SFW_ERR_STATE, "sfw_exchange_connect_local before sfw_exchange_export");
    if (c->xchg.max_scenes != ctxs[0]->xchg.max_scenes)
      return sfw_fail(c, SFW_ERR_ARG, "sfw_exchange_connect_local: max_scenes differs between the contexts");
    SFW_CK(c, cudaSetDevice(c->device));
    for (uint32_t q = 0; q < world; ++q) {
      if (ctxs[q]->device != c->device) {
        int can = 0;
        SFW_CK(c, cudaDeviceCanAccessPeer(&can, c->device, ctxs[q]->device));
        if (!can)
          return sfw_fail(c, SFW_ERR_UNSUPPORTED, "device %d cannot access device %d", c->device, ctxs[q]->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return sfw_fail(c, SFW_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
      }
      c->xchg.peer[q] = ctxs[q]->xchg.local;
    }
    c->xchg.rank = r;
    c->xchg.world = world;
    c->xchg.connected = true;
    c->xchg.local_peers = true;
    c->xchg.counts_set = false;
  }
  return SFW_OK;
}

int sfw_exchange_expect(sfw_ctx *c, const uint32_t *scenes_per_rank) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->xchg.connected)
    return sfw_fail(c, SFW_ERR_STATE, "sfw_exchange_expect before sfw_exchange_connect");
  if (!scenes_per_rank) { // back to "every rank stages what this rank stages"
    c->xchg.counts_set = false;
    return SFW_OK;
  }
  for (uint32_t q = 0; q < c->xchg.world; ++q) {
    if (scenes_per_rank[q] > c->xchg.max_scenes)
      return sfw_fail(c, SFW_ERR_ARG, "sfw_exchange_expect: rank %u announces %u scenes, the gather buffer holds %u", q,
                      scenes_per_rank[q], c->xchg.max_scenes);
    c->xchg.counts[q] = scenes_per_rank[q];
  }
  c->xchg.counts_set = true;
  return SFW_OK;
}

int sfw_exchange_set_timeout(sfw_ctx *c, double seconds) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!(seconds > 0.0) || seconds > 3600.0)
    return sfw_fail(c, SFW_ERR_ARG, "sfw_exchange_set_timeout: need 0 < seconds <= 3600");
  c->xchg.timeout_s = seconds;
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
      reinterpret_cast<const unsigned int *>(c->xchg.local + c->xchg.off_arrived), c->xchg.world, wait_args(c),
      (unsigned long long)(c->xchg.timeout_s * 1e9), c->xchg.status);
  SFW_CK(c, cudaGetLastError());
  c->launches += 1;
  return SFW_OK;
}

int sfw_exchange_fetch(sfw_ctx *c, SfwBest *all_best_out) {
  if (!c)
    return SFW_ERR_ARG;
  int rc = sfw_exchange_sync(c); // the records of the latest run must have ARRIVED before the local buffer is read
  if (rc != SFW_OK)
    return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!all_best_out)
    return sfw_fail(c, SFW_ERR_ARG, "sfw_exchange_fetch: null output");
  const uint32_t ms = c->xchg.max_scenes, world = c->xchg.world;
  const size_t slot = (size_t)((c->xchg.epoch - 1) & 1u);
  const size_t off = slot * world * ms * sizeof(SfwBest), bytes = (size_t)world * ms * sizeof(SfwBest);
  SFW_CK(c, cudaMemcpyAsync(c->xchg.host, c->xchg.local + off, bytes, cudaMemcpyDeviceToHost, c->stream));
  SFW_CK(c, cudaStreamSynchronize(c->stream));
  rc = check_status(c, "sfw_exchange_fetch");
  if (rc != SFW_OK)
    return rc;
  size_t at = 0;
  for (uint32_t q = 0; q < world; ++q) { // rank order = global scene order of a block partition
    const uint32_t n = c->xchg.last_counts[q];
    memcpy(all_best_out + at, c->xchg.host + (size_t)q * ms * sizeof(SfwBest), (size_t)n * sizeof(SfwBest));
    at += n;
  }
  return SFW_OK;
}

int sfw_exchange_merge(sfw_ctx *c, SfwBest *merged_out) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->xchg.connected || !c->xchg.epoch)
    return sfw_fail(c, SFW_ERR_STATE, "sfw_exchange_merge: no exported launch to merge");
  const uint32_t n = c->B.n_scenes, world = c->xchg.world, ms = c->xchg.max_scenes;
  for (uint32_t q = 0; q < world; ++q)
    if (c->xchg.last_counts[q] != n)
      return sfw_fail(c, SFW_ERR_STATE, "sfw_exchange_merge: row-slab mode needs the same %u scenes on every rank (rank %u: %u)",
                      n, q, c->xchg.last_counts[q]);
  SFW_CK(c, cudaSetDevice(c->device));
  const size_t slot = (size_t)((c->xchg.epoch - 1) & 1u);
  const SfwBest *gathered = reinterpret_cast<const SfwBest *>(c->xchg.local) + slot * world * ms;
  SfwBest *merged = reinterpret_cast<SfwBest *>(c->xchg.local + c->xchg.off_merged);
  sfw_exchange_merge_kernel<<<1, 256, 0, c->stream>>>(
      reinterpret_cast<const unsigned int *>(c->xchg.local + c->xchg.off_arrived), world, wait_args(c),
      (unsigned long long)(c->xchg.timeout_s * 1e9), c->xchg.status, gathered, ms, n, c->B.linvels, c->B.angvels, c->B.n_w,
      merged);
  SFW_CK(c, cudaGetLastError());
  c->launches += 1;
  if (!merged_out) // asynchronous use: the merged records stay on the device (sfw_exchange_merged_device)
    return SFW_OK;
  SFW_CK(c, cudaMemcpyAsync(c->xchg.host, merged, (size_t)n * sizeof(SfwBest), cudaMemcpyDeviceToHost, c->stream));
  SFW_CK(c, cudaStreamSynchronize(c->stream));
  const int rc = check_status(c, "sfw_exchange_merge");
  if (rc != SFW_OK)
    return rc;
  memcpy(merged_out, c->xchg.host, (size_t)n * sizeof(SfwBest));
  return SFW_OK;
}

const void *sfw_exchange_device_buffer(sfw_ctx *c) {
  if (!c || !c->xchg.connected || !c->xchg.epoch)
    return nullptr;
  const size_t slot = (size_t)((c->xchg.epoch - 1) & 1u);
  return c->xchg.local + slot * c->xchg.world * c->xchg.max_scenes * sizeof(SfwBest);
}

const void *sfw_exchange_merged_device(sfw_ctx *c) {
  if (!c || !c->xchg.connected)
    return nullptr;
  return c->xchg.local + c->xchg.off_merged;
}

} // extern "C"
