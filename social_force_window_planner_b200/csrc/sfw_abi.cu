// sfw_abi.cu — host side of the C ABI declared in include/sfw_b200.h.
//
// Packs caller scenes (FP64, AoS, reference-shaped: sfm::Agent / ControllerParams / Costmap2D) into
// the device layout of sfw_dev.h inside ONE pinned arena, ships it with ONE async H2D copy on the
// context stream, launches the scorer and brings back SfwBest (+ the cost vector on request).
// There is no CPU scoring path in this file: every result comes from the kernels.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "sfw_ctx.h"
#include "sfw_dev.h"
#include "sfw_kernels.h"

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
typedef SfwPlan Plan;
typedef SfwArena Arena;

} // namespace

std::string &sfw_create_error() {
  thread_local std::string e;
  return e;
}

int sfw_arena_reserve(sfw_ctx *c, SfwArena &a, size_t bytes) {
  if (bytes <= a.cap)
    return SFW_OK;
  // the streams may still be reading the old buffers
  SFW_CK(c, cudaStreamSynchronize(c->stream));
  if (c->copy_stream)
    SFW_CK(c, cudaStreamSynchronize(c->copy_stream));
  if (a.host)
    cudaFreeHost(a.host);
  if (a.dev)
    cudaFree(a.dev);
  a.host = nullptr;
  a.dev = nullptr;
  a.cap = 0;
  size_t cap = align_up(bytes + bytes / 4, 1 << 16);
  SFW_CK(c, cudaMallocHost((void **)&a.host, cap));
  SFW_CK(c, cudaMalloc((void **)&a.dev, cap));
  a.cap = cap;
  return SFW_OK;
}

namespace {

#define fail sfw_fail
#define CK SFW_CK
#define arena_reserve sfw_arena_reserve

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 3-D uint8 tensor [scene][row][col] over the costmap slots; box = one scene's window.
int make_tensor_map(sfw_ctx *c, const uint8_t *maps, uint32_t pitch, uint32_t rows, uint32_t scenes,
                    uint32_t wp, uint32_t wh) {
  if (c->tm_ptr == maps && c->tm_pitch == pitch && c->tm_rows == rows && c->tm_scenes == scenes &&
      c->tm_wp == wp && c->tm_h == wh)
    return SFW_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc)
    return fail(c, SFW_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {pitch, rows, scenes};
  cuuint64_t strides[2] = {pitch, (cuuint64_t)pitch * rows};
  cuuint32_t box[3] = {wp, wh, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)maps, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(c, SFW_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  c->tm_ptr = maps;
  c->tm_pitch = pitch;
  c->tm_rows = rows;
  c->tm_scenes = scenes;
  c->tm_wp = wp;
  c->tm_h = wh;
  return SFW_OK;
}

const SfwSfmParams kDefaultSfm = {2.0, 10.0, 0.2, 2.1, 3.0, 2.0, 1.0, 2.0, 0.35, 2.0, 3.0, 0.5};

// one run = run_prepare (once) + run_launch per range of scenes (defined next to sfw_run)
struct RunState {
  uint32_t rb = 0, re = 0;
  bool slab_share = false;
};
int run_prepare(sfw_ctx *c, RunState &rs);
int run_launch(sfw_ctx *c, const RunState &rs, uint32_t s0, uint32_t cnt);

// Choose block size / tiling for the thread-per-trajectory kernel: maximise resident threads per
// SM under the shared-memory and register limits, then shrink the block so a single-wave launch
// is spread evenly over all SMs.
// family: -1 = choose the kernel family here; 0 / 1 = keep thread-per-trajectory / block-per-trajectory (a row
// slab must be scored by the kernel its full grid was planned for, so that sharding never changes a bit);
// crowd_threads: likewise the block size of the block-per-trajectory kernel (its two instantiations sum the
// per-warp accumulator rows in different groupings), 0 = choose.
int make_plan(sfw_ctx *c, uint32_t n_scenes, uint32_t samples, uint32_t P, uint32_t M, uint32_t F,
              uint32_t win_wp, uint32_t win_h, int steps, int family = -1, uint32_t crowd_threads = 0) {
  Plan &pl = c->plan;
  if (pl.valid && pl.n_scenes == n_scenes && pl.samples == samples && pl.maxP == P && pl.maxM == M &&
      pl.maxF == F && pl.win_wp == win_wp && pl.win_h == win_h && pl.steps == steps)
    return SFW_OK;
  uint32_t bestT = 0, bestK = 0;
  size_t max_dyn = 0;
  CK(c, sfw_small_max_dynamic_smem(&max_dyn));
  // Thread-per-trajectory needs a whole crowd in one thread's shared-memory column; it stops paying
  // (and its 64-bit goal mask stops fitting) beyond SFW_MAX_PEDS_SMALL.  Denser crowds go to the
  // block-per-trajectory kernel.
  bool try_small = P <= SFW_MAX_PEDS_SMALL;
  if (try_small && sfw_small_smem_bytes(win_wp, win_h, P, M, F, 128) > max_dyn)
    try_small = false; // fewer than 4 warps per SM would fit
  if (family == 1 || (family < 0 && c->policy == SFW_POLICY_LATENCY))
    try_small = false;
  // Block-per-trajectory kernel: 256-thread blocks (two per SM), or — a small crowd on a grid of a few waves —
  // 128-thread blocks (four per SM) when that saves waves.  Measured end to end on one wave of 5 x 9 samples
  // (profiles/r2n_block_size_variants.txt): a 128-thread block is 1.06 - 1.09 x slower per trajectory up to 5 pedestrians and
  // 1.25 - 1.32 x at 20 - 40; BASELINE's 21 x 21 / 5-pedestrian tick goes from two waves to one: 0.100 -> 0.066 ms.
  uint32_t crowd_T = crowd_threads ? crowd_threads : (uint32_t)SFW_CROWD_THREADS;
  size_t crowd_smem = sfw_crowd_smem_bytes(P, M, F, (uint32_t)steps, crowd_T);
  int crowd_k = 0;
  double crowd_waves = 0.0; // waves of the chosen block size, in units of a 256-thread wave's time
  const uint64_t crowd_total = (uint64_t)n_scenes * samples;
  if (crowd_smem <= max_dyn) {
    CK(c, sfw_crowd_prepare(crowd_T, crowd_smem, &crowd_k));
    if (crowd_k > 0) {
      const uint64_t slots = (uint64_t)c->sm_count * crowd_k;
      const uint64_t waves = (crowd_total + slots - 1) / slots;
      crowd_waves = (double)waves;
      if (!crowd_threads && P <= SFW_MAX_PEDS_SMALL && waves >= 2 && waves <= 8) {
        const size_t smem_s = sfw_crowd_smem_bytes(P, M, F, (uint32_t)steps, SFW_CROWD_THREADS_SMALL);
        int k_s = 0;
        CK(c, sfw_crowd_prepare(SFW_CROWD_THREADS_SMALL, smem_s, &k_s));
        if (k_s > 0) {
          const uint64_t slots_s = (uint64_t)c->sm_count * k_s;
          const double waves_s = (double)((crowd_total + slots_s - 1) / slots_s) * (P <= 6 ? 1.1 : 1.3);
          if (waves_s < crowd_waves) {
            crowd_T = SFW_CROWD_THREADS_SMALL;
            crowd_smem = smem_s;
            crowd_k = k_s;
            crowd_waves = waves_s;
          }
        }
      }
    }
  }
  if (try_small && family < 0 && c->policy == SFW_POLICY_AUTO && crowd_smem <= max_dyn) {
    // Small grids are latency bound in the thread-per-trajectory kernel: one thread walks every pair of a
    // trajectory, so a tick costs what ONE warp costs however few trajectories there are.  The block-per-trajectory
    // kernel spreads a trajectory's pairs (and obstacle clusters) over threads and runs one block per trajectory.
    // Measured per rollout step, end to end through sfw_score (scripts/latency_policy_probe.py, latency_probe.py;
    // 0 ... 40 pedestrians): thread per trajectory 2.1 + 0.63 P + 0.016 P^2 us, block per trajectory
    // 1.65 + 0.085 P us per WAVE of blocks.  A single wave always wins (116 vs 135 us without pedestrians,
    // 132 vs 188 us with one, 40 steps); two waves from 3 pedestrians on; at 20 pedestrians up to 8 waves.
    if (crowd_k > 0) {
      const double Pd = (double)P;
      // (from 4 pedestrian pairs on the block-per-trajectory kernel spreads its force phase over all warps:
      // 2.0 + 0.027 P us per step and wave, 155 us end to end at 20 pedestrians, 170 us at 40)
      const double wave_us = P >= 7 ? 2.0 + 0.027 * Pd : 1.65 + 0.085 * Pd;
      if (crowd_waves * wave_us < 2.1 + 0.63 * Pd + 0.016 * Pd * Pd)
        try_small = false;
    }
  }
  if (!try_small) {
    const size_t smem = crowd_smem;
    const int k = crowd_k;
    if (smem > max_dyn)
      return fail(c, SFW_ERR_UNSUPPORTED,
                  "scene does not fit the block-per-trajectory kernel (peds=%u steps=%d: %zu B of shared memory)",
                  P, steps, smem);
    if (k <= 0)
      return fail(c, SFW_ERR_UNSUPPORTED, "block-per-trajectory kernel cannot be resident (peds=%u)", P);
    const uint64_t total = (uint64_t)n_scenes * samples;
    pl.n_scenes = n_scenes;
    pl.samples = samples;
    pl.maxP = P;
    pl.maxM = M;
    pl.maxF = F;
    pl.win_wp = win_wp;
    pl.win_h = win_h;
    pl.steps = steps;
    pl.crowd = true;
    pl.grid = (uint32_t)std::min<uint64_t>(total, (uint64_t)c->sm_count * k);
    pl.T = crowd_T;
    pl.tiles = 1;
    pl.smem = smem;
    pl.valid = true;
    return SFW_OK;
  }
  pl.crowd = false;
  pl.steps = steps;
  for (uint32_t T = SFW_MAX_BLOCK_SMALL; T >= 32; T -= 32) {
    size_t smem = sfw_small_smem_bytes(win_wp, win_h, P, M, F, T);
    if (smem > max_dyn)
      continue;
    int k = 0;
    CK(c, sfw_small_occupancy(T, smem, &k));
    if (k <= 0)
      continue;
    if ((uint64_t)k * T > (uint64_t)bestK * bestT) {
      bestK = (uint32_t)k;
      bestT = T;
    }
  }
  if (!bestT)
    return fail(c, SFW_ERR_UNSUPPORTED,
                "scene does not fit the thread-per-trajectory kernel (peds=%u obstacles=%u)", P, M);
  const uint64_t total = (uint64_t)n_scenes * samples;
  const uint64_t resident = (uint64_t)c->sm_count * bestK * bestT;
  uint32_t T = bestT;
  if (total <= resident) {
    // single wave: spread evenly, keep the same number of blocks per SM
    uint64_t per_block = (total + (uint64_t)c->sm_count * bestK - 1) / ((uint64_t)c->sm_count * bestK);
    T = (uint32_t)std::min<uint64_t>(bestT, std::max<uint64_t>(32, align_up(per_block, 32)));
  } else {
    const uint64_t waves = (total + resident - 1) / resident;
    uint64_t per_block = (total + waves * c->sm_count * bestK - 1) / (waves * c->sm_count * bestK);
    T = (uint32_t)std::min<uint64_t>(bestT, std::max<uint64_t>(32, align_up(per_block, 32)));
  }
  if (T > samples)
    T = (uint32_t)std::max<size_t>(32, align_up(samples, 32));
  pl.n_scenes = n_scenes;
  pl.samples = samples;
  pl.maxP = P;
  pl.maxM = M;
  pl.maxF = F;
  pl.win_wp = win_wp;
  pl.win_h = win_h;
  pl.T = T;
  pl.k = bestK;
  pl.tiles = (samples + T - 1) / T;
  pl.smem = sfw_small_smem_bytes(win_wp, win_h, P, M, F, T);
  pl.valid = true;
  return SFW_OK;
}


// Obstacle points -> clusters of 8 behind a {centre, reach^2} header (sfw_dev.h, obstacle_sum2).  `pts` are in
// the kernel's units (metres * log2(e)/sigma, scene frame).  Clusters come from recursive median splits along
// the wider axis, cut at multiples of 8, ties broken by the other coordinate and the input position, so the
// layout only depends on the points.  reach = cluster radius + r_max (largest agent radius, same units) +
// cutoff: beyond it every term exp2(-(|p - o| - r)) of the cluster is below 2^-cutoff.  cutoff <= 0: reach = inf.
void split_obstacles(std::vector<uint32_t> &idx, const std::vector<float2> &pts, size_t lo, size_t hi) {
  const size_t n = hi - lo;
  if (n <= SFW_OBST_CLUSTER)
    return;
  float x0 = pts[idx[lo]].x, x1 = x0, y0 = pts[idx[lo]].y, y1 = y0;
  for (size_t i = lo; i < hi; ++i) {
    const float2 p = pts[idx[i]];
    x0 = std::min(x0, p.x), x1 = std::max(x1, p.x), y0 = std::min(y0, p.y), y1 = std::max(y1, p.y);
  }
  const bool along_x = (x1 - x0) >= (y1 - y0);
  // NaN coordinates (a caller's problem, but not a reason for an invalid comparator) sort last
  auto key = [](float v) { return v == v ? v : std::numeric_limits<float>::infinity(); };
  std::sort(idx.begin() + lo, idx.begin() + hi, [&](uint32_t a, uint32_t b) {
    const float2 p = pts[a], q = pts[b];
    const float pa = key(along_x ? p.x : p.y), qa = key(along_x ? q.x : q.y);
    if (pa != qa)
      return pa < qa;
    const float pb = key(along_x ? p.y : p.x), qb = key(along_x ? q.y : q.x);
    if (pb != qb)
      return pb < qb;
    return a < b;
  });
  const size_t clusters = (n + SFW_OBST_CLUSTER - 1) / SFW_OBST_CLUSTER;
  const size_t mid = lo + SFW_OBST_CLUSTER * ((clusters + 1) / 2);
  split_obstacles(idx, pts, lo, mid);
  split_obstacles(idx, pts, mid, hi);
}

void pack_obstacles(const std::vector<float2> &pts, float2 *out, double cutoff_log2, double r_max,
                    std::vector<uint32_t> &idx) {
  const size_t n = pts.size();
  idx.resize(n);
  for (size_t i = 0; i < n; ++i)
    idx[i] = (uint32_t)i;
  split_obstacles(idx, pts, 0, n);
  for (size_t lo = 0, g = 0; lo < n; lo += SFW_OBST_CLUSTER, ++g) {
    const size_t hi = std::min(n, lo + SFW_OBST_CLUSTER);
    float2 *rec = out + g * SFW_OBST_CLUSTER_SLOTS;
    float x0 = pts[idx[lo]].x, x1 = x0, y0 = pts[idx[lo]].y, y1 = y0;
    for (size_t i = lo; i < hi; ++i) {
      const float2 p = pts[idx[i]];
      x0 = std::min(x0, p.x), x1 = std::max(x1, p.x), y0 = std::min(y0, p.y), y1 = std::max(y1, p.y);
    }
    const float cx = 0.5f * (x0 + x1), cy = 0.5f * (y0 + y1);
    double rad = 0.0;
    for (size_t i = lo; i < hi; ++i)
      rad = std::max(rad, std::hypot((double)pts[idx[i]].x - cx, (double)pts[idx[i]].y - cy));
    float reach2 = std::numeric_limits<float>::infinity();
    if (cutoff_log2 > 0.0) {
      const double reach = (rad + r_max + cutoff_log2) * (1.0 + 1e-6); // rounding of the device's FP32 test
      const double r2 = reach * reach;
      reach2 = r2 < 3.0e38 ? (float)r2 : std::numeric_limits<float>::infinity();
    }
    rec[0] = make_float2(cx, cy);
    rec[1] = make_float2(reach2, 0.f);
    for (size_t i = 0; i < SFW_OBST_CLUSTER; ++i)
      rec[2 + i] = (lo + i < hi) ? pts[idx[lo + i]] : make_float2(SFW_FAR_AWAY, 0.f);
  }
}
} // namespace

extern "C" {

int sfw_abi_version(void) { return SFW_ABI_VERSION; }

void sfw_default_params(SfwParams *p) {
  if (!p)
    return;
  memset(p, 0, sizeof(*p));
  p->max_vel_x = 0.7;
  p->max_trans_acc = 1.0;
  p->max_rot_acc = 1.0;
  p->sim_time = 1.0;
  p->sim_granularity = 0.025;
  p->robot_radius = 0.35f;
  p->social_weight = 1.2;
  p->costmap_weight = 2.0;
  p->angle_weight = 0.7;
  p->distance_weight = 1.0;
  p->vel_weight = 1.0;
}

void sfw_default_sfm_params(SfwSfmParams *p) {
  if (p)
    *p = kDefaultSfm;
}

const char *sfw_last_error(const sfw_ctx *ctx) { return ctx ? ctx->err.c_str() : sfw_create_error().c_str(); }

int sfw_create(sfw_ctx **out, int device, void *stream, const SfwLimits *limits) {
  if (!out)
    return fail(nullptr, SFW_ERR_ARG, "sfw_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, SFW_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n)
    return fail(nullptr, SFW_ERR_ARG, "device %d out of range (have %d)", device, n);
  e = cudaSetDevice(device);
  if (e != cudaSuccess)
    return fail(nullptr, SFW_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess)
    return fail(nullptr, SFW_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major < 10)
    return fail(nullptr, SFW_ERR_UNSUPPORTED,
                "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  sfw_ctx *c = new sfw_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  {
    // host workers for batches: a fair share of the cores when one process per GPU runs on every visible device
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    c->host_threads = (int)std::min(16u, std::max(1u, hw / (unsigned)std::max(n, 1)));
  }
  memset(&c->B, 0, sizeof(c->B));
  memset(&c->tmap, 0, sizeof(c->tmap));
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete c;
      return fail(nullptr, SFW_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    c->own_stream = true;
  }
  e = cudaHostAlloc((void **)&c->status, 64, cudaHostAllocMapped);
  if (e == cudaSuccess)
    e = cudaEventCreateWithFlags(&c->h2d_done, cudaEventDisableTiming);
  if (e == cudaSuccess)
    e = cudaEventCreateWithFlags(&c->ev_compute, cudaEventDisableTiming);
  if (e == cudaSuccess)
    e = cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming);
  if (e == cudaSuccess)
    e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    sfw_fail(nullptr, SFW_ERR_CUDA, "sfw_create: %s", cudaGetErrorString(e));
    sfw_destroy(c);
    return SFW_ERR_CUDA;
  }
  memset(c->status, 0, 64);
  c->xchg.status = c->status + 1;
  c->pdl = getenv("SFW_B200_NO_PDL") == nullptr;
  if (limits) {
    size_t in_est = (size_t)limits->max_scenes *
                        (sizeof(SfwSceneDev) + (size_t)limits->max_peds * 48 +
                         (size_t)sfw_obst_slots(limits->max_obstacles) * 8 + 1024 + align_up(limits->max_cells, 256)) +
                    65536;
    size_t out_est = (size_t)limits->max_scenes * (sizeof(SfwBest) + (size_t)limits->max_samples * 6 + 4096);
    if (arena_reserve(c, c->in, in_est) != SFW_OK || arena_reserve(c, c->out, out_est) != SFW_OK) {
      sfw_create_error() = c->err;
      sfw_destroy(c);
      return SFW_ERR_CUDA;
    }
  }
  *out = c;
  return SFW_OK;
}

int sfw_destroy(sfw_ctx *c) {
  if (!c)
    return SFW_OK;
  cudaSetDevice(c->device);
  if (c->stream)
    cudaStreamSynchronize(c->stream);
  if (c->in.host)
    cudaFreeHost(c->in.host);
  if (c->in.dev)
    cudaFree(c->in.dev);
  if (c->out.host)
    cudaFreeHost(c->out.host);
  if (c->out.dev)
    cudaFree(c->out.dev);
  for (SfwArena *a : {&c->sensor_in, &c->sensor_out}) {
    if (a->host)
      cudaFreeHost(a->host);
    if (a->dev)
      cudaFree(a->dev);
  }
  for (uint32_t q = 0; q < c->xchg.world && c->xchg.connected && !c->xchg.local_peers; ++q)
    if (q != c->xchg.rank && c->xchg.peer[q])
      cudaIpcCloseMemHandle(c->xchg.peer[q]);
  if (c->share_buf)
    cudaFree(c->share_buf);
  if (c->xchg.local)
    cudaFree(c->xchg.local);
  if (c->xchg.host)
    cudaFreeHost(c->xchg.host);
  if (c->status)
    cudaFreeHost(c->status);
  if (c->h2d_done)
    cudaEventDestroy(c->h2d_done);
  if (c->ev_compute)
    cudaEventDestroy(c->ev_compute);
  if (c->ev_copy)
    cudaEventDestroy(c->ev_copy);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  if (c->d_points)
    cudaFree(c->d_points);
  if (c->h_points)
    cudaFreeHost(c->h_points);
  if (c->own_stream && c->stream)
    cudaStreamDestroy(c->stream);
  delete c;
  return SFW_OK;
}

static int upload_impl(sfw_ctx *c, const SfwParams *params, const SfwSfmParams *sfm_in, const SfwScene *scenes,
                       uint32_t n_scenes, const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                       bool launch_too);
// sfw_upload proper.  launch_too (sfw_score_batch): a big batch of the thread-per-trajectory family is LAUNCHED
// piece by piece as its costmaps land on the device — the scorer works on the first scenes while the host workers
// are still packing the last ones — and the call returns with the run already enqueued (c->ran).
extern "C++" int upload_impl(sfw_ctx *c, const SfwParams *params, const SfwSfmParams *sfm_in, const SfwScene *scenes,
                       uint32_t n_scenes, const double *linvels, uint32_t n_v, const double *angvels,
                       uint32_t n_w, bool launch_too) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!params || !scenes || !n_scenes || !linvels || !angvels || !n_v || !n_w)
    return fail(c, SFW_ERR_ARG, "sfw_upload: null/empty argument");
  if (!(params->sim_granularity > 0.0) || !(params->sim_time >= 0.0) || !(params->max_vel_x != 0.0))
    return fail(c, SFW_ERR_ARG, "sfw_upload: sim_granularity must be > 0, sim_time >= 0, max_vel_x != 0");
  if ((uint64_t)n_v * n_w > 0x7fffffffull)
    return fail(c, SFW_ERR_ARG, "sfw_upload: too many samples");
  CK(c, cudaSetDevice(c->device));
  const SfwSfmParams &sfm = sfm_in ? *sfm_in : kDefaultSfm;
  if (!(sfm.force_sigma_obstacle > 0.0) || !(sfm.gamma > 0.0) || !(sfm.relaxation_time > 0.0))
    return fail(c, SFW_ERR_ARG, "sfm parameters: force_sigma_obstacle, gamma and relaxation_time must be > 0");
  c->staged = false;
  c->ran = false;
  SfwScratch &X = c->scratch;
  // a batch of scenes is packed by the context's host workers (one scene = one item); a lone scene — the control
  // tick — stays on the calling thread
  if (n_scenes >= 8 && c->pool.size() < (unsigned)c->host_threads)
    c->pool.resize((unsigned)c->host_threads);
  if (X.w.size() < c->pool.size())
    X.w.resize(c->pool.size());

  // ---- shape scan: sizes and per-scene prefix offsets ---------------------------------------------
  uint32_t maxP = 0, maxM = 0, maxF = 0, max_sx = 0, max_sy = 0;
  X.off_pairs.assign(n_scenes + 1, 0u);
  X.off_obst.assign(n_scenes + 1, 0u);
  X.off_fp.assign(n_scenes + 1, 0u);
  X.off_grp.assign(n_scenes + 1, 0u);
  X.off_ped.assign(n_scenes + 1, 0u);
  X.grp_cnt.assign(n_scenes, 0u);
  bool any_groups = false;
  for (uint32_t s = 0; s < n_scenes; ++s) {
    const SfwScene &sc = scenes[s];
    if (!sc.costmap || !sc.size_x || !sc.size_y || !(sc.resolution > 0.0))
      return fail(c, SFW_ERR_ARG, "scene %u: costmap missing or empty", s);
    if ((sc.n_peds && !sc.peds) || (sc.n_obstacles && !sc.obstacles_xy) ||
        (sc.n_footprint && !sc.footprint_xy))
      return fail(c, SFW_ERR_ARG, "scene %u: null array with non-zero count", s);
    if (sc.n_peds > SFW_MAX_PEDS_CROWD)
      return fail(c, SFW_ERR_UNSUPPORTED, "scene %u: %u pedestrians > %d supported by this build", s,
                  sc.n_peds, SFW_MAX_PEDS_CROWD);
    if (sc.n_footprint > SFW_MAX_FOOTPRINT)
      return fail(c, SFW_ERR_UNSUPPORTED, "scene %u: footprint with %u vertices > %d", s,
                  sc.n_footprint, SFW_MAX_FOOTPRINT);
    maxP = std::max(maxP, sc.n_peds);
    maxM = std::max(maxM, sc.n_obstacles);
    maxF = std::max(maxF, sc.n_footprint);
    max_sx = std::max(max_sx, sc.size_x);
    max_sy = std::max(max_sy, sc.size_y);
    uint32_t tagged = 0;
    for (uint32_t j = 0; j < sc.n_peds; ++j)
      tagged += sc.peds[j].group_id >= 0 ? 1u : 0u;
    any_groups = any_groups || tagged >= 2u;
    X.off_pairs[s + 1] = X.off_pairs[s] + (sc.n_peds + 1) / 2;         // pedestrians are stored as pairs
    X.off_obst[s + 1] = X.off_obst[s] + sfw_obst_slots(sc.n_obstacles); // clusters of 8 behind a 2-slot header
    X.off_fp[s + 1] = X.off_fp[s] + sc.n_footprint;
    X.off_ped[s + 1] = X.off_ped[s] + sc.n_peds;
    // group table of a scene: start[0 .. G] + 2 words per member; at most tagged / 2 groups of >= 2 members
    X.off_grp[s + 1] = X.off_grp[s] + (tagged >= 2u ? tagged / 2u + 1u + 2u * tagged : 0u);
  }
  const uint64_t totP = X.off_pairs[n_scenes], totM = X.off_obst[n_scenes], totF = X.off_fp[n_scenes];
  const uint64_t totG = X.off_grp[n_scenes];

  // ---- rollout constants --------------------------------------------------------------------
  int num_steps = (int)(params->sim_time / params->sim_granularity + 0.5); // sfw_planner.cpp:519
  if (num_steps == 0)
    num_steps = 1;
  if (num_steps > 65535)
    return fail(c, SFW_ERR_UNSUPPORTED, "num_steps %d > 65535", num_steps);
  const double dt = params->sim_time / num_steps; // :527

  // ---- staged window: everything the footprint can touch ------------------------------------
  double max_lin = 0.0;
  for (uint32_t i = 0; i < n_v; ++i)
    max_lin = std::max(max_lin, std::fabs(linvels[i]));
  uint32_t win_wp = 0, win_h = 0;
  X.wx0.assign(n_scenes, 0);
  X.wy0.assign(n_scenes, 0);
  {
    int64_t need = 0;
    bool ok = true;
    for (uint32_t s = 0; s < n_scenes && ok; ++s) {
      const SfwScene &sc = scenes[s];
      double circ = 0.0;
      for (uint32_t k = 0; k < sc.n_footprint; ++k)
        circ = std::max(circ, std::hypot(sc.footprint_xy[2 * k], sc.footprint_xy[2 * k + 1]));
      const double speed = std::hypot(std::max(max_lin, std::fabs(sc.robot.vx)), sc.robot.vy);
      const double reach = speed * params->sim_time + circ;
      const double rc_d = std::ceil(reach / sc.resolution) + 2.0;
      if (!(rc_d < 4096.0)) {
        ok = false;
        break;
      }
      const int64_t rc = (int64_t)rc_d;
      const double cxd = std::floor((sc.robot.x - sc.origin_x) / sc.resolution);
      const double cyd = std::floor((sc.robot.y - sc.origin_y) / sc.resolution);
      if (!(std::fabs(cxd) < 1e9) || !(std::fabs(cyd) < 1e9)) {
        ok = false;
        break;
      }
      // TMA (measured on B200): the innermost start coordinate times the element size must be a
      // multiple of 16 bytes or the copy traps as an illegal instruction -> floor x0 to 16 cells
      // and widen the box by the slack.
      const int64_t x0 = (int64_t)cxd - rc;
      const int64_t x0a = (x0 >= 0) ? (x0 / 16) * 16 : -(((-x0) + 15) / 16) * 16;
      X.wx0[s] = (int32_t)x0a;
      X.wy0[s] = (int32_t)((int64_t)cyd - rc);
      need = std::max<int64_t>(need, 2 * rc + 1);
    }
    if (ok && need > 0) {
      const uint32_t wp = (uint32_t)align_up((size_t)need + 15, 16);
      if (wp <= 256 && need <= 256 && (size_t)wp * need <= 48 * 1024) {
        win_wp = wp;
        win_h = (uint32_t)need;
      }
    }
  }

  // ---- launch plan (block size, kernel family): needed by the sharing decision below ----------
  const uint32_t samples = n_v * n_w;
  {
    const int prc = make_plan(c, n_scenes, samples, maxP, maxM, maxF, win_wp, win_h, num_steps);
    if (prc != SFW_OK)
      return prc;
  }

  // ---- rollout prefix sharing: leading saturated updates of every row / column (SfwShareDev) ----
  // Replays the scalar recurrences of step_velocity exactly as the kernel evaluates them (IEEE double add /
  // compare): update k of row r is "saturated" iff the ramp vi +- a*dt does not reach the target yet.
  const bool share_candidate = c->share_allowed && c->policy != SFW_POLICY_LATENCY && (uint64_t)n_v * n_w >= 1024 &&
                               num_steps >= 8;
  uint32_t sh_kmax = 0;
  double sh_mean_s0 = 0.0;
  X.kv.clear();
  X.kw.clear();
  X.dv.clear();
  X.dw.clear();
  X.perm.clear();
  X.rperm.clear();
  X.lvl_rows.clear();
  X.lvl_cols.clear();
  X.chunk_map.clear();
  if (share_candidate) {
    const int kcap = std::min(num_steps - 1, 96);
    const double ax_dt = params->max_trans_acc * dt, ath_dt = params->max_rot_acc * dt;
    X.kv.resize((size_t)n_scenes * n_v);
    X.dv.resize((size_t)n_scenes * n_v);
    X.kw.resize((size_t)n_scenes * n_w);
    X.dw.resize((size_t)n_scenes * n_w);
    for (SfwScratch::Worker &w : X.w) {
      w.kmax = 0;
      w.tot = 0.0;
      w.up.resize(kcap + 1);
      w.dn.resize(kcap + 1);
    }
    // the saturated ramps are the same for every row / column of a scene: build them once per scene (the
    // additions are the kernel's own, sequentially rounded), then count how far each target lets them run
    c->pool.run(n_scenes, [&](uint32_t s, unsigned wi) {
      SfwScratch::Worker &W = X.w[wi];
      std::vector<double> &up = W.up, &dn = W.dn;
      auto build = [&](double v0, double a_dt) {
        up[0] = dn[0] = v0;
        for (int k = 1; k <= kcap; ++k) {
          up[k] = up[k - 1] + a_dt;
          dn[k] = dn[k - 1] - a_dt;
        }
      };
      auto ramp = [&](double target, double v0, double a_dt, uint16_t &k_out, uint8_t &dir_out) {
        const bool rising = (target - v0) >= 0.0;
        int k = 0;
        if (a_dt > 0.0) {
          if (rising)
            while (k < kcap && target >= up[k + 1])
              ++k;
          else
            while (k < kcap && target <= dn[k + 1])
              ++k;
        }
        k_out = (uint16_t)k;
        dir_out = rising ? 1 : 0;
        W.kmax = std::max<uint32_t>(W.kmax, (uint32_t)k);
      };
      build(scenes[s].robot.vx, ax_dt);
      for (uint32_t r = 0; r < n_v; ++r)
        ramp(linvels[r], scenes[s].robot.vx, ax_dt, X.kv[(size_t)s * n_v + r], X.dv[(size_t)s * n_v + r]);
      build(scenes[s].robot.vtheta, ath_dt);
      for (uint32_t q = 0; q < n_w; ++q)
        ramp(angvels[q], scenes[s].robot.vtheta, ath_dt, X.kw[(size_t)s * n_w + q], X.dw[(size_t)s * n_w + q]);
    }, 64);
    for (const SfwScratch::Worker &w : X.w)
      sh_kmax = std::max(sh_kmax, w.kmax);
    // per scene: rows sorted by kv, columns by kw (counting sorts), how many fork before each step, and the mean
    // fork step max(kv, kw) over the grid (from the prefix counts)
    const uint32_t L = sh_kmax + 2u;
    X.perm.resize((size_t)n_scenes * n_w);
    X.rperm.resize((size_t)n_scenes * n_v);
    X.lvl_rows.assign((size_t)n_scenes * L, 0u);
    X.lvl_cols.assign((size_t)n_scenes * L, 0u);
    c->pool.run(n_scenes, [&](uint32_t s, unsigned wi) {
      SfwScratch::Worker &W = X.w[wi];
      W.cur.resize(L);
      const uint16_t *kvs = X.kv.data() + (size_t)s * n_v, *kws = X.kw.data() + (size_t)s * n_w;
      uint32_t *lr = X.lvl_rows.data() + (size_t)s * L, *lc = X.lvl_cols.data() + (size_t)s * L;
      for (uint32_t r = 0; r < n_v; ++r)
        ++lr[kvs[r] + 1u];
      for (uint32_t q = 0; q < n_w; ++q)
        ++lc[kws[q] + 1u];
      for (uint32_t k = 1; k < L; ++k) {
        lr[k] += lr[k - 1u];
        lc[k] += lc[k - 1u];
      }
      for (uint32_t k = 1; k <= sh_kmax; ++k) // sum of max(kv, kw) = sum over k >= 1 of #{max >= k}
        W.tot += (double)n_v * n_w - (double)lr[k] * (double)lc[k];
      std::copy(lr, lr + L, W.cur.begin());
      for (uint32_t r = 0; r < n_v; ++r)
        X.rperm[(size_t)s * n_v + W.cur[kvs[r]]++] = r;
      std::copy(lc, lc + L, W.cur.begin());
      for (uint32_t q = 0; q < n_w; ++q)
        X.perm[(size_t)s * n_w + W.cur[kws[q]]++] = q;
    }, 64);
    double tot = 0.0;
    for (const SfwScratch::Worker &w : X.w)
      tot += w.tot;
    sh_mean_s0 = tot / ((double)n_scenes * n_v * n_w);
  }
  // ---- worth it?  Sharing removes the first max(kv, kw) steps of every sample and costs two latency-bound
  // path launches of kmax steps each.  Cost model from scripts/share_probe.py / latency_probe.py (us).
  bool share_on = false, share_dealt = false, share_merged = false;
  uint32_t share_warp = 0; // bit 0 / 1: launch 1 / 2 uses the warp-per-path writer
  uint32_t share_rec = 0;
  size_t share_need = 0;
  const uint32_t share_paths = 4u + 2u * n_w + 2u * n_v;
  if (share_candidate && sh_kmax >= 2) {
    const bool crowd = c->plan.crowd;
    const double P = (double)maxP, S = (double)num_steps;
    const double total = (double)n_scenes * samples;
    const uint32_t P2max = (maxP + 1u) / 2u;
    const double resident = crowd ? (double)c->plan.grid : (double)c->sm_count * c->plan.k * c->plan.T;
    // thread-per-trajectory: throughput cost of a trajectory-step / one step of a lone warp;
    // block-per-trajectory: a trajectory-step at full occupancy / one step of one block
    const double t_ts_ns = crowd ? 0.2 + 0.00085 * P * P : 0.1 + 0.017 * P + 0.0008 * P * P;
    const double t_lat_us = crowd ? 2.0 + 0.00025 * P * P : 1.3 + 0.25 * P + 0.028 * P * P;
    const bool single = !crowd && total <= resident;
    const double mean_s0 = sh_mean_s0;
    double saved_us, margin;
    if (single) {
      // one wave: T = c1 (N + 1.5) per step with N warps per scheduler and c1 = t_lat / 2.5 (profiles/README.md);
      // the plain launch ends with its fullest scheduler, the dealt one (below) with the average
      const double N = (double)c->plan.T * c->plan.k / 128.0, c1 = t_lat_us / 2.5;
      saved_us = c1 * (S * (std::ceil(N) + 1.5) - (S - mean_s0) * (N + 1.5));
      margin = 1.5;
    } else {
      saved_us = total * mean_s0 * t_ts_ns * 1e-3;
      margin = crowd ? 2.0 : 1.5;
    }
    // the writers (thread-per-trajectory family): a thread per path (step = 2 + 0.94 P + 0.0027 P^2 us, measured
    // 11.7 / 23 us at 10 / 21 pedestrians) or, without pedestrian groups, a warp per path (3 + 0.16 P us, and
    // ~20 % more per extra warp on the scheduler); chosen per launch
    double cost_us;
    if (crowd) {
      const double path_waves = std::ceil((double)n_scenes * share_paths / resident);
      cost_us = (1.0 + path_waves) * sh_kmax * t_lat_us + 30.0;
    } else {
      const double t_thr = 2.0 + 0.94 * P + 0.0027 * P * P, t_warp = 3.0 + 0.16 * P;
      const size_t smw = sfw_small_smem_bytes(win_wp, win_h, maxP, maxM, maxF, SFW_PATH_WARP_THREADS);
      const double per_sm = std::max(1.0, std::min(16.0, std::floor(227.0 * 1024.0 / (double)(smw + 1024))));
      const double slots = c->sm_count * per_sm;
      auto warp_cost = [&](double blocks) {
        const double waves = std::ceil(blocks / slots), wps = std::min(per_sm, std::ceil(blocks / c->sm_count));
        return sh_kmax * t_warp * waves * (1.0 + 0.2 * (wps - 1.0)) + 20.0;
      };
      auto thread_cost = [&](double blocks) { return sh_kmax * t_thr * std::ceil(blocks / slots) + 20.0; };
      const double paths2 = (double)(share_paths - 4u);
      double c1 = thread_cost((double)n_scenes), c2 = thread_cost((double)n_scenes * std::ceil(paths2 / 128.0));
      if (!any_groups) {
        const double w1 = warp_cost((double)n_scenes);
        const double w2 = warp_cost((double)n_scenes * std::ceil(paths2 / (SFW_PATH_WARP_THREADS / 32.0)));
        if (w1 < c1) {
          c1 = w1;
          share_warp |= 1u;
        }
        if (w2 < c2) {
          c2 = w2;
          share_warp |= 2u;
        }
      }
      cost_us = c1 + c2 + 10.0;
      if (share_warp == 3u) {
        // both stages in one launch (the second waits for the first's records): only when every block is resident
        // at once, so that no assumption about the dispatch order is needed
        int occ = 0;
        CK(c, sfw_warp_paths_occupancy(smw, &occ));
        const double blocks = (double)n_scenes * (1.0 + std::ceil(paths2 / (SFW_PATH_WARP_THREADS / 32.0)));
        share_merged = blocks <= (double)c->sm_count * occ;
      }
    }
    share_rec = crowd ? (uint32_t)align_up(16u + 32u * P2max + 2u * P2max, 16)
                      : (uint32_t)align_up(sizeof(SfwCkptHdr) + 32u * P2max, 16);
    share_need = (size_t)n_scenes * share_paths * (sh_kmax + 1u) * share_rec;
    share_on = (saved_us > margin * cost_us || c->share_allowed == 2) && share_need <= ((size_t)8 << 30);
    share_dealt = share_on && !crowd && total < 2.5 * resident; // few waves: every block gets the same mix
    if (getenv("SFW_B200_TRACE_SHARING"))
      fprintf(stderr,
              "sfw_b200 sharing: %s kernel, %u scenes x %u samples, P=%u S=%d kmax=%u mean fork step %.2f, %s, "
              "saved %.0f us vs cost %.0f us (%s writers) -> %s\n",
              crowd ? "block-per-trajectory" : "thread-per-trajectory", n_scenes, samples, maxP, num_steps, sh_kmax,
              mean_s0, single ? "one wave" : "several waves", saved_us, cost_us,
              share_warp == 3u ? (share_merged ? "warp, one launch" : "warp") : share_warp == 1u ? "warp + thread" : share_warp == 2u ? "thread + warp" : "thread",
              share_on ? "on" : "off");
    if (share_dealt) {
      // One wave: a block-sorted order would leave the blocks of unshortened samples as long as before.  Deal
      // the fork-sorted 32-sample chunks over (block, scheduler) bins instead -- warp w of a block issues from
      // scheduler w % 4 -- longest chunks to the bins that hold the fewest warps (448 threads = 14 warps sit
      // 4/4/3/3 on the schedulers), boustrophedon within a class of equal capacity so that the sums level out.
      const uint32_t T = c->plan.T, wpb = T / 32u, nchunk = (samples + 31u) / 32u, nblk = c->plan.tiles;
      struct Bin {
        uint32_t cap, blk, sch;
      };
      std::vector<Bin> bins;
      bins.reserve((size_t)nblk * 4u);
      for (uint32_t b = 0; b < nblk; ++b) {
        const uint32_t warps = std::min(wpb, nchunk - std::min(nchunk, b * wpb));
        for (uint32_t q = 0; q < 4u; ++q)
          if (q < warps)
            bins.push_back({(warps - q + 3u) / 4u, b, q});
      }
      std::stable_sort(bins.begin(), bins.end(), [](const Bin &a, const Bin &b) { return a.cap < b.cap; });
      X.chunk_map.assign((size_t)nblk * wpb, 0u);
      uint32_t ch = 0; // chunks in ascending fork step = longest first
      for (size_t b0 = 0; b0 < bins.size();) {
        size_t b1 = b0;
        while (b1 < bins.size() && bins[b1].cap == bins[b0].cap)
          ++b1;
        const size_t nb = b1 - b0;
        for (uint32_t slot = 0; slot < bins[b0].cap; ++slot)
          for (size_t i = 0; i < nb; ++i, ++ch) {
            const Bin &bn = bins[b0 + ((slot & 1u) ? nb - 1 - i : i)];
            X.chunk_map[(size_t)bn.blk * wpb + bn.sch + 4u * slot] = ch;
          }
        b0 = b1;
      }
    }
  }

  // ---- arena layout --------------------------------------------------------------------------
  const uint32_t map_pitch = (uint32_t)align_up(max_sx, 16);
  const uint32_t map_rows = max_sy;
  const size_t slot = (size_t)map_pitch * map_rows;
  size_t off = 0;
  const size_t o_scenes = off;
  off = align_up(off + sizeof(SfwSceneDev) * n_scenes, kAlign);
  const size_t o_pos = off;
  off = align_up(off + 16 * totP, kAlign);
  const size_t o_vel = off;
  off = align_up(off + 16 * totP, kAlign);
  const size_t o_goal = off;
  off = align_up(off + 16 * totP, kAlign);
  const size_t o_par = off;
  off = align_up(off + 16 * totP, kAlign);
  const size_t o_par2 = off;
  off = align_up(off + 16 * totP, kAlign);
  const size_t o_gbits = off;
  off = align_up(off + 2 * totP, kAlign);
  const size_t o_obs = off;
  off = align_up(off + 8 * totM, kAlign);
  const size_t o_fp = off;
  off = align_up(off + 16 * totF, kAlign);
  const size_t o_grp = off;
  off = align_up(off + 4 * totG, kAlign);
  const size_t o_lin = off;
  off = align_up(off + 8 * (size_t)n_v, kAlign);
  const size_t o_ang = off;
  off = align_up(off + 8 * (size_t)n_w, kAlign);
  const size_t o_skv = off;
  off = align_up(off + 2 * X.kv.size(), kAlign);
  const size_t o_skw = off;
  off = align_up(off + 2 * X.kw.size(), kAlign);
  const size_t o_sdv = off;
  off = align_up(off + X.dv.size(), kAlign);
  const size_t o_sdw = off;
  off = align_up(off + X.dw.size(), kAlign);
  const size_t o_sperm = off;
  off = align_up(off + 4 * X.perm.size(), kAlign);
  const size_t o_srperm = off;
  off = align_up(off + 4 * X.rperm.size(), kAlign);
  const size_t o_slr = off;
  off = align_up(off + 4 * X.lvl_rows.size(), kAlign);
  const size_t o_slc = off;
  off = align_up(off + 4 * X.lvl_cols.size(), kAlign);
  const size_t o_scm = off;
  off = align_up(off + 4 * X.chunk_map.size(), kAlign);
  const size_t o_maps = off;
  off = align_up(off + slot * n_scenes, kAlign);
  const size_t in_bytes = off;
  int rc = arena_reserve(c, c->in, in_bytes);
  if (rc != SFW_OK)
    return rc;

  // ---- outputs -------------------------------------------------------------------------------
  const double inv_sigma = 1.0 / sfm.force_sigma_obstacle;
  const uint32_t max_tiles = (samples + 31) / 32; // any slab / block size fits
  size_t oo = 0;
  c->off_best = oo;
  oo = align_up(oo + sizeof(SfwBest) * n_scenes, kAlign);
  c->off_costs = oo;
  oo = align_up(oo + 4 * (size_t)samples * n_scenes, kAlign);
  c->off_npts = oo;
  oo = align_up(oo + 2 * (size_t)samples * n_scenes, kAlign);
  c->off_bb = oo;
  oo = align_up(oo + sizeof(SfwBlockBest) * (size_t)max_tiles * n_scenes, kAlign);
  c->off_cnt = oo;
  oo = align_up(oo + 4 * (size_t)n_scenes, kAlign);
  c->off_work = oo;
  oo = align_up(oo + 64, kAlign);
  rc = arena_reserve(c, c->out, oo);
  if (rc != SFW_OK)
    return rc;
  // The tile counters must start at 0 (they self-reset after every launch).  Their offset moves with
  // the sample count, so clear them on every upload: 4 bytes per scene on the same stream.
  // (the work / done counters of the block-per-trajectory kernel, right behind them, re-arm themselves too)
  CK(c, cudaMemsetAsync(c->out.dev + c->off_cnt, 0, c->off_work + 64 - c->off_cnt, c->stream));
  c->out_scenes = n_scenes;
  c->out_samples = samples;

  // ---- batch descriptor ------------------------------------------------------------------------
  SfwBatchDev &B = c->B;
  memset(&B, 0, sizeof(B));
  uint8_t *dv = c->in.dev;
  B.scenes = reinterpret_cast<const SfwSceneDev *>(dv + o_scenes);
  B.pedPos = reinterpret_cast<const float4 *>(dv + o_pos);
  B.pedVel = reinterpret_cast<const float4 *>(dv + o_vel);
  B.pedGoal = reinterpret_cast<const float4 *>(dv + o_goal);
  B.pedPar = reinterpret_cast<const float4 *>(dv + o_par);
  B.pedPar2 = reinterpret_cast<const float4 *>(dv + o_par2);
  B.goal_bits = dv + o_gbits;
  B.obst = reinterpret_cast<const float2 *>(dv + o_obs);
  B.footprint = reinterpret_cast<const double2 *>(dv + o_fp);
  B.groups = reinterpret_cast<const uint32_t *>(dv + o_grp);
  B.maps = dv + o_maps;
  B.linvels = reinterpret_cast<const double *>(dv + o_lin);
  B.angvels = reinterpret_cast<const double *>(dv + o_ang);
  B.costs = reinterpret_cast<float *>(c->out.dev + c->off_costs);
  B.npts = reinterpret_cast<uint16_t *>(c->out.dev + c->off_npts);
  B.best = reinterpret_cast<SfwBest *>(c->out.dev + c->off_best);
  // small results (a control tick's cost vector and winner) also land in the pinned host buffer directly
  c->zero_copy_out = (sizeof(SfwBest) + 4 * (size_t)samples) * n_scenes <= ((size_t)1 << 20);
  B.costs_host = c->zero_copy_out ? reinterpret_cast<float *>(c->out.host + c->off_costs) : nullptr;
  B.best_host = c->zero_copy_out ? reinterpret_cast<SfwBest *>(c->out.host + c->off_best) : nullptr;
  B.blockbest = reinterpret_cast<SfwBlockBest *>(c->out.dev + c->off_bb);
  B.counters = reinterpret_cast<unsigned int *>(c->out.dev + c->off_cnt);
  B.map_pitch = map_pitch;
  B.map_rows = map_rows;
  B.n_scenes = n_scenes;
  B.n_v = n_v;
  B.n_w = n_w;
  B.row_begin = 0;
  B.row_end = n_v;
  B.tiles_per_scene = c->plan.tiles;
  B.win_wp = win_wp;
  B.win_h = win_h;
  B.num_steps = num_steps;
  B.score_zero = c->score_zero ? 1u : 0u;
  B.status = c->status;
  B.dt = dt;
  B.max_vel_x = params->max_vel_x;
  B.acc_x = params->max_trans_acc;
  B.acc_th = params->max_rot_acc;
  B.w_vel = params->vel_weight;
  B.w_dist = params->distance_weight;
  B.w_ang = params->angle_weight;
  B.w_map = params->costmap_weight;
  B.w_soc = params->social_weight;
  B.rr2 = params->robot_radius * params->robot_radius; // float product (sfw_planner.cpp:617)
  const double log2e = 1.4426950408889634;
  B.lambda = (float)sfm.lambda;
  B.gamma = (float)sfm.gamma;
  B.c_d = (float)(log2e / sfm.gamma);
  B.c_np = (float)(sfm.n_prime * sfm.n_prime * log2e);
  B.c_n = (float)(sfm.n * sfm.n * log2e);
  B.k_soc = (float)sfm.force_factor_social;
  B.kd_tau = (float)(sfm.force_factor_desired / sfm.relaxation_time);
  B.inv_tau = (float)(1.0 / sfm.relaxation_time);
  B.c_obs = (float)(log2e * inv_sigma);
  B.dtf = (float)dt;
  B.k_gaze = (float)sfm.force_factor_group_gaze;
  B.k_coh = (float)sfm.force_factor_group_coherence;
  B.k_rep = (float)sfm.force_factor_group_repulsion;
  // ---- prefix sharing decided above: records + tables ----------------------------------------
  c->share_active = false;
  c->share_warp = 0;
  if (share_on) {
    if (share_need > c->share_cap) {
      CK(c, cudaStreamSynchronize(c->stream));
      if (c->share_buf)
        cudaFree(c->share_buf);
      c->share_buf = nullptr;
      c->share_cap = 0;
      CK(c, cudaMalloc((void **)&c->share_buf, share_need));
      CK(c, cudaMemsetAsync(c->share_buf, 0, share_need, c->stream)); // record flags (SfwCkptHdr::epoch) start clear
      c->share_cap = share_need;
    }
    B.share.records = c->share_buf;
    B.share.kv = reinterpret_cast<const uint16_t *>(dv + o_skv);
    B.share.kw = reinterpret_cast<const uint16_t *>(dv + o_skw);
    B.share.dirv = dv + o_sdv;
    B.share.dirw = dv + o_sdw;
    B.share.col_perm = reinterpret_cast<const uint32_t *>(dv + o_sperm);
    B.share.row_perm = reinterpret_cast<const uint32_t *>(dv + o_srperm);
    B.share.lvl_rows = reinterpret_cast<const uint32_t *>(dv + o_slr);
    B.share.lvl_cols = reinterpret_cast<const uint32_t *>(dv + o_slc);
    B.share.chunk_map = X.chunk_map.empty() ? nullptr : reinterpret_cast<const uint32_t *>(dv + o_scm);
    B.share.scene_stride = (uint64_t)share_paths * (sh_kmax + 1u) * share_rec;
    B.share.rec_bytes = share_rec;
    B.share.kmax = sh_kmax;
    B.share.mode = 0;
    c->share_active = true;
    c->share_warp = share_warp;
    c->share_merged = share_merged;
    c->share_paths = share_paths;
    c->share_mean_s0 = sh_mean_s0;
    c->share_off_rperm = o_srperm;
    c->share_off_lvl_rows = o_slr;
    c->share_kmax = sh_kmax;
    c->share_rows_begin = 0;
    c->share_rows_end = n_v;
  }
  if (win_wp) {
    rc = make_tensor_map(c, B.maps, map_pitch, map_rows, n_scenes, win_wp, win_h);
    if (rc != SFW_OK)
      return rc;
  }

  // ---- pack ----------------------------------------------------------------------------------
  // sfw_upload is asynchronous: the previous call's H2D copy may still be reading the pinned staging buffer
  // (upload(A); run; upload(B) without a sync in between) — wait for that copy, not for the whole stream
  CK(c, cudaEventSynchronize(c->h2d_done));
  uint8_t *h = c->in.host;
  SfwSceneDev *hs = reinterpret_cast<SfwSceneDev *>(h + o_scenes);
  float4 *hPos = reinterpret_cast<float4 *>(h + o_pos);
  float4 *hVel = reinterpret_cast<float4 *>(h + o_vel);
  float4 *hGoal = reinterpret_cast<float4 *>(h + o_goal);
  float4 *hPar = reinterpret_cast<float4 *>(h + o_par);
  float4 *hPar2 = reinterpret_cast<float4 *>(h + o_par2);
  uint8_t *hBits = h + o_gbits;
  float2 *hO = reinterpret_cast<float2 *>(h + o_obs);
  double2 *hF = reinterpret_cast<double2 *>(h + o_fp);
  uint32_t *hG = reinterpret_cast<uint32_t *>(h + o_grp);
  if (share_candidate && !X.chunk_map.empty())
    memcpy(h + o_scm, X.chunk_map.data(), 4 * X.chunk_map.size());
  memcpy(h + o_lin, linvels, 8 * (size_t)n_v);
  memcpy(h + o_ang, angvels, 8 * (size_t)n_w);
  const double c_obs_d = (double)(float)(1.4426950408889634 * inv_sigma); // the float the kernel multiplies by
  const uint32_t L = sh_kmax + 2u;
  for (SfwScratch::Worker &w : X.w)
    w.cull_skipped = w.cull_tests = 0;

  // One scene, start to finish: obstacle clusters, the pedestrian order that goes with them, group table, scene
  // record, pedestrian pairs, footprint, costmap slot, its slices of the sharing tables.  Scenes write disjoint
  // ranges of the arena, so any number of workers may run this at once.
  auto pack_scene = [&](uint32_t s, unsigned wi) {
    SfwScratch::Worker &W = X.w[wi];
    const SfwScene &sc = scenes[s];
    const SfwRobot &R = sc.robot;
    const uint32_t pP = X.off_pairs[s], pM = X.off_obst[s], pF = X.off_fp[s];
    // -- obstacle clusters + pedestrian order.  Obstacle points are stored as compact clusters a pedestrian PAIR
    // skips when both members are out of reach (obstacle_sum2).  Pedestrians are therefore packed in the order of
    // the clusters they reach from their start positions, so that the two members of a pair (and neighbouring
    // pairs: the lanes of the block-per-trajectory kernel) skip the same clusters.  Any order is a valid one: the
    // social force sums over all others.
    float r_max = (float)R.agent_radius;
    for (uint32_t j = 0; j < sc.n_peds; ++j)
      r_max = std::max(r_max, (float)sc.peds[j].radius);
    W.pts.resize(sc.n_obstacles);
    for (uint32_t k = 0; k < sc.n_obstacles; ++k)
      W.pts[k] = make_float2((float)((sc.obstacles_xy[2 * k] - R.x) * c_obs_d),
                             (float)((sc.obstacles_xy[2 * k + 1] - R.y) * c_obs_d));
    float2 *rec = hO + pM;
    pack_obstacles(W.pts, rec, c->obst_cutoff_log2, (double)r_max * c_obs_d, W.idx);
    const uint32_t slots = sfw_obst_slots(sc.n_obstacles), n_cl = slots / SFW_OBST_CLUSTER_SLOTS;
    W.order.resize(sc.n_peds);
    W.slot.resize(sc.n_peds);
    uint32_t *order = W.order.data();
    for (uint32_t j = 0; j < sc.n_peds; ++j)
      order[j] = j;
    if (c->obst_cutoff_log2 > 0.0 && n_cl && sc.n_peds) {
      const uint32_t words = (n_cl + 63u) / 64u;
      std::vector<uint64_t> &reach = W.reach; // per pedestrian: bit g = cluster g within reach
      std::vector<uint32_t> &reach_n = W.reach_n;
      reach.assign((size_t)sc.n_peds * words, 0ull);
      reach_n.assign(sc.n_peds, 0u);
      for (uint32_t j = 0; j < sc.n_peds; ++j) {
        const float qx = (float)(sc.peds[j].x - R.x) * (float)c_obs_d, qy = (float)(sc.peds[j].y - R.y) * (float)c_obs_d;
        for (uint32_t g = 0; g < n_cl; ++g) {
          const float dx = qx - rec[g * SFW_OBST_CLUSTER_SLOTS].x, dy = qy - rec[g * SFW_OBST_CLUSTER_SLOTS].y;
          if (dx * dx + dy * dy <= rec[g * SFW_OBST_CLUSTER_SLOTS + 1].x) {
            reach[(size_t)j * words + g / 64u] |= 1ull << (g & 63u);
            ++reach_n[j];
          }
        }
      }
      std::sort(order, order + sc.n_peds, [&](uint32_t a, uint32_t b) {
        if (reach_n[a] != reach_n[b])
          return reach_n[a] > reach_n[b];
        for (uint32_t w = 0; w < words; ++w)
          if (reach[(size_t)a * words + w] != reach[(size_t)b * words + w])
            return reach[(size_t)a * words + w] < reach[(size_t)b * words + w];
        return a < b;
      });
      // share of (pair, cluster) tests that skip the cluster at the start positions (reported only)
      for (uint32_t k = 0; k < sc.n_peds; k += 2) {
        const uint32_t a = order[k], b = (k + 1 < sc.n_peds) ? order[k + 1] : order[k];
        for (uint32_t w = 0; w < words; ++w) {
          const uint64_t either = reach[(size_t)a * words + w] | reach[(size_t)b * words + w];
          W.cull_skipped += (w + 1 < words ? 64u : n_cl - 64u * w) - (uint32_t)__builtin_popcountll(either);
        }
        W.cull_tests += n_cl;
      }
    }
    for (uint32_t j = 0; j < sc.n_peds; ++j)
      W.slot[order[j]] = j; // caller's index -> packed position

    // -- pedestrian groups: lightsfm only applies group forces to groups with >= 2 members
    uint32_t n_groups = 0;
    if (X.off_grp[s + 1] > X.off_grp[s]) {
      std::vector<std::pair<int32_t, uint32_t>> &tagged = W.tagged;
      tagged.clear();
      for (uint32_t j = 0; j < sc.n_peds; ++j)
        if (sc.peds[j].group_id >= 0)
          tagged.emplace_back(sc.peds[j].group_id, j);
      std::stable_sort(tagged.begin(), tagged.end(),
                       [](const std::pair<int32_t, uint32_t> &a, const std::pair<int32_t, uint32_t> &b) { return a.first < b.first; });
      W.starts.clear();
      W.members.clear();
      for (size_t i = 0; i < tagged.size();) {
        size_t e = i;
        while (e < tagged.size() && tagged[e].first == tagged[i].first)
          ++e;
        if (e - i >= 2) {
          W.starts.push_back((uint32_t)(W.members.size() / 2));
          for (size_t m = i; m < e; ++m) {
            const float rad = (float)sc.peds[tagged[m].second].radius;
            uint32_t bits;
            memcpy(&bits, &rad, 4);
            W.members.push_back(W.slot[tagged[m].second]); // index in the packed order
            W.members.push_back(bits);
          }
        }
        i = e;
      }
      n_groups = (uint32_t)W.starts.size();
      if (n_groups) {
        W.starts.push_back((uint32_t)(W.members.size() / 2));
        uint32_t *g = hG + X.off_grp[s];
        memcpy(g, W.starts.data(), 4 * W.starts.size());
        memcpy(g + W.starts.size(), W.members.data(), 4 * W.members.size());
      }
    }
    X.grp_cnt[s] = n_groups;

    // -- the scene record
    SfwSceneDev &d = hs[s];
    memset(&d, 0, sizeof(d));
    d.rx = R.x;
    d.ry = R.y;
    d.rth = R.theta;
    d.rvx = R.vx;
    d.rvy = R.vy;
    d.rvth = R.vtheta;
    d.wpx = R.wpx;
    d.wpy = R.wpy;
    d.origin_x = sc.origin_x;
    d.origin_y = sc.origin_y;
    d.resolution = sc.resolution;
    d.ax = (float)(R.agent_x - R.x);
    d.ay = (float)(R.agent_y - R.y);
    d.avx = (float)R.agent_vx;
    d.avy = (float)R.agent_vy;
    const double obs_norm = sc.n_obstacles ? sfm.force_factor_obstacle / (double)sc.n_obstacles : 0.0;
    d.a_obs_scale = (float)(obs_norm * std::exp(R.agent_radius * inv_sigma));
    d.size_x = sc.size_x;
    d.size_y = sc.size_y;
    d.win_x0 = X.wx0[s];
    d.win_y0 = X.wy0[s];
    const uint32_t n_pairs = (sc.n_peds + 1) / 2;
    d.n_peds = sc.n_peds;
    d.n_pairs = n_pairs;
    d.n_obst = slots;
    d.n_fp = sc.n_footprint;
    d.ped_off = pP;
    d.obs_off = pM;
    d.fp_off = pF;
    d.map_off = slot * s;
    d.goal_mask = 0;
    d.n_groups = n_groups;
    d.grp_off = X.off_grp[s];
    {
      double circ = 0.0;
      for (uint32_t k = 0; k < sc.n_footprint; ++k)
        circ = std::max(circ, std::hypot(sc.footprint_xy[2 * k], sc.footprint_xy[2 * k + 1]));
      const double rcells = std::floor(circ / sc.resolution) + 2.0;
      d.fp_rc = (sc.n_footprint >= 3 && rcells < 64.0) ? (uint32_t)rcells : 0u;
    }
    // Pedestrian pair (2k, 2k+1): one float4 per quantity = (q0, q1) x (x, y).  An odd crowd is padded
    // with an agent SFW_FAR_AWAY from everything: all its pair/obstacle terms underflow to exactly 0.
    for (uint32_t k = 0; k < n_pairs; ++k) {
      float q[2][10];
      for (uint32_t hlf = 0; hlf < 2; ++hlf) {
        const uint32_t j = 2 * k + hlf;
        float *v = q[hlf];
        if (j < sc.n_peds) {
          const SfwPed &p = sc.peds[order[j]];
          v[0] = (float)(p.x - R.x);
          v[1] = (float)(p.y - R.y);
          v[2] = (float)p.vx;
          v[3] = (float)p.vy;
          v[4] = (float)(p.goal_x - R.x);
          v[5] = (float)(p.goal_y - R.y);
          v[6] = (float)(p.goal_radius * p.goal_radius);
          v[7] = (float)p.desired_velocity;
          v[8] = (float)(obs_norm * std::exp(p.radius * inv_sigma));
          v[9] = (float)(p.desired_velocity * p.desired_velocity);
          hBits[2 * (pP + k) + hlf] = p.has_goal ? 1 : 0;
          if (p.has_goal && j < 64)
            d.goal_mask |= (1ull << j);
        } else {
          hBits[2 * (pP + k) + hlf] = 0;
          v[0] = v[4] = SFW_FAR_AWAY;
          v[1] = v[2] = v[3] = v[5] = v[6] = v[7] = v[8] = v[9] = 0.f;
        }
      }
      hPos[pP + k] = make_float4(q[0][0], q[1][0], q[0][1], q[1][1]);
      hVel[pP + k] = make_float4(q[0][2], q[1][2], q[0][3], q[1][3]);
      hGoal[pP + k] = make_float4(q[0][4], q[1][4], q[0][5], q[1][5]);
      hPar[pP + k] = make_float4(q[0][6], q[1][6], q[0][7], q[1][7]);
      hPar2[pP + k] = make_float4(q[0][8], q[1][8], q[0][9], q[1][9]);
    }
    for (uint32_t k = 0; k < sc.n_footprint; ++k)
      hF[pF + k] = make_double2(sc.footprint_xy[2 * k], sc.footprint_xy[2 * k + 1]);
    // -- this scene's slices of the sharing tables
    if (share_candidate) {
      memcpy(h + o_skv + 2 * (size_t)s * n_v, X.kv.data() + (size_t)s * n_v, 2 * (size_t)n_v);
      memcpy(h + o_skw + 2 * (size_t)s * n_w, X.kw.data() + (size_t)s * n_w, 2 * (size_t)n_w);
      memcpy(h + o_sdv + (size_t)s * n_v, X.dv.data() + (size_t)s * n_v, n_v);
      memcpy(h + o_sdw + (size_t)s * n_w, X.dw.data() + (size_t)s * n_w, n_w);
      memcpy(h + o_sperm + 4 * (size_t)s * n_w, X.perm.data() + (size_t)s * n_w, 4 * (size_t)n_w);
      memcpy(h + o_srperm + 4 * (size_t)s * n_v, X.rperm.data() + (size_t)s * n_v, 4 * (size_t)n_v);
      memcpy(h + o_slr + 4 * (size_t)s * L, X.lvl_rows.data() + (size_t)s * L, 4 * (size_t)L);
      memcpy(h + o_slc + 4 * (size_t)s * L, X.lvl_cols.data() + (size_t)s * L, 4 * (size_t)L);
    }
  };
  // ... and its costmap slot: the bulk of the bytes
  auto pack_map = [&](uint32_t s, unsigned) {
    const SfwScene &sc = scenes[s];
    uint8_t *dst = h + o_maps + slot * s;
    if (sc.size_x == map_pitch) {
      memcpy(dst, sc.costmap, (size_t)sc.size_x * sc.size_y);
    } else {
      for (uint32_t r = 0; r < sc.size_y; ++r)
        memcpy(dst + (size_t)r * map_pitch, sc.costmap + (size_t)r * sc.size_x, sc.size_x);
    }
  };

  // A small batch (and the single scene of a control tick) is packed in one pass and shipped in one copy on the
  // context stream.  A big one goes in pieces on the context's COPY stream — 8 MB, 16 MB, 32 MB ... of costmap slots,
  // each packed by the workers (scene records and costmap in one pass) while the previous piece is on its way —
  // and, from sfw_score_batch, is LAUNCHED in two groups: the scorer starts on the first quarter of the scenes as
  // soon as they have landed and runs beside the packing and the copies of the other three quarters.
  const bool pieces = slot * n_scenes > ((size_t)16 << 20);
  bool launched = false;
  if (!pieces) {
    c->pool.run(n_scenes, [&](uint32_t s, unsigned wi) {
      pack_scene(s, wi);
      pack_map(s, wi);
    }, 8);
    CK(c, cudaMemcpyAsync(c->in.dev, c->in.host, in_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaEventRecord(c->h2d_done, c->stream));
  } else {
    const bool launch_groups = launch_too && !c->plan.crowd;
    RunState rs;
    if (launch_groups) {
      c->slab_begin = 0;
      c->slab_end = 0xffffffffu;
      rc = run_prepare(c, rs);
      if (rc != SFW_OK)
        return rc;
    }
    // the device arena may still be read by what is queued on the context stream: copies start behind it
    CK(c, cudaEventRecord(c->ev_compute, c->stream));
    CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_compute, 0));
    const uint32_t first_group = launch_groups ? std::max<uint32_t>(1u, (n_scenes + 3u) / 4u) : n_scenes;
    size_t piece_bytes = (size_t)8 << 20;
    for (uint32_t g0 = 0; g0 < n_scenes;) {
      const uint32_t g1 = g0 == 0 ? first_group : n_scenes;
      for (uint32_t s0 = g0; s0 < g1;) {
        const uint32_t cnt = std::min<uint32_t>(std::max<uint32_t>(1u, (uint32_t)(piece_bytes / slot)), g1 - s0);
        c->pool.run(cnt, [&](uint32_t i, unsigned wi) {
          pack_scene(s0 + i, wi);
          pack_map(s0 + i, wi);
        }, 8);
        CK(c, cudaMemcpyAsync(c->in.dev + o_maps + slot * s0, c->in.host + o_maps + slot * s0, slot * cnt,
                              cudaMemcpyHostToDevice, c->copy_stream));
        s0 += cnt;
        piece_bytes = std::min<size_t>(piece_bytes * 2, (size_t)64 << 20);
      }
      // everything that is not a costmap (a few per cent of the bytes), for the scenes packed so far; the second
      // group's copy re-sends the first group's bytes unchanged
      CK(c, cudaMemcpyAsync(c->in.dev, c->in.host, o_maps, cudaMemcpyHostToDevice, c->copy_stream));
      CK(c, cudaEventRecord(c->ev_copy, c->copy_stream));
      CK(c, cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); // the context stream continues behind this group's copies
      if (launch_groups) {
        rc = run_launch(c, rs, g0, g1 - g0);
        if (rc != SFW_OK)
          return rc;
      }
      g0 = g1;
    }
    CK(c, cudaEventRecord(c->h2d_done, c->copy_stream));
    launched = launch_groups;
  }
  c->in_bytes = in_bytes;
  c->scene_host.assign(hs, hs + n_scenes);
  uint64_t cull_skipped = 0, cull_tests = 0;
  for (const SfwScratch::Worker &w : X.w) {
    cull_skipped += w.cull_skipped;
    cull_tests += w.cull_tests;
  }

  // SURVEY.md 8(d) algorithmic bytes
  uint64_t ab = 0;
  for (uint32_t s = 0; s < n_scenes; ++s) {
    const SfwScene &sc = scenes[s];
    ab += (uint64_t)sc.size_x * sc.size_y + 32ull * sc.n_peds + 8ull * sc.n_obstacles +
          16ull * sc.n_footprint + 128ull + 4ull * (n_v + n_w) + 4ull * samples + 16ull;
  }
  c->algo_bytes = ab;
  c->obst_skip_frac = cull_tests ? (double)cull_skipped / (double)cull_tests : 0.0;
  c->slab_begin = 0;
  c->slab_end = 0xffffffffu;
  c->staged = true;
  c->ran = launched;
  return SFW_OK;
}

int sfw_upload(sfw_ctx *c, const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scenes,
               uint32_t n_scenes, const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w) {
  return upload_impl(c, params, sfm, scenes, n_scenes, linvels, n_v, angvels, n_w, false);
}

int sfw_set_host_threads(sfw_ctx *c, int n_threads) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (n_threads < 1 || n_threads > 64)
    return fail(c, SFW_ERR_ARG, "sfw_set_host_threads: need 1 <= n_threads <= 64");
  c->host_threads = n_threads;
  if (c->pool.size() > (unsigned)n_threads || (unsigned)n_threads == 1u)
    c->pool.resize((unsigned)n_threads);
  return SFW_OK;
}

int sfw_set_policy(sfw_ctx *c, int policy) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (policy < SFW_POLICY_AUTO || policy > SFW_POLICY_LATENCY)
    return fail(c, SFW_ERR_ARG, "sfw_set_policy: unknown policy %d", policy);
  c->policy = policy;
  c->plan.valid = false;
  return SFW_OK;
}

int sfw_set_zero_sample(sfw_ctx *c, int score_it) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  c->score_zero = score_it ? 1 : 0;
  return SFW_OK;
}

int sfw_set_obstacle_cutoff(sfw_ctx *c, double cutoff_log2) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (cutoff_log2 != cutoff_log2)
    return fail(c, SFW_ERR_ARG, "sfw_set_obstacle_cutoff: NaN");
  c->obst_cutoff_log2 = cutoff_log2;
  return SFW_OK;
}

uint32_t sfw_obstacle_layout(const double *obstacles_xy, uint32_t n, double ref_x, double ref_y, double sigma,
                             double r_max, double cutoff_log2, float *slots_out, uint32_t slots_cap) {
  const uint32_t slots = sfw_obst_slots(n);
  if (!n || !obstacles_xy || !slots_out || !(sigma > 0.0))
    return (obstacles_xy && sigma > 0.0) ? slots : 0u;
  const double c_obs_d = (double)(float)(1.4426950408889634 / sigma); // as sfw_upload
  std::vector<float2> pts(n), rec(slots);
  for (uint32_t k = 0; k < n; ++k)
    pts[k] = make_float2((float)((obstacles_xy[2 * k] - ref_x) * c_obs_d),
                         (float)((obstacles_xy[2 * k + 1] - ref_y) * c_obs_d));
  std::vector<uint32_t> idx;
  pack_obstacles(pts, rec.data(), cutoff_log2, (double)(float)r_max * c_obs_d, idx);
  memcpy(slots_out, rec.data(), sizeof(float2) * std::min(slots, slots_cap));
  return slots;
}

int sfw_set_prefix_sharing(sfw_ctx *c, int on) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  c->share_allowed = on < 0 ? 0 : on > 2 ? 2 : on;
  return SFW_OK;
}

int sfw_set_row_slab(sfw_ctx *c, uint32_t row_begin, uint32_t row_end) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->staged)
    return fail(c, SFW_ERR_STATE, "sfw_set_row_slab before sfw_upload");
  if (row_end > c->B.n_v)
    row_end = c->B.n_v;
  if (row_begin > row_end)
    return fail(c, SFW_ERR_ARG, "row_begin > row_end");
  c->slab_begin = row_begin;
  c->slab_end = row_end;
  return SFW_OK;
}

extern "C++" {
namespace {

// What one sfw_run does, in two halves, so that a batch can also be launched piece by piece while the rest of it
// is still being packed (sfw_score_batch): run_prepare = row slab + fork tables of the slab + winner-exchange
// bookkeeping (once per run), run_launch = the kernel launches for scenes [s0, s0 + cnt).
int run_prepare(sfw_ctx *c, RunState &rs) {
  SfwBatchDev &B = c->B;
  const uint32_t rb = std::min(c->slab_begin, B.n_v), re = std::min(c->slab_end, B.n_v);
  rs.rb = rb;
  rs.re = re;
  if (rb != B.row_begin || re != B.row_end) {
    // rows outside the slab keep SFW_COST_SKIPPED
    B.row_begin = rb;
    B.row_end = re;
    const uint32_t samples = (re - rb) * B.n_w;
    Plan saved = c->plan;
    c->plan.valid = false;
    int rc = make_plan(c, B.n_scenes, std::max(samples, 1u), saved.maxP, saved.maxM, saved.maxF,
                       saved.win_wp, saved.win_h, saved.steps, saved.crowd ? 1 : 0, saved.crowd ? saved.T : 0u);
    if (rc != SFW_OK)
      return rc;
    B.tiles_per_scene = c->plan.tiles;
    // fill the whole cost vector with SKIPPED once; the kernel overwrites the slab
    const size_t n = (size_t)B.n_scenes * B.n_v * B.n_w;
    {
      // cudaMemsetAsync only sets bytes; -2.0f = 0xC0000000 needs a 32-bit fill
      CUresult (*fill32)(CUdeviceptr, unsigned int, size_t, CUstream) = nullptr;
      void *p = nullptr;
      cudaDriverEntryPointQueryResult qr;
      if (cudaGetDriverEntryPoint("cuMemsetD32Async", &p, cudaEnableDefault, &qr) == cudaSuccess &&
          qr == cudaDriverEntryPointSuccess) {
        fill32 = (CUresult(*)(CUdeviceptr, unsigned int, size_t, CUstream))p;
        if (fill32((CUdeviceptr)(uintptr_t)B.costs, 0xC0000000u, n, (CUstream)c->stream) != CUDA_SUCCESS)
          return fail(c, SFW_ERR_CUDA, "cuMemsetD32Async failed");
      } else {
        return fail(c, SFW_ERR_CUDA, "cuMemsetD32Async entry point unavailable");
      }
    }
    CK(c, cudaMemsetAsync(B.npts, 0, n * 2, c->stream));
  }
  // Rollout prefix sharing on a row slab (thread-per-trajectory family): the sample launch walks the SLAB's samples in
  // fork order, so row_perm / lvl_rows must describe the slab's rows.  Rebuilt here (a few KB per scene) whenever
  // the slab changes; the path launches still write every shared path (columns are not sharded).
  rs.slab_share = re > rb && c->share_active && !c->plan.crowd && (rb == 0 && re == B.n_v ? true : (uint64_t)(re - rb) * 8u >= B.n_v);
  if (rs.slab_share && (c->share_rows_begin != rb || c->share_rows_end != re)) {
    SfwScratch &X = c->scratch;
    const uint32_t L = c->share_kmax + 2u, n_v = B.n_v;
    CK(c, cudaEventSynchronize(c->h2d_done)); // the staging arena is ours again
    uint32_t *hr = reinterpret_cast<uint32_t *>(c->in.host + c->share_off_rperm);
    uint32_t *hl = reinterpret_cast<uint32_t *>(c->in.host + c->share_off_lvl_rows);
    std::vector<uint32_t> cur(L);
    for (uint32_t s = 0; s < B.n_scenes; ++s) {
      const uint16_t *kvs = X.kv.data() + (size_t)s * n_v;
      uint32_t *lr = hl + (size_t)s * L, *rp = hr + (size_t)s * n_v;
      std::fill(lr, lr + L, 0u);
      for (uint32_t r = rb; r < re; ++r)
        ++lr[kvs[r] + 1u];
      for (uint32_t k = 1; k < L; ++k)
        lr[k] += lr[k - 1u];
      std::copy(lr, lr + L, cur.begin());
      for (uint32_t r = rb; r < re; ++r)
        rp[cur[kvs[r]]++] = r;
    }
    CK(c, cudaMemcpyAsync(c->in.dev + c->share_off_rperm, hr, 4 * (size_t)B.n_scenes * n_v, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(c->in.dev + c->share_off_lvl_rows, hl, 4 * (size_t)B.n_scenes * L, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaEventRecord(c->h2d_done, c->stream));
    c->share_rows_begin = rb;
    c->share_rows_end = re;
  }
  // fused winner exchange: this run writes epoch parity `slot` of every rank's gather buffer
  memset(&B.xchg, 0, sizeof(B.xchg));
  if (c->xchg.connected) {
    if (B.n_scenes > c->xchg.max_scenes)
      return fail(c, SFW_ERR_ARG, "%u scenes staged but the exchange buffer holds %u", B.n_scenes, c->xchg.max_scenes);
    if (c->xchg.counts_set && c->xchg.counts[c->xchg.rank] != B.n_scenes)
      return fail(c, SFW_ERR_ARG, "%u scenes staged but sfw_exchange_expect announced %u for this rank", B.n_scenes,
                  c->xchg.counts[c->xchg.rank]);
    for (uint32_t q = 0; q < c->xchg.world; ++q) {
      B.xchg.peer_best[q] = reinterpret_cast<SfwBest *>(c->xchg.peer[q]);
      B.xchg.peer_arrived[q] = reinterpret_cast<unsigned int *>((uint8_t *)c->xchg.peer[q] + c->xchg.off_arrived);
    }
    B.xchg.rank = c->xchg.rank;
    B.xchg.world = c->xchg.world;
    B.xchg.max_scenes = c->xchg.max_scenes;
    B.xchg.enabled = 1;
    B.xchg.slot = (uint32_t)(c->xchg.epoch & 1u);
    c->xchg.epoch += 1;
    for (uint32_t q = 0; q < c->xchg.world; ++q) { // what every rank delivers this tick (sfw_exchange_expect)
      const uint32_t cnt = c->xchg.counts_set ? c->xchg.counts[q] : B.n_scenes;
      c->xchg.last_counts[q] = cnt;
      c->xchg.expected[q] += cnt;
    }
  }
  return SFW_OK;
}

// The launches of one run for scenes [s0, s0 + cnt).  The block-per-trajectory family pulls work items of the whole
// batch from one counter, so it only takes the full range.
int run_launch(sfw_ctx *c, const RunState &rs, uint32_t s0, uint32_t cnt) {
  SfwBatchDev &B = c->B;
  const uint32_t rb = rs.rb, re = rs.re;
  const bool slab_share = rs.slab_share;
  const bool whole = s0 == 0 && cnt == B.n_scenes;
  if (c->plan.crowd && !whole)
    return fail(c, SFW_ERR_STATE, "internal: the block-per-trajectory family is launched for the whole batch");
  if (re > rb && c->plan.crowd && c->share_active && rb == 0 && re == B.n_v) {
    // rollout prefix sharing, block-per-trajectory flavour: paths, paths, samples (+ arg-min)
    unsigned int *wc = reinterpret_cast<unsigned int *>(c->out.dev + c->off_work);
    SfwBatchDev W = B;
    for (uint32_t mode = 1; mode <= 3; ++mode) {
      W.share.mode = mode;
      const uint64_t items = (uint64_t)B.n_scenes * (mode == 1 ? 4u : mode == 2 ? c->share_paths - 4u : c->out_samples);
      CK(c, sfw_launch_crowd(W, wc, (uint32_t)std::min<uint64_t>(items, c->plan.grid), c->plan.T, c->plan.smem, c->stream, mode == 3));
    }
    c->launches += sfw_crowd_fuses_argmin(W) ? 3 : 4;
    c->last_kernel = "sfw_score_crowd,share";
  } else if (re > rb && c->plan.crowd) {
    SfwBatchDev W = B;
    W.share.mode = 0;
    CK(c, sfw_launch_crowd(W, reinterpret_cast<unsigned int *>(c->out.dev + c->off_work), c->plan.grid, c->plan.T,
                           c->plan.smem, c->stream, true));
    c->launches += sfw_crowd_fuses_argmin(W) ? 1 : 2; // scorer (+ arg-min)
    c->last_kernel = "sfw_score_crowd";
  } else if (slab_share) {
    // rollout prefix sharing: the 4 doubly saturated paths, the 2 (n_v + n_w) singly saturated ones (each
    // continuing one of the 4), then every sample from the record of its own fork point
    // (the path launches are latency bound: small blocks, so that every scene's few paths are resident at once --
    // but not smaller than 128 threads: the block's prologue, the free-space bit map of the window, is shared work)
    SfwBatchDev W = B;
    W.scene_base = s0;
    W.launch_scenes = cnt;
    const uint32_t T1 = 128, T2 = 128;
    const uint32_t per = SFW_PATH_WARP_THREADS / 32u;
    const size_t smw = sfw_small_smem_bytes(B.win_wp, B.win_h, c->plan.maxP, c->plan.maxM, c->plan.maxF,
                                            SFW_PATH_WARP_THREADS);
    if (c->share_warp == 3u && c->share_merged) {
      // both path stages in one launch: tile 0 of a scene writes the 4 doubly saturated paths, the other tiles
      // wait (per record) for the one they continue from
      W.share.mode = 4;
      c->share_epoch = c->share_epoch == 0xffffffffu ? 1u : c->share_epoch + 1u;
      W.share.epoch = c->share_epoch;
      W.tiles_per_scene = 1u + (c->share_paths - 4u + per - 1u) / per;
      CK(c, sfw_launch_warp_paths(W, c->tmap, smw, c->stream));
      c->launches -= 1; // (3 is added below)
    } else {
      W.share.mode = 1;
      W.tiles_per_scene = 1;
      if (c->share_warp & 1u) // a warp per path, a lane per pedestrian pair
        CK(c, sfw_launch_warp_paths(W, c->tmap, smw, c->stream));
      else
        CK(c, sfw_launch_small(W, c->tmap, T1,
                               sfw_small_smem_bytes(B.win_wp, B.win_h, c->plan.maxP, c->plan.maxM, c->plan.maxF, T1),
                               c->stream));
      W.share.mode = 2;
      if (c->share_warp & 2u) {
        W.tiles_per_scene = (c->share_paths - 4u + per - 1u) / per;
        CK(c, sfw_launch_warp_paths(W, c->tmap, smw, c->stream, c->pdl));
      } else {
        W.tiles_per_scene = (c->share_paths - 4u + T2 - 1u) / T2;
        CK(c, sfw_launch_small(W, c->tmap, T2,
                               sfw_small_smem_bytes(B.win_wp, B.win_h, c->plan.maxP, c->plan.maxM, c->plan.maxF, T2),
                               c->stream, c->pdl));
      }
    }
    W.share.mode = 3;
    W.tiles_per_scene = B.tiles_per_scene;
    if (rb != 0 || re != B.n_v)
      W.share.chunk_map = nullptr; // the dealt order was built for the whole grid's blocks
    CK(c, sfw_launch_small(W, c->tmap, c->plan.T, c->plan.smem, c->stream, c->pdl));
    c->launches += 3;
    c->last_kernel = sfw_small_kernel_name(c->plan.T, true);
  } else if (re > rb) {
    SfwBatchDev W = B;
    W.scene_base = s0;
    W.launch_scenes = cnt;
    W.share.mode = 0;
    CK(c, sfw_launch_small(W, c->tmap, c->plan.T, c->plan.smem, c->stream));
    c->launches += 1;
    c->last_kernel = sfw_small_kernel_name(c->plan.T);
  } else {
    // empty slab: nothing to score, every scene reports "no valid trajectory" — to the peers too, who wait for
    // this rank's records like for anyone else's
    CK(c, cudaMemsetAsync(B.best, 0, sizeof(SfwBest) * B.n_scenes, c->stream));
    if (B.xchg.enabled) {
      CK(c, sfw_launch_export_invalid(B.xchg, B.n_scenes, c->stream));
      c->launches += 1;
    }
  }
  return SFW_OK;
}

} // namespace
} // extern "C++"

int sfw_run(sfw_ctx *c) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->staged)
    return fail(c, SFW_ERR_STATE, "sfw_run before sfw_upload");
  CK(c, cudaSetDevice(c->device));
  RunState rs;
  int rc = run_prepare(c, rs);
  if (rc != SFW_OK)
    return rc;
  rc = run_launch(c, rs, 0, c->B.n_scenes);
  if (rc != SFW_OK)
    return rc;
  c->ran = true;
  return SFW_OK;
}

int sfw_sync(sfw_ctx *c) {
  if (!c)
    return SFW_ERR_ARG;
  CK(c, cudaStreamSynchronize(c->stream));
  return SFW_OK;
}

int sfw_download(sfw_ctx *c, float *costs_out, SfwBest *best_out) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->ran)
    return fail(c, SFW_ERR_STATE, "sfw_download before sfw_run");
  CK(c, cudaSetDevice(c->device));
  const size_t nb = sizeof(SfwBest) * c->B.n_scenes;
  const size_t nc = 4 * (size_t)c->out_samples * c->B.n_scenes;
  // (a row slab fills the rows outside it on the device only, and an empty slab writes nothing: copy then)
  const bool landed = c->zero_copy_out && c->B.row_begin == 0 && c->B.row_end == c->B.n_v;
  if (!landed) {
    if (best_out)
      CK(c, cudaMemcpyAsync(c->out.host + c->off_best, c->out.dev + c->off_best, nb,
                            cudaMemcpyDeviceToHost, c->stream));
    if (costs_out)
      CK(c, cudaMemcpyAsync(c->out.host + c->off_costs, c->out.dev + c->off_costs, nc,
                            cudaMemcpyDeviceToHost, c->stream));
  }
  CK(c, cudaStreamSynchronize(c->stream));
  if (c->status[0]) { // a kernel gave up (it says why) instead of trapping the context
    const unsigned int code = c->status[0];
    c->status[0] = 0u;
    return fail(c, SFW_ERR_STATE, "scorer kernel gave up (device status %u: %s); results discarded", code,
                code == SFW_DEVSTAT_PATH_WAIT ? "a shared-path record it continues from never appeared" : "unknown");
  }
  if (best_out)
    memcpy(best_out, c->out.host + c->off_best, nb);
  if (costs_out) {
    // a batch's cost vectors are tens of megabytes: the host workers copy them out of the landing buffer in slices
    const size_t slice = (size_t)1 << 20;
    const uint32_t n_slices = (uint32_t)((nc + slice - 1) / slice);
    const uint8_t *src = c->out.host + c->off_costs;
    uint8_t *dst = reinterpret_cast<uint8_t *>(costs_out);
    c->pool.run(n_slices, [&](uint32_t i, unsigned) {
      const size_t o = (size_t)i * slice;
      memcpy(dst + o, src + o, std::min(slice, nc - o));
    }, 8);
  }
  return SFW_OK;
}

int sfw_score_batch(sfw_ctx *c, const SfwParams *params, const SfwSfmParams *sfm,
                    const SfwScene *scenes, uint32_t n_scenes, const double *linvels, uint32_t n_v,
                    const double *angvels, uint32_t n_w, float *costs_out, SfwBest *best_out) {
  // a big batch is launched piece by piece from inside the upload (the scorer overlaps the packing of the rest)
  int rc = upload_impl(c, params, sfm, scenes, n_scenes, linvels, n_v, angvels, n_w, true);
  if (rc != SFW_OK)
    return rc;
  if (!c->ran) {
    rc = sfw_run(c);
    if (rc != SFW_OK)
      return rc;
  }
  return sfw_download(c, costs_out, best_out);
}

int sfw_score(sfw_ctx *c, const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
              const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
              float *costs_out, SfwBest *best_out) {
  return sfw_score_batch(c, params, sfm, scene, 1, linvels, n_v, angvels, n_w, costs_out, best_out);
}

int sfw_trajectory_points(sfw_ctx *c, uint32_t scene, uint32_t sample_index, double *xyz_out,
                          uint32_t max_points, uint32_t *n_points) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->ran)
    return fail(c, SFW_ERR_STATE, "sfw_trajectory_points before sfw_run");
  if (scene >= c->B.n_scenes || sample_index >= c->out_samples || !n_points)
    return fail(c, SFW_ERR_ARG, "sfw_trajectory_points: index out of range");
  CK(c, cudaSetDevice(c->device));
  uint16_t np16 = 0;
  CK(c, cudaMemcpyAsync(&np16, c->B.npts + (size_t)scene * c->out_samples + sample_index, 2,
                        cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  *n_points = np16;
  const uint32_t n = std::min<uint32_t>(np16, max_points);
  if (!n || !xyz_out)
    return SFW_OK;
  if (n > c->points_cap) {
    if (c->d_points)
      cudaFree(c->d_points);
    if (c->h_points)
      cudaFreeHost(c->h_points);
    c->d_points = nullptr;
    c->h_points = nullptr;
    c->points_cap = 0;
    CK(c, cudaMalloc((void **)&c->d_points, 24 * (size_t)n));
    CK(c, cudaMallocHost((void **)&c->h_points, 24 * (size_t)n));
    c->points_cap = n;
  }
  CK(c, sfw_launch_points(c->B, scene, sample_index, n, c->d_points, c->stream));
  c->launches += 1;
  CK(c, cudaMemcpyAsync(c->h_points, c->d_points, 24 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  memcpy(xyz_out, c->h_points, 24 * (size_t)n);
  return SFW_OK;
}

int sfw_marker_points(sfw_ctx *c, uint32_t scene, uint32_t first, uint32_t stride, uint32_t count, double *xyz_out,
                      uint32_t max_points, uint16_t *n_points_out) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->ran)
    return fail(c, SFW_ERR_STATE, "sfw_marker_points before sfw_run");
  if (!count)
    return SFW_OK;
  if (!stride)
    stride = 1;
  if (scene >= c->B.n_scenes || !n_points_out || (!xyz_out && max_points) ||
      (uint64_t)first + (uint64_t)(count - 1) * stride >= c->out_samples)
    return fail(c, SFW_ERR_ARG, "sfw_marker_points: sample range out of bounds");
  CK(c, cudaSetDevice(c->device));
  const size_t o_n = 0, o_xyz = align_up(2 * (size_t)count, kAlign);
  const size_t bytes = o_xyz + 24 * (size_t)count * max_points;
  int rc = arena_reserve(c, c->sensor_out, bytes); // scratch arena (shared with the laser path)
  if (rc != SFW_OK)
    return rc;
  uint8_t *dv = c->sensor_out.dev, *hv = c->sensor_out.host;
  CK(c, sfw_launch_marker_points(c->B, scene, first, stride, count, max_points,
                                 reinterpret_cast<double *>(dv + o_xyz), reinterpret_cast<uint16_t *>(dv + o_n),
                                 c->stream));
  c->launches += 1;
  CK(c, cudaMemcpyAsync(hv, dv, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  memcpy(n_points_out, hv + o_n, 2 * (size_t)count);
  if (max_points)
    memcpy(xyz_out, hv + o_xyz, 24 * (size_t)count * max_points);
  return SFW_OK;
}

int sfw_may_i_stop(sfw_ctx *c, uint32_t scene, double vl_x, double vl_y, double va, double x, double y, double th,
                   double dt, int32_t *can_stop, uint32_t *steps) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->staged)
    return fail(c, SFW_ERR_STATE, "sfw_may_i_stop before sfw_upload");
  if (scene >= c->B.n_scenes || !can_stop || !(dt > 0.0))
    return fail(c, SFW_ERR_ARG, "sfw_may_i_stop: scene out of range / null output / dt <= 0");
  CK(c, cudaSetDevice(c->device));
  int rc = arena_reserve(c, c->sensor_out, 256);
  if (rc != SFW_OK)
    return rc;
  int *d = reinterpret_cast<int *>(c->sensor_out.dev), *h = reinterpret_cast<int *>(c->sensor_out.host);
  CK(c, sfw_launch_may_i_stop(c->B, scene, vl_x, vl_y, va, x, y, th, dt, d, c->stream));
  c->launches += 1;
  CK(c, cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  *can_stop = h[0];
  if (steps)
    *steps = (uint32_t)h[1];
  return SFW_OK;
}

void *sfw_stream(sfw_ctx *c) { return c ? (void *)c->stream : nullptr; }
const float *sfw_device_costs(sfw_ctx *c) { return (c && c->staged) ? c->B.costs : nullptr; }
const void *sfw_device_best(sfw_ctx *c) { return (c && c->staged) ? (const void *)c->B.best : nullptr; }
uint64_t sfw_kernel_launches(const sfw_ctx *c) { return c ? c->launches : 0; }
uint64_t sfw_algorithmic_bytes(const sfw_ctx *c) { return c ? c->algo_bytes : 0; }
uint64_t sfw_h2d_bytes(const sfw_ctx *c) { return c ? c->in_bytes : 0; }
uint64_t sfw_d2h_bytes(const sfw_ctx *c) {
  return (c && c->staged) ? (sizeof(SfwBest) + 4ull * c->out_samples) * c->B.n_scenes : 0;
}
const char *sfw_last_kernel(const sfw_ctx *c) { return c ? c->last_kernel : "none"; }
uint32_t sfw_block_threads(const sfw_ctx *c) { return (c && c->staged && c->plan.valid) ? c->plan.T : 0u; }
double sfw_obstacle_skip_fraction(const sfw_ctx *c) { return (c && c->staged) ? c->obst_skip_frac : 0.0; }
double sfw_shared_prefix_steps(const sfw_ctx *c) { return (c && c->share_active) ? c->share_mean_s0 : 0.0; }

} // extern "C"
