// sfw_ctx.h — the context object behind the C ABI (include/sfw_b200.h), shared by the translation units
// that implement entry points (sfw_abi.cu: scoring path, sfw_sensor.cu: laser scan -> obstacle points).
#ifndef SFW_CTX_H
#define SFW_CTX_H
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "sfw_dev.h"

struct SfwArena {
  uint8_t *host = nullptr; // pinned
  uint8_t *dev = nullptr;
  size_t cap = 0;
};

struct SfwPlan {
  // shape key
  uint32_t n_scenes = 0, samples = 0, maxP = 0, maxM = 0, maxF = 0, win_wp = 0, win_h = 0;
  // result
  uint32_t T = 0, tiles = 0, k = 1;
  size_t smem = 0;
  bool crowd = false; // block-per-trajectory kernel (sfw_crowd.cu)
  uint32_t grid = 0;
  int steps = 0;
  bool valid = false;
};

struct sfw_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  std::string err;
  std::mutex mu;

  SfwArena in;   // packed inputs
  SfwArena out;  // best | costs | npts | blockbest | counters
  SfwArena sensor_in, sensor_out; // sfw_laser_obstacles / sfw_marker_points staging
  bool laser_attr_set = false;
  int policy = 0; // SFW_POLICY_*
  int score_zero = 0; // sfw_set_zero_sample
  unsigned int *status = nullptr; // mapped pinned [16]: [0] scorer kernels (SFW_DEVSTAT_*), [1] winner exchange
  cudaEvent_t h2d_done = nullptr; // the staging buffer `in.host` may be overwritten once this has fired
  double obst_cutoff_log2 = SFW_OBST_CUTOFF_LOG2; // sfw_set_obstacle_cutoff; <= 0: off
  double obst_skip_frac = 0.0;                    // of the staged batch, at the start poses

  // rollout prefix sharing (SfwShareDev)
  int share_allowed = 1;      // sfw_set_prefix_sharing: 0 never, 1 when the cost model says it pays, 2 whenever possible
  bool share_active = false; // decided per upload
  uint32_t share_warp = 0;   // bit 0 / 1: path launch 1 / 2 uses the warp-per-path writer
  bool share_merged = false; // both path stages in one launch (all blocks co-resident)
  uint32_t share_epoch = 0;  // run counter of the merged path launch (record flags)
  uint8_t *share_buf = nullptr;
  size_t share_cap = 0;
  uint32_t share_paths = 0;
  double share_mean_s0 = 0.0;

  // fused multi-GPU winner exchange (csrc/sfw_exchange.cu)
  struct {
    bool exported = false, connected = false;
    bool local_peers = false; // peers are contexts of this process (sfw_exchange_connect_local): nothing to close
    uint32_t rank = 0, world = 1, max_scenes = 0;
    uint8_t *local = nullptr; // [2][world_max][max_scenes] SfwBest | arrived[SFW_MAX_RANKS] | merged[max_scenes]
    size_t bytes = 0, off_arrived = 0, off_merged = 0;
    void *peer[SFW_MAX_RANKS] = {};
    uint64_t epoch = 0;     // launches exported so far
    uint64_t expected[SFW_MAX_RANKS] = {}; // records rank q must have delivered by now
    uint32_t counts[SFW_MAX_RANKS] = {};   // sfw_exchange_expect: scenes rank q stages per tick
    uint32_t last_counts[SFW_MAX_RANKS] = {}; // ... of the latest run
    bool counts_set = false;               // false: every rank stages what this rank stages
    double timeout_s = 10.0;               // bound of the device-side arrival wait
    unsigned int *status = nullptr;        // = ctx status + 1 (mapped pinned): 0 ok, 1 + q = rank q did not deliver in time
    uint8_t *host = nullptr; // pinned landing buffer of sfw_exchange_fetch
  } xchg;
  double *d_points = nullptr;
  double *h_points = nullptr;
  uint32_t points_cap = 0;

  SfwBatchDev B;
  CUtensorMap tmap;
  bool staged = false, ran = false;
  // tensor-map cache key
  const void *tm_ptr = nullptr;
  uint32_t tm_pitch = 0, tm_rows = 0, tm_scenes = 0, tm_wp = 0, tm_h = 0;
  SfwPlan plan;
  size_t off_best = 0, off_costs = 0, off_npts = 0, off_bb = 0, off_cnt = 0, off_work = 0;
  uint32_t out_scenes = 0, out_samples = 0, out_tiles = 0;
  uint64_t launches = 0;
  uint64_t algo_bytes = 0;
  size_t in_bytes = 0;
  const char *last_kernel = "none";
  uint32_t slab_begin = 0, slab_end = 0xffffffffu;
  std::vector<SfwSceneDev> scene_host; // host copy for trajectory_points
};

// error text of a failed sfw_create (no context to hang it on)
std::string &sfw_create_error();

inline int sfw_fail(sfw_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c)
    c->err = buf;
  else
    sfw_create_error() = buf;
  return code;
}

#define SFW_CK(ctx, call)                                                                          \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return sfw_fail((ctx), SFW_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                      __FILE__, __LINE__);                                                         \
  } while (0)

// grow a pinned-host + device buffer pair (contents are not preserved)
int sfw_arena_reserve(sfw_ctx *c, SfwArena &a, size_t bytes);

#endif
