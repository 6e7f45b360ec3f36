// sfw_ctx.h — the context object behind the C ABI (include/sfw_b200.h), shared by the translation units
// that implement entry points (sfw_abi.cu: scoring path, sfw_sensor.cu: laser scan -> obstacle points).
#ifndef SFW_CTX_H
#define SFW_CTX_H
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "sfw_dev.h"

struct SfwArena {
  uint8_t *host = nullptr; // pinned
  uint8_t *dev = nullptr;
  size_t cap = 0;
};

// Host worker pool of one context: packs the scenes of a batch in parallel (sfw_upload) and copies big results
// out of the pinned landing buffer (sfw_download).  run(n, fn) calls fn(item, worker) for item = 0 .. n - 1 with
// dynamic scheduling; the calling thread is worker 0 and takes part.  Small jobs run inline (no wake-up latency
// on the single-scene control tick).
class SfwPool {
 public:
  ~SfwPool() { stop(); }
  unsigned size() const { return (unsigned)th_.size() + 1u; }
  void resize(unsigned n_threads) { // total workers including the caller
    stop();
    quit_ = false;
    for (unsigned w = 1; w < n_threads; ++w)
      th_.emplace_back([this, w] { loop(w); });
  }
  void run(uint32_t n, const std::function<void(uint32_t, unsigned)> &fn, uint32_t inline_below = 4) {
    if (n < inline_below || th_.empty()) {
      for (uint32_t i = 0; i < n; ++i)
        fn(i, 0u);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      n_ = n;
      next_.store(0u);
      busy_ = (unsigned)th_.size();
      ++gen_;
    }
    cv_.notify_all();
    work(0u);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return busy_ == 0u; });
    fn_ = nullptr;
  }

 private:
  void work(unsigned w) {
    for (;;) {
      const uint32_t i = next_.fetch_add(1u);
      if (i >= n_)
        break;
      (*fn_)(i, w);
    }
  }
  void loop(unsigned w) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
        if (quit_)
          return;
        seen = gen_;
      }
      work(w);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--busy_ == 0u)
          done_.notify_one();
      }
    }
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
    }
    cv_.notify_all();
    for (std::thread &t : th_)
      t.join();
    th_.clear();
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(uint32_t, unsigned)> *fn_ = nullptr;
  uint32_t n_ = 0;
  std::atomic<uint32_t> next_{0};
  unsigned busy_ = 0;
  uint64_t gen_ = 0;
  bool quit_ = false;
};

// Scratch of sfw_upload, kept in the context so that a control tick allocates nothing once it has warmed up.
struct SfwScratch {
  std::vector<uint32_t> off_pairs, off_obst, off_fp, off_grp, off_ped; // per-scene prefix offsets [n_scenes + 1]
  std::vector<uint32_t> grp_cnt;                                       // groups (>= 2 members) per scene
  std::vector<int32_t> wx0, wy0;
  std::vector<uint16_t> kv, kw;
  std::vector<uint8_t> dv, dw;
  std::vector<uint32_t> perm, rperm, lvl_rows, lvl_cols, chunk_map;
  struct Worker {
    std::vector<float2> pts, rec;
    std::vector<uint64_t> reach;
    std::vector<uint32_t> reach_n, order, slot, idx, starts, members, cur;
    std::vector<std::pair<int32_t, uint32_t>> tagged;
    std::vector<double> up, dn;
    uint64_t cull_skipped = 0, cull_tests = 0;
    uint32_t kmax = 0;
    double tot = 0.0;
  };
  std::vector<Worker> w;
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

  int host_threads = 1; // workers a batch is packed with (sfw_set_host_threads); set from the host at sfw_create
  SfwPool pool;       // host workers, started by the first batch of >= 8 scenes
  SfwScratch scratch; // sfw_upload's reusable buffers
  SfwArena in;   // packed inputs
  SfwArena out;  // best | costs | npts | blockbest | counters
  SfwArena sensor_in, sensor_out; // sfw_laser_obstacles / sfw_marker_points staging
  bool laser_attr_set = false;
  int policy = 0; // SFW_POLICY_*
  int score_zero = 0; // sfw_set_zero_sample
  bool pdl = true;    // launches that continue from the previous launch's records use programmatic dependent launch
                      // (SFW_B200_NO_PDL=1 in the environment at sfw_create switches it off: A/B measurements)
  unsigned int *status = nullptr; // mapped pinned [16]: [0] scorer kernels (SFW_DEVSTAT_*), [1] winner exchange
  cudaEvent_t h2d_done = nullptr; // the staging buffer `in.host` may be overwritten once this has fired
  cudaStream_t copy_stream = nullptr; // H2D pieces of a big batch (the context stream keeps computing beside them)
  cudaEvent_t ev_compute = nullptr, ev_copy = nullptr; // context stream -> copy stream, and back
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
  size_t share_off_rperm = 0, share_off_lvl_rows = 0; // byte offsets of row_perm / lvl_rows in the input arena
  uint32_t share_kmax = 0;
  uint32_t share_rows_begin = 0, share_rows_end = 0;  // rows the device tables currently describe

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
  bool zero_copy_out = false; // the kernels also store costs / winners into out.host (small results)
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
