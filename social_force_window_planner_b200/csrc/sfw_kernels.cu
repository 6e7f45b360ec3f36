// sfw_kernels.cu — sm_100a kernels of the DWA + social-force trajectory scorer.
//
// Kernel "small crowd" (sfw_score_small): ONE THREAD PER TRAJECTORY.
//   * grid  = n_scenes x tiles_per_scene blocks, block = T threads (T chosen by the host so that
//     the whole (v,w) grid fits the machine in an integral number of balanced waves).
//   * per block: the scene's reachable costmap window is staged to shared memory with ONE TMA 2D/3D
//     tensor copy (cp.async.bulk.tensor), pedestrian/obstacle/footprint arrays with bulk copies
//     (cp.async.bulk), all completing on one mbarrier.
//   * per thread: the robot rollout (FP64, bit-faithful to the reference's kinematics and cell
//     indexing) plus a private FP32 social-force simulation of all pedestrians whose state lives
//     in a shared-memory column owned by the thread (conflict-free float4 / float2 accesses).
//   * algebra that removes work without changing results (DESIGN.md "Exact savings"):
//       - the lightsfm pair force is antisymmetric (F_ab = -F_ba when all agents share
//         sfm::Parameters, which the reference guarantees), so each unordered pair is evaluated
//         once: N(N-1)/2 instead of N(N-1) evaluations per step;
//       - the per-pedestrian "social work" pair evaluation of computeSocialWork at step i uses
//         exactly the states that computeForces sees at step i+1, so it is taken from there; only
//         the last step needs one extra robot-pedestrian pass;
//       - the desired/obstacle forces computeSocialWork computes on its copies are never read by
//         the reference and are not computed here.
//   * epilogue: block arg-min with the reference's tie-break order, then the last block of each
//     scene reduces the tile winners (threadfence + counter), all in the same launch.
//
// Reference semantics followed (paths relative to the reference repo):
//   src/sfw_planner.cpp:338-417 (sample loop + arg-min), :475-676 (scoreTrajectory),
//   :678-705 (computeSocialWork), include/.../sfw_planner.hpp:399-463 (kinematics),
//   include/.../world_model.hpp:45-75 + src/costmap_model.cpp:21-121 +
//   include/.../line_iterator.hpp:37-124 (footprint rasterisation), lightsfm.
#include <algorithm>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "sfw_dev.h"
#include "sfw_kernels.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#include "sfw_forces.cuh"

#ifndef SFW_SMALL_PREFETCH
#define SFW_SMALL_PREFETCH 1
#endif

// ================================================================================================
// Kernel: one thread per trajectory
// ================================================================================================
// MAXT: largest block the variant may be launched with.  One block per SM is resident (the per-thread
// shared-memory columns fill the SM), so the register budget is 64K / MAXT: 512 -> 128, 448 -> 144,
// 384 -> 168.  The host picks the variant with the most registers that still covers its block size.
// SHARE: rollout prefix sharing (SfwShareDev): the same kernel simulates the shared paths and writes their
// per-step records (B.share.mode 1, 2) or starts every sample from the record of its fork point (mode 3).
// The plain instantiation compiles all of that out.
// SHARE == 2: the path writer with one WARP per path and one lane per pedestrian pair.  A lone thread walking all
// pairs of a path is latency bound (23 us per step with 20 pedestrians); here lane k evaluates only the
// interactions of its own pair -- the reactions from the pairs before it, then its actions on the pairs after it,
// with the very calls and in the very order the thread-per-trajectory pass uses -- so the records are bit-identical
// and a step costs what P2 pair evaluations cost, not P2^2 / 2.  Rollout and footprint are computed redundantly by
// every lane (uniform).  Scenes with pedestrian groups use the thread-per-path writer (SHARE == 1).
#define SFW_HUGE_TARGET 1.0e300 /* velocity target that keeps step_velocity saturated for ever */
template <int MAXT, int SHARE>
__global__ void __launch_bounds__(MAXT, 1)
sfw_score_small(const __grid_constant__ SfwBatchDev B, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool WARP = SHARE == 2;
  const uint32_t T = blockDim.x;
  const uint32_t tid = threadIdx.x;
  const uint32_t scene_local = blockIdx.x / B.tiles_per_scene;
  const uint32_t tile = blockIdx.x - scene_local * B.tiles_per_scene;
  const uint32_t scene = B.scene_base + scene_local; // every index below is the scene's GLOBAL one
  const SfwSceneDev *__restrict__ scp = B.scenes + scene;
  const uint32_t P2 = scp->n_pairs, M = scp->n_obst, F = scp->n_fp;
  const uint32_t n_groups = scp->n_groups;

  // ---- shared memory carve-up -------------------------------------------------------------
  // [window][pos P2][vel P2][goal P2][par P2][par2 P2][obst M][footprint F][mbar][tile best]
  // [per-thread columns: pos P2*T | vel P2*T | force P2*T]
  const uint32_t win_bytes = B.win_wp * B.win_h; // multiple of 16
  uint8_t *s_win = smem_raw;
  uint32_t off = (win_bytes + 127u) & ~127u;
  float4 *s_pos0 = reinterpret_cast<float4 *>(smem_raw + off);
  off += P2 * 16u;
  float4 *s_vel0 = reinterpret_cast<float4 *>(smem_raw + off);
  off += P2 * 16u;
  float4 *s_goal = reinterpret_cast<float4 *>(smem_raw + off);
  off += P2 * 16u;
  float4 *s_par = reinterpret_cast<float4 *>(smem_raw + off);
  off += P2 * 16u;
  float4 *s_par2 = reinterpret_cast<float4 *>(smem_raw + off);
  off += P2 * 16u;
  float2 *s_obs = reinterpret_cast<float2 *>(smem_raw + off);
  off += M * 8u; // M counts float2 slots: a multiple of 10
  double2 *s_fp = reinterpret_cast<double2 *>(smem_raw + off);
  off += F * 16u;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + off);
  off += 16u;
  float *s_redc = reinterpret_cast<float *>(smem_raw + off);
  uint32_t *s_redi = reinterpret_cast<uint32_t *>(smem_raw + off + 32u * 4u);
  off += 64u * 4u;
  // free-space bit maps over the window (row pass, then the final one)
  const uint32_t fw32 = (B.win_wp + 31u) >> 5;
  uint32_t *s_rowfree = reinterpret_cast<uint32_t *>(smem_raw + off);
  off += ((win_bytes ? B.win_h * fw32 * 4u : 0u) + 15u) & ~15u;
  uint32_t *s_free = reinterpret_cast<uint32_t *>(smem_raw + off);
  off += ((win_bytes ? B.win_h * fw32 * 4u : 0u) + 15u) & ~15u;
  float4 *s_pos = reinterpret_cast<float4 *>(smem_raw + off);
  off += P2 * T * 16u;
  float4 *s_vel = reinterpret_cast<float4 *>(smem_raw + off);
  off += P2 * T * 16u;
  float4 *s_frc = reinterpret_cast<float4 *>(smem_raw + off);

  __shared__ bool s_last;

  // ---- stage the scene with the TMA engine ---------------------------------------------------
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
    const uint32_t obs_bytes = M * 8u;
    const uint32_t tx = win_bytes + 5u * P2 * 16u + obs_bytes + F * 16u;
    mbar_expect_tx(s_bar, tx);
    if (win_bytes)
      tma_load_3d(s_win, &tmap, scp->win_x0, scp->win_y0, (int)scene, s_bar);
    if (P2) {
      bulk_g2s(s_pos0, B.pedPos + scp->ped_off, P2 * 16u, s_bar);
      bulk_g2s(s_vel0, B.pedVel + scp->ped_off, P2 * 16u, s_bar);
      bulk_g2s(s_goal, B.pedGoal + scp->ped_off, P2 * 16u, s_bar);
      bulk_g2s(s_par, B.pedPar + scp->ped_off, P2 * 16u, s_bar);
      bulk_g2s(s_par2, B.pedPar2 + scp->ped_off, P2 * 16u, s_bar);
    }
    if (obs_bytes)
      bulk_g2s(s_obs, B.obst + scp->obs_off, obs_bytes, s_bar);
    if (F)
      bulk_g2s(s_fp, B.footprint + scp->fp_off, F * 16u, s_bar);
  }
  __syncthreads(); // barrier init visible to the waiters
  mbar_wait(s_bar, 0);

  // ---- free-space map of the window: bit (lx, ly) set <=> all cells within fp_rc of it are inside the
  // map and cost 0.  Separable: a row pass (ballot per 32 cells), then an AND over 2 fp_rc + 1 rows.
  const uint32_t fp_rc = scp->fp_rc;
  const bool use_free = win_bytes != 0u && fp_rc != 0u && F >= 3u;
  if (use_free) {
    const int r = (int)fp_rc, ww = (int)B.win_wp, wh = (int)B.win_h;
    const int gx0 = scp->win_x0, gy0 = scp->win_y0, sx = (int)scp->size_x, sy = (int)scp->size_y;
    const uint32_t n_words = B.win_h * fw32;
    for (uint32_t wd = tid >> 5; wd < n_words; wd += T >> 5) { // T is a multiple of 32
      const int ly = (int)(wd / fw32), lx = (int)((wd % fw32) * 32u + (tid & 31u));
      bool ok = lx - r >= 0 && lx + r < ww && gx0 + lx - r >= 0 && gx0 + lx + r < sx && gy0 + ly >= 0 && gy0 + ly < sy;
      if (ok) {
        const uint8_t *row = s_win + ly * ww + lx;
        uint32_t acc = 0;
        for (int d = -r; d <= r; ++d)
          acc |= row[d];
        ok = acc == 0u;
      }
      const uint32_t bits = __ballot_sync(0xffffffffu, ok);
      if ((tid & 31u) == 0)
        s_rowfree[wd] = bits;
    }
    __syncthreads();
    for (uint32_t wd = tid; wd < n_words; wd += T) {
      const int ly = (int)(wd / fw32);
      uint32_t bits = 0u;
      if (ly - r >= 0 && ly + r < wh) {
        bits = 0xffffffffu;
        for (int d = -r; d <= r; ++d)
          bits &= s_rowfree[wd + d * (int)fw32];
      }
      s_free[wd] = bits;
    }
    __syncthreads();
  }

  MapView mv;
  mv.win = win_bytes ? s_win : nullptr;
  mv.glob = B.maps + scp->map_off;
  mv.ox = scp->origin_x;
  mv.oy = scp->origin_y;
  mv.res = scp->resolution;
  mv.rinv = __ddiv_rn(1.0, scp->resolution);
  mv.sx = scp->size_x;
  mv.sy = scp->size_y;
  mv.pitch = B.map_pitch;
  mv.wx0 = scp->win_x0;
  mv.wy0 = scp->win_y0;
  mv.wwp = win_bytes ? B.win_wp : 0u;
  mv.wh = win_bytes ? B.win_h : 0u;
  mv.free_bits = use_free ? s_free : nullptr;
  mv.fw32 = fw32;

  const SfmConst K = make_sfm_const(B);

  // ---- which trajectory is mine ----------------------------------------------------------------
  const uint32_t n_w = B.n_w;
  const uint32_t first = B.row_begin * n_w, last = B.row_end * n_w;
  uint32_t idx = first + tile * T + tid;
  bool in_range = idx < last;
  double v_s = 0.0, w_s = 0.0;
  int s0 = 0;                           // first step this thread simulates itself
  const uint8_t *ck_in = nullptr;       // record it starts from (nullptr: the scene's initial state)
  uint8_t *ck_out = nullptr;            // writer launches: this path's records [step count]
  uint32_t share_mode = SHARE ? B.share.mode : 0u;
  uint32_t ptile = tile;
  const bool merged = WARP && share_mode == 4u; // both path stages in this launch
  if (merged) {
    share_mode = tile == 0u ? 1u : 2u;
    ptile = tile ? tile - 1u : 0u;
  }
  const bool writer = SHARE && (share_mode == 1u || share_mode == 2u);
  if (SHARE && share_mode) {
    const SfwShareDev &H = B.share;
    const uint16_t *kv = H.kv + (size_t)scene * B.n_v, *kw = H.kw + (size_t)scene * n_w;
    const uint8_t *dirv = H.dirv + (size_t)scene * B.n_v, *dirw = H.dirw + (size_t)scene * n_w;
    uint8_t *base = H.records + (size_t)scene * H.scene_stride;
    const size_t R = H.rec_bytes, K1 = (size_t)H.kmax + 1u;
    const uint32_t p = WARP ? ptile * (T >> 5) + (tid >> 5) : tile * T + tid; // writer launches: path number
    if (share_mode == 3u) {
      // sorted position of this thread among the samples of the row slab [row_begin, row_end) (the whole grid
      // unless the caller shards rows over GPUs: then row_perm / lvl_rows describe the slab's rows only)
      const uint32_t n_slab = last - first;
      uint32_t pos = tile * T + tid;
      if (H.chunk_map) {
        const uint32_t ch = pos >> 5;
        pos = ch < ((n_slab + 31u) >> 5) ? H.chunk_map[ch] * 32u + (pos & 31u) : n_slab;
      }
      in_range = pos < n_slab;
      if (in_range) {
        // level of pos: the largest k with lvl_rows[k] * lvl_cols[k] <= pos
        const uint32_t *lvl_rows = H.lvl_rows + (size_t)scene * (H.kmax + 2u);
        const uint32_t *lvl_cols = H.lvl_cols + (size_t)scene * (H.kmax + 2u);
        const uint32_t *row_perm = H.row_perm + (size_t)scene * B.n_v, *col_perm = H.col_perm + (size_t)scene * n_w;
        uint32_t lo = 0u, hi = H.kmax + 1u;
        while (hi - lo > 1u) {
          const uint32_t mid = (lo + hi) >> 1;
          if (lvl_rows[mid] * lvl_cols[mid] <= pos)
            lo = mid;
          else
            hi = mid;
        }
        const uint32_t r0 = lvl_rows[lo], r1 = lvl_rows[lo + 1u], c0 = lvl_cols[lo], c1 = lvl_cols[lo + 1u];
        const uint32_t local = pos - r0 * c0, n_a = (r1 - r0) * c1;
        uint32_t r, c;
        if (local < n_a) { // rows of this level x columns up to this level
          const uint32_t q = local / c1;
          r = row_perm[r0 + q];
          c = col_perm[local - q * c1];
        } else { // earlier rows x columns of this level
          const uint32_t l2 = local - n_a, nc = c1 - c0, q = l2 / nc;
          r = row_perm[q];
          c = col_perm[c0 + (l2 - q * nc)];
        }
        idx = r * n_w + c;
        v_s = B.linvels[r];
        w_s = B.angvels[c];
        const int kvr = kv[r], kwc = kw[c];
        s0 = max(kvr, kwc);
        if (s0 > 0) {
          size_t path;
          if (kvr == kwc)
            path = (size_t)dirv[r] * 2u + dirw[c];
          else if (kvr > kwc)
            path = 4u + (size_t)dirv[r] * n_w + c;
          else
            path = 4u + 2u * (size_t)n_w + (size_t)r * 2u + dirw[c];
          ck_in = base + (path * K1 + (size_t)s0) * R;
        }
      }
    } else if (share_mode == 1u) { // the 4 doubly saturated paths
      in_range = p < 4u;
      if (in_range) {
        v_s = (p >> 1) ? SFW_HUGE_TARGET : -SFW_HUGE_TARGET;
        w_s = (p & 1u) ? SFW_HUGE_TARGET : -SFW_HUGE_TARGET;
        ck_out = base + (size_t)p * K1 * R;
      }
    } else { // mode 2: (v saturated, column c) and (row r, w saturated)
      in_range = p < 2u * n_w + 2u * B.n_v;
      if (in_range) {
        size_t parent;
        if (p < 2u * n_w) {
          const uint32_t dv = p / n_w, c = p - dv * n_w;
          v_s = dv ? SFW_HUGE_TARGET : -SFW_HUGE_TARGET;
          w_s = B.angvels[c];
          s0 = kw[c];
          parent = (size_t)dv * 2u + dirw[c];
        } else {
          const uint32_t q = p - 2u * n_w, r = q >> 1, dw = q & 1u;
          v_s = B.linvels[r];
          w_s = dw ? SFW_HUGE_TARGET : -SFW_HUGE_TARGET;
          s0 = kv[r];
          parent = (size_t)dirv[r] * 2u + dw;
        }
        if (s0 > 0)
          ck_in = base + (parent * K1 + (size_t)s0) * R;
        ck_out = base + (4u + (size_t)p) * K1 * R;
      }
    }
  } else if (in_range) {
    v_s = B.linvels[idx / n_w];
    w_s = B.angvels[idx % n_w];
  }
  const bool skipped = !writer && in_range && !B.score_zero && (v_s == 0.0 && w_s == 0.0); // sfw_planner.cpp:349-352
  bool alive = in_range && !skipped;

  // ---- per-thread state: one shared-memory column per thread (conflict-free LDS.128) ------------
  // (one column per warp in the warp-per-path writer: every lane stores the same values)
  const uint32_t col = WARP ? (tid & ~31u) : tid;
  float4 *pos = s_pos + col, *vel = s_vel + col, *frc = s_frc + col; // [k * T]
  for (uint32_t k = WARP ? (tid & 31u) : 0u; k < P2; k += WARP ? 32u : 1u) { // warp writer: lane k owns pair k
    pos[k * T] = s_pos0[k];
    vel[k * T] = s_vel0[k];
    if (!WARP)
      frc[k * T] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  uint64_t goalmask = scp->goal_mask;

  double x = scp->rx, y = scp->ry, th = scp->rth;
  double vx = scp->rvx, vth = scp->rvth;
  const double vy = scp->rvy; // acc_y == 0: vy never changes (sfw_planner.cpp:357,582)
  const double base_x = scp->rx, base_y = scp->ry;
  const double dt = B.dt;
  const double ax_dt = __dmul_rn(B.acc_x, dt), ath_dt = __dmul_rn(B.acc_th, dt);
  const float dtf = B.dtf;
  // SFM view of the robot (agents[0]): starts as the sensor-interface snapshot
  float prx = scp->ax, pry = scp->ay, rvxf = scp->avx, rvyf = scp->avy;
  const float a_obs_scale = scp->a_obs_scale;
  const float rr2 = B.rr2;
  double social_work = 0.0, costmap_sum = 0.0;
  int npts = 0;
  const int S = B.num_steps;
  if (SHARE) {
    // Everything above — the TMA staging, the free-space map, the fork tables — only read the caller's inputs.
    // The shared-path records below come from the PREVIOUS launch of this tick: with programmatic dependent
    // launch this kernel's prologue overlapped that launch's tail, and here it waits for it.
    if (writer)
      griddep_launch_dependents();
    griddep_wait();
  }
  if (WARP && merged && ck_in) {
    // the record this path continues from is written by tile 0 of the scene in this very launch (a lower block
    // index, hence already dispatched): wait for its flag
    const uint32_t *flag = &reinterpret_cast<const SfwCkptHdr *>(ck_in)->epoch;
    uint32_t seen;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
      if (seen == B.share.epoch)
        break;
      __nanosleep(256);
      if (clock64() - t0 > (1ll << 33)) { // seconds: the host only merges launches whose blocks are all co-resident
        // give up WITHOUT killing the context: this path (and what forks from it) is garbage, the host sees the
        // status word after the stream synchronises and fails the call with SFW_ERR_STATE
        if ((tid & 31u) == 0u) {
          *reinterpret_cast<volatile unsigned int *>(B.status) = SFW_DEVSTAT_PATH_WAIT; // mapped pinned host word
          __threadfence_system();
        }
        in_range = false;
        ck_in = nullptr;
        break;
      }
    }
  }
  if (SHARE && ck_in) { // start from the shared path's record of this thread's fork point
    // (the warp-per-path writer may be reading what another block wrote moments ago: L2 loads only)
    const SfwCkptHdr *h = reinterpret_cast<const SfwCkptHdr *>(ck_in);
    auto ld = [](const auto *q) { return WARP ? __ldcg(q) : *q; };
    x = ld(&h->x);
    y = ld(&h->y);
    th = ld(&h->th);
    vx = ld(&h->vx);
    vth = ld(&h->vth);
    social_work = ld(&h->social_work);
    costmap_sum = ld(&h->costmap_sum);
    prx = ld(&h->prx);
    pry = ld(&h->pry);
    rvxf = ld(&h->rvxf);
    rvyf = ld(&h->rvyf);
    goalmask = ld(reinterpret_cast<const unsigned long long *>(&h->goalmask));
    npts = ld(&h->npts);
    alive = alive && ld(&h->alive) != 0;
    const float4 *pv = reinterpret_cast<const float4 *>(ck_in + sizeof(SfwCkptHdr));
    for (uint32_t k = WARP ? (tid & 31u) : 0u; k < P2; k += WARP ? 32u : 1u) {
      pos[k * T] = ld(&pv[2u * k]);
      vel[k * T] = ld(&pv[2u * k + 1u]);
    }
  }
  if (WARP)
    __syncwarp(); // the column is shared by the warp
  const int S_end = writer ? (int)B.share.kmax : S;
  const int i_first = SHARE ? __reduce_min_sync(0xffffffffu, in_range ? s0 : S_end) : 0;

  for (int i = i_first; i < S_end; ++i) {
    if (!writer && !__any_sync(0xffffffffu, alive))
      break;
    const bool started = !SHARE || i >= s0; // lanes of a warp may fork at different steps
    const bool was_alive = alive;
    alive = alive && started;
    // -- legality of the current pose (sfw_planner.cpp:545-575) --
    double sn = 0.0, cs = 1.0;
    int fc = -1;
    if (alive) {
      sincos(th, &sn, &cs);
      fc = footprint_cost(mv, s_fp, (int)F, x, y, sn, cs);
    }
    // The Bresenham loops above leave the warp split; without this the whole social-force step
    // below would run once per fragment (measured: 19.8 of 32 lanes active, profiles/r1a).
    __syncwarp();
    if (fc < 0)
      alive = false;
    if (alive) {
      {
        costmap_sum = __dadd_rn(costmap_sum, __ddiv_rn((double)fc, 255.0));
        ++npts;
        // -- velocities and pose (sfw_planner.cpp:581-588, hpp:418-463) --
        vx = step_velocity(v_s, vx, ax_dt);
        vth = step_velocity(w_s, vth, ath_dt);
        double lx = __dmul_rn(vx, cs), ly = __dmul_rn(vx, sn);
        if (vy != 0.0) {
          double sn2, cs2;
          sincos(__dadd_rn(1.57079632679489661923, th), &sn2, &cs2);
          lx = __dadd_rn(lx, __dmul_rn(vy, cs2));
          ly = __dadd_rn(ly, __dmul_rn(vy, sn2));
        }
        x = __dadd_rn(x, __dmul_rn(lx, dt));
        y = __dadd_rn(y, __dmul_rn(ly, dt));
        th = __dadd_rn(th, __dmul_rn(vth, dt));
        const float nrx = (float)(x - base_x), nry = (float)(y - base_y);

        // -- social force step (sfw_planner.cpp:592-629), pedestrians processed as pairs --
        f2 rfx2 = bc2(0.f), rfy2 = bc2(0.f), wp2 = bc2(0.f); // robot force / social-work partial sums
        bool hit = false;
        const f2 RX = bc2(prx), RY = bc2(pry), RVX = bc2(rvxf), RVY = bc2(rvyf);
        const f2 DT = bc2(dtf);
        if constexpr (WARP) {
          const uint32_t k = tid & 31u; // my pedestrian pair
          const bool mine = k < P2;
          const float4 pa = mine ? pos[k * T] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 va = mine ? vel[k * T] : make_float4(0.f, 0.f, 0.f, 0.f);
          const f2 AX = mk2(pa.x, pa.y), AY = mk2(pa.z, pa.w);
          const f2 AVX = mk2(va.x, va.y), AVY = mk2(va.z, va.w);
          f2 fx = bc2(0.f), fy = bc2(0.f), fm = bc2(0.f);
          f2 FX = bc2(0.f), FY = bc2(0.f);                                        // reactions from earlier pairs
          f2 s0x = bc2(0.f), s0y = bc2(0.f), s1x = bc2(0.f), s1y = bc2(0.f);      // actions on later pairs
          // Cross-pair forces: the P2 (P2 - 1) / 2 unordered tiles (pair a < pair b) are DEALT over all 32 lanes
          // — a lane per pair would walk P2 - 1 tiles one after the other, and two thirds of the warp would watch —
          // each evaluated exactly as the thread-per-trajectory pass evaluates it at iteration a for target b
          // (a0 <- (b0, b1), a1 <- (b0, b1)); the four packed results go to the warp's share of the (otherwise
          // unused) force columns, 16 tiles per row.  Lane k then REPLAYS its pair's sums in that pass's order:
          // the reactions of the pairs before it, the actions on the pairs after it — same operands, same order,
          // bit-identical records.
          {
            const uint32_t lane = tid & 31u, n_tiles = P2 * (P2 - 1u) / 2u;
            for (uint32_t t = lane; t < n_tiles; t += 32u) {
              uint32_t a = 0u, rem = t; // tile t = (a, b): row a of the upper triangle holds P2 - 1 - a tiles
              while (rem >= P2 - 1u - a) {
                rem -= P2 - 1u - a;
                ++a;
              }
              const uint32_t b = a + 1u + rem;
              const float4 pA = pos[a * T], vA = vel[a * T], pB = pos[b * T], vB = vel[b * T];
              const f2 BX = mk2(pB.x, pB.y), BY = mk2(pB.z, pB.w);
              const f2 BVX = mk2(vB.x, vB.y), BVY = mk2(vB.z, vB.w);
              f2 hx, hy, gx2, gy2, hm;
              pair_force2<false>(K, bc2(pA.x), bc2(pA.z), bc2(vA.x), bc2(vA.z), BX, BY, BVX, BVY, hx, hy, hm);
              pair_force2<false>(K, bc2(pA.y), bc2(pA.w), bc2(vA.y), bc2(vA.w), BX, BY, BVX, BVY, gx2, gy2, hm);
              float4 *slot = frc + (t >> 4) * T + 2u * (t & 15u);
              float l, h, l2, h2;
              un2(hx, l, h);
              un2(hy, l2, h2);
              slot[0] = make_float4(l, h, l2, h2);
              un2(gx2, l, h);
              un2(gy2, l2, h2);
              slot[1] = make_float4(l, h, l2, h2);
            }
            __syncwarp();
            if (mine) {
              for (uint32_t j = 0; j < P2; ++j) {
                if (j == k)
                  continue;
                const uint32_t a = j < k ? j : k, b = j < k ? k : j;
                const uint32_t t = a * P2 - a * (a + 1u) / 2u + (b - a - 1u);
                const float4 *slot = frc + (t >> 4) * T + 2u * (t & 15u);
                const float4 h4 = slot[0], g4 = slot[1];
                const f2 hx = mk2(h4.x, h4.y), hy = mk2(h4.z, h4.w), gx2 = mk2(g4.x, g4.y), gy2 = mk2(g4.z, g4.w);
                if (j > k) { // pair k acted on pair j
                  s0x = add2(s0x, hx);
                  s0y = add2(s0y, hy);
                  s1x = add2(s1x, gx2);
                  s1y = add2(s1y, gy2);
                } else { // the reaction of pair j's action on pair k
                  FX = sub2(FX, add2(hx, gx2));
                  FY = sub2(FY, add2(hy, gy2));
                }
              }
            }
          }
          float npx0 = 0.f, npx1 = 0.f, npy0 = 0.f, npy1 = 0.f, nvx0 = 0.f, nvx1 = 0.f, nvy0 = 0.f, nvy1 = 0.f;
          uint32_t clr_lo = 0u, clr_hi = 0u;
          bool myhit = false;
          if (mine) {
            pair_force2<true>(K, AX, AY, AVX, AVY, RX, RY, RVX, RVY, fx, fy, fm);
            FX = add2(FX, fx);
            FY = add2(FY, fy);
            {
              float gx_, gy_, gm_;
              pair_force<false>(K, pa.x, pa.z, va.x, va.z, pa.y, pa.w, va.y, va.w, gx_, gy_, gm_);
              FX = add2(FX, mk2(gx_, -gx_));
              FY = add2(FY, mk2(gy_, -gy_));
            }
            {
              float l0, h0, l1, h1;
              un2(s0x, l0, h0);
              un2(s1x, l1, h1);
              FX = add2(FX, mk2(l0 + h0, l1 + h1));
              un2(s0y, l0, h0);
              un2(s1y, l1, h1);
              FY = add2(FY, mk2(l0 + h0, l1 + h1));
            }
            f2 ox, oy;
            obstacle_sum2(s_obs, (int)M, B.c_obs, AX, AY, ox, oy);
            const float4 G = s_goal[k], Pp = s_par[k], Pc = s_par2[k];
            const f2 gdx = sub2(mk2(G.x, G.y), AX), gdy = sub2(mk2(G.z, G.w), AY);
            float g20, g21;
            un2(fma2(gdx, gdx, mul2(gdy, gdy)), g20, g21);
            const bool hg0 = (goalmask >> (2u * k)) & 1ull, hg1 = (goalmask >> (2u * k + 1u)) & 1ull;
            const bool go0 = hg0 && g20 > Pp.x, go1 = hg1 && g21 > Pp.y;
            const f2 gs = mk2(go0 ? rsqrt_approx(g20) * Pp.z : 0.f, go1 ? rsqrt_approx(g21) * Pp.w : 0.f);
            const f2 c1 = mk2(go0 ? B.kd_tau : B.inv_tau, go1 ? B.kd_tau : B.inv_tau);
            const f2 dfx = mul2(c1, sub2(mul2(gdx, gs), AVX));
            const f2 dfy = mul2(c1, sub2(mul2(gdy, gs), AVY));
            const f2 OS = mk2(Pc.x, Pc.y);
            const f2 Fx = add2(add2(dfx, FX), mul2(OS, ox));
            const f2 Fy = add2(add2(dfy, FY), mul2(OS, oy));
            f2 nvx = fma2(Fx, DT, AVX), nvy = fma2(Fy, DT, AVY);
            float v20, v21;
            un2(fma2(nvx, nvx, mul2(nvy, nvy)), v20, v21);
            const f2 sc = mk2(v20 > Pc.z ? Pp.z * rsqrt_approx(v20) : 1.0f, v21 > Pc.w ? Pp.w * rsqrt_approx(v21) : 1.0f);
            nvx = mul2(nvx, sc);
            nvy = mul2(nvy, sc);
            const f2 npx = fma2(nvx, DT, AX), npy = fma2(nvy, DT, AY);
            un2(npx, npx0, npx1);
            un2(npy, npy0, npy1);
            un2(nvx, nvx0, nvx1);
            un2(nvy, nvy0, nvy1);
            {
              const f2 hx = sub2(mk2(G.x, G.y), npx), hy = sub2(mk2(G.z, G.w), npy);
              float h0, h1;
              un2(fma2(hx, hx, mul2(hy, hy)), h0, h1);
              uint64_t clr = 0ull;
              if (hg0 && h0 <= Pp.x)
                clr |= 1ull << (2u * k);
              if (hg1 && h1 <= Pp.y)
                clr |= 1ull << (2u * k + 1u);
              clr_lo = (uint32_t)clr;
              clr_hi = (uint32_t)(clr >> 32);
            }
            {
              const f2 cx = sub2(bc2(nrx), npx), cy = sub2(bc2(nry), npy);
              float c0, c1_;
              un2(fma2(cx, cx, mul2(cy, cy)), c0, c1_);
              myhit = (c0 <= rr2) | (c1_ <= rr2);
            }
          }
          __syncwarp(); // every lane has read the old state
          if (mine) {
            pos[k * T] = make_float4(npx0, npx1, npy0, npy1);
            vel[k * T] = make_float4(nvx0, nvx1, nvy0, nvy1);
          }
          // ordered sums over the pairs, as the one-thread pass accumulates them
          {
            float fxl, fxh, fyl, fyh, fml, fmh;
            un2(fx, fxl, fxh);
            un2(fy, fyl, fyh);
            un2(fm, fml, fmh);
            for (uint32_t q = 0; q < P2; ++q) {
              rfx2 = sub2(rfx2, mk2(__shfl_sync(0xffffffffu, fxl, q), __shfl_sync(0xffffffffu, fxh, q)));
              rfy2 = sub2(rfy2, mk2(__shfl_sync(0xffffffffu, fyl, q), __shfl_sync(0xffffffffu, fyh, q)));
              wp2 = add2(wp2, mk2(__shfl_sync(0xffffffffu, fml, q), __shfl_sync(0xffffffffu, fmh, q)));
            }
          }
          hit = __any_sync(0xffffffffu, myhit);
          goalmask &= ~((uint64_t)__reduce_or_sync(0xffffffffu, clr_lo) |
                        ((uint64_t)__reduce_or_sync(0xffffffffu, clr_hi) << 32));
          __syncwarp(); // new state visible to the warp
        } else {
        // group forces of computeForces (lightsfm computeGroupForce): added to the (zeroed) accumulators
        if (n_groups) {
          const uint32_t *gt = B.groups + scp->grp_off;
          const uint32_t *mem0 = gt + n_groups + 1u;
          const float *Pf = reinterpret_cast<const float *>(pos);
          float *Ff = reinterpret_cast<float *>(frc);
          const uint32_t st = 4u * T;
          for (uint32_t g = 0; g < n_groups; ++g) {
            const uint32_t s0 = gt[g], c = gt[g + 1u] - s0;
            const uint32_t *mem = mem0 + 2u * s0;
            float cx, cy;
            group_centre(mem, c, Pf, st, cx, cy);
            for (uint32_t m = 0; m < c; ++m) {
              const uint32_t j = mem[2u * m];
              float ddx, ddy, gfx, gfy;
              desired_direction(Pf, st, s_goal, s_par, j, (goalmask >> j) & 1ull, ddx, ddy);
              group_member_force(mem, c, m, Pf, st, cx, cy, ddx, ddy, B.k_gaze, B.k_coh, B.k_rep, gfx, gfy);
              Ff[(j >> 1) * st + (j & 1u)] += gfx;
              Ff[(j >> 1) * st + 2u + (j & 1u)] += gfy;
            }
          }
        }
        for (uint32_t k = 0; k < P2; ++k) {
          const float4 pa = pos[k * T], va = vel[k * T], fa4 = frc[k * T];
          const f2 AX = mk2(pa.x, pa.y), AY = mk2(pa.z, pa.w);
          const f2 AVX = mk2(va.x, va.y), AVY = mk2(va.z, va.w);
          f2 FX = mk2(fa4.x, fa4.y), FY = mk2(fa4.z, fa4.w);
          f2 fx, fy, fm;
          // both pedestrians of the pair <- robot (robot at the pose the previous step left it)
          pair_force2<true>(K, AX, AY, AVX, AVY, RX, RY, RVX, RVY, fx, fy, fm);
          FX = add2(FX, fx);
          FY = add2(FY, fy);
          rfx2 = sub2(rfx2, fx);
          rfy2 = sub2(rfy2, fy);
          wp2 = add2(wp2, fm); // = computeSocialWork's per-pedestrian terms of the PREVIOUS step
          // the two pedestrians of the pair against each other
          {
            float gx_, gy_, gm_;
            pair_force<false>(K, pa.x, pa.z, va.x, va.z, pa.y, pa.w, va.y, va.w, gx_, gy_, gm_);
            FX = add2(FX, mk2(gx_, -gx_));
            FY = add2(FY, mk2(gy_, -gy_));
          }
          // against every later pair: a0 <- (b0, b1) and a1 <- (b0, b1), antisymmetric push to b
          const f2 A0X = bc2(pa.x), A0Y = bc2(pa.z), A0VX = bc2(va.x), A0VY = bc2(va.z);
          const f2 A1X = bc2(pa.y), A1Y = bc2(pa.w), A1VX = bc2(va.y), A1VY = bc2(va.w);
          f2 s0x = bc2(0.f), s0y = bc2(0.f), s1x = bc2(0.f), s1y = bc2(0.f);
#if SFW_SMALL_PREFETCH
          // the next pair's state is fetched one trip ahead: the reaction push at the end of a trip is a shared-memory
          // store the compiler must keep in order with every later shared load, which would put the load latency of
          // pos / vel at the head of every trip
          // (one array past the last pair on the final trip: pos runs into vel, vel into frc — in bounds, unused)
          uint32_t o = (k + 1u) * T;
          float4 pb = pos[o], vb = vel[o];
          for (uint32_t j = k + 1; j < P2; ++j, o += T) {
            const float4 pbn = pos[o + T], vbn = vel[o + T];
#else
          for (uint32_t j = k + 1; j < P2; ++j) {
            const float4 pb = pos[j * T], vb = vel[j * T];
#endif
            const f2 BX = mk2(pb.x, pb.y), BY = mk2(pb.z, pb.w);
            const f2 BVX = mk2(vb.x, vb.y), BVY = mk2(vb.z, vb.w);
            f2 hx, hy, gx2, gy2, hm;
            pair_force2<false>(K, A0X, A0Y, A0VX, A0VY, BX, BY, BVX, BVY, hx, hy, hm);
            pair_force2<false>(K, A1X, A1Y, A1VX, A1VY, BX, BY, BVX, BVY, gx2, gy2, hm);
            s0x = add2(s0x, hx);
            s0y = add2(s0y, hy);
            s1x = add2(s1x, gx2);
            s1y = add2(s1y, gy2);
#if SFW_SMALL_PREFETCH
            float4 *fb = frc + o;
#else
            float4 *fb = frc + j * T;
#endif
            const float4 fb4 = *fb;
            float bx0, bx1, by0, by1;
            un2(sub2(mk2(fb4.x, fb4.y), add2(hx, gx2)), bx0, bx1);
            un2(sub2(mk2(fb4.z, fb4.w), add2(hy, gy2)), by0, by1);
            *fb = make_float4(bx0, bx1, by0, by1);
#if SFW_SMALL_PREFETCH
            pb = pbn;
            vb = vbn;
#endif
          }
          {
            float l0, h0, l1, h1;
            un2(s0x, l0, h0);
            un2(s1x, l1, h1);
            FX = add2(FX, mk2(l0 + h0, l1 + h1));
            un2(s0y, l0, h0);
            un2(s1y, l1, h1);
            FY = add2(FY, mk2(l0 + h0, l1 + h1));
          }
          // obstacle force
          f2 ox, oy;
          obstacle_sum2(s_obs, (int)M, B.c_obs, AX, AY, ox, oy);
          const float4 G = s_goal[k], Pp = s_par[k], Pc = s_par2[k];
          // desired force (App. B-1): k_des/tau * (e_goal * v_des - v) with a live goal, else -v/tau
          const f2 gdx = sub2(mk2(G.x, G.y), AX), gdy = sub2(mk2(G.z, G.w), AY);
          float g20, g21;
          un2(fma2(gdx, gdx, mul2(gdy, gdy)), g20, g21);
          const bool hg0 = (goalmask >> (2u * k)) & 1ull, hg1 = (goalmask >> (2u * k + 1u)) & 1ull;
          const bool go0 = hg0 && g20 > Pp.x, go1 = hg1 && g21 > Pp.y;
          const f2 gs = mk2(go0 ? rsqrt_approx(g20) * Pp.z : 0.f, go1 ? rsqrt_approx(g21) * Pp.w : 0.f);
          const f2 c1 = mk2(go0 ? B.kd_tau : B.inv_tau, go1 ? B.kd_tau : B.inv_tau);
          const f2 dfx = mul2(c1, sub2(mul2(gdx, gs), AVX));
          const f2 dfy = mul2(c1, sub2(mul2(gdy, gs), AVY));
          const f2 OS = mk2(Pc.x, Pc.y);
          const f2 Fx = add2(add2(dfx, FX), mul2(OS, ox));
          const f2 Fy = add2(add2(dfy, FY), mul2(OS, oy));
          // updatePosition (App. B-5)
          f2 nvx = fma2(Fx, DT, AVX), nvy = fma2(Fy, DT, AVY);
          float v20, v21;
          un2(fma2(nvx, nvx, mul2(nvy, nvy)), v20, v21);
          const f2 sc = mk2(v20 > Pc.z ? Pp.z * rsqrt_approx(v20) : 1.0f, v21 > Pc.w ? Pp.w * rsqrt_approx(v21) : 1.0f);
          nvx = mul2(nvx, sc);
          nvy = mul2(nvy, sc);
          const f2 npx = fma2(nvx, DT, AX), npy = fma2(nvy, DT, AY);
          float px0, px1, py0, py1, vx0, vx1, vy0, vy1;
          un2(npx, px0, px1);
          un2(npy, py0, py1);
          un2(nvx, vx0, vx1);
          un2(nvy, vy0, vy1);
          pos[k * T] = make_float4(px0, px1, py0, py1);
          vel[k * T] = make_float4(vx0, vx1, vy0, vy1);
          frc[k * T] = make_float4(0.f, 0.f, 0.f, 0.f);
          // goal reached -> pop (App. B-5)
          {
            const f2 hx = sub2(mk2(G.x, G.y), npx), hy = sub2(mk2(G.z, G.w), npy);
            float h0, h1;
            un2(fma2(hx, hx, mul2(hy, hy)), h0, h1);
            if (hg0 && h0 <= Pp.x)
              goalmask &= ~(1ull << (2u * k));
            if (hg1 && h1 <= Pp.y)
              goalmask &= ~(1ull << (2u * k + 1u));
          }
          // robot / pedestrian collision with the NEW robot pose (sfw_planner.cpp:613-627)
          {
            const f2 cx = sub2(bc2(nrx), npx), cy = sub2(bc2(nry), npy);
            float c0, c1_;
            un2(fma2(cx, cx, mul2(cy, cy)), c0, c1_);
            hit |= (c0 <= rr2) | (c1_ <= rr2);
          }
        }
        } // !WARP
        float rfx, rfy, wp;
        {
          float l, h;
          un2(rfx2, l, h);
          rfx = l + h;
          un2(rfy2, l, h);
          rfy = l + h;
          un2(wp2, l, h);
          wp = l + h;
        }
        // robot's own obstacle force at the pose the forces were evaluated at
        float rox, roy;
        obstacle_sum1(s_obs, (int)M, B.c_obs, prx, pry, rox, roy);
        rox *= a_obs_scale;
        roy *= a_obs_scale;
        const float wr = sqrt_approx(fmaf(rfx, rfx, rfy * rfy)) + sqrt_approx(fmaf(rox, rox, roy * roy));
        // wp accumulated at step i belongs to step i-1; the i == 0 pass must not count
        social_work += (double)(wr + ((i > 0) ? wp : 0.f));
        prx = nrx;
        pry = nry;
        rvxf = (float)vx;
        rvyf = (float)vy;
        if (hit)
          alive = false;
      }
    }
    if (SHARE) {
      if (!started)
        alive = was_alive; // not forked yet: untouched
      if (writer && in_range && started) { // the path's state after i + 1 steps
        uint8_t *rec = ck_out + (size_t)(i + 1) * B.share.rec_bytes;
        SfwCkptHdr *h = reinterpret_cast<SfwCkptHdr *>(rec);
        h->x = x;
        h->y = y;
        h->th = th;
        h->vx = vx;
        h->vth = vth;
        h->social_work = social_work;
        h->costmap_sum = costmap_sum;
        h->prx = prx;
        h->pry = pry;
        h->rvxf = rvxf;
        h->rvyf = rvyf;
        h->goalmask = goalmask;
        h->npts = npts;
        h->alive = alive ? 1 : 0;
        if (alive) {
          float4 *pv = reinterpret_cast<float4 *>(rec + sizeof(SfwCkptHdr));
          for (uint32_t k = WARP ? (tid & 31u) : 0u; k < P2; k += WARP ? 32u : 1u) {
            pv[2u * k] = pos[k * T];
            pv[2u * k + 1u] = vel[k * T];
          }
        }
        if (WARP && merged && share_mode == 1u) { // publish: every lane's stores, then the flag
          __threadfence();
          __syncwarp();
          if ((tid & 31u) == 0u)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&h->epoch), "r"(B.share.epoch) : "memory");
        }
      }
    }
  }
  if (writer)
    return; // shared paths have no cost of their own

  // ---- terminal costs (sfw_planner.cpp:643-675) -----------------------------------------------
  float cost = in_range ? (skipped ? SFW_COST_SKIPPED : SFW_COST_INVALID) : SFW_COST_SKIPPED;
  if (alive) {
    // computeSocialWork's pedestrian term of the last step (updated states, robot at final pose)
    f2 wp2 = bc2(0.f);
    const f2 RX = bc2(prx), RY = bc2(pry), RVX = bc2(rvxf), RVY = bc2(rvyf);
    for (uint32_t k = 0; k < P2; ++k) {
      const float4 pa = pos[k * T], va = vel[k * T];
      f2 fx, fy, fm;
      pair_force2<true>(K, mk2(pa.x, pa.y), mk2(pa.z, pa.w), mk2(va.x, va.y), mk2(va.z, va.w), RX, RY, RVX,
                        RVY, fx, fy, fm);
      wp2 = add2(wp2, fm);
    }
    float wp, wph;
    un2(wp2, wp, wph);
    wp += wph;
    social_work += (double)wp;
    const double dx = __dsub_rn(scp->wpx, x), dy = __dsub_rn(scp->wpy, y);
    const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const double dtheta = atan2(dy, dx);
    float angf = (float)__dsub_rn(dtheta, th);
    angf = normalize_angle_f(angf, (float)(-M_PI), (float)M_PI);
    const double ang_diff = __ddiv_rn(fabs((double)angf), M_PI);
    const double vel_diff = __ddiv_rn(fabs(__dsub_rn(B.max_vel_x, vx)), B.max_vel_x);
    const double cm = __ddiv_rn(costmap_sum, (double)S);
    double c = __dmul_rn(B.w_vel, vel_diff);
    c = __dadd_rn(c, __dmul_rn(B.w_dist, d));
    c = __dadd_rn(c, __dmul_rn(B.w_ang, ang_diff));
    c = __dadd_rn(c, __dmul_rn(B.w_map, cm));
    c = __dadd_rn(c, __dmul_rn(B.w_soc, social_work));
    cost = (float)c;
  }
  if (in_range) {
    const size_t o = (size_t)scene * B.n_v * n_w + idx;
    B.costs[o] = cost;
    if (B.costs_host)
      B.costs_host[o] = cost;
    B.npts[o] = (uint16_t)npts;
  }

  // ---- arg-min: warp shuffle, then block, then last block of the scene -----------------------
  float bc = (in_range && eligible(cost, v_s)) ? cost : -1.f;
  uint32_t bi = idx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oc = __shfl_xor_sync(0xffffffffu, bc, o);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(oc, oi, bc, bi, B.linvels, B.angvels, n_w)) {
      bc = oc;
      bi = oi;
    }
  }
  const uint32_t warp = tid >> 5, lane = tid & 31u, nwarps = (T + 31u) >> 5;
  if (lane == 0) {
    s_redc[warp] = bc;
    s_redi[warp] = bi;
  }
  __syncthreads();
  if (warp == 0) {
    bc = (lane < nwarps) ? s_redc[lane] : -1.f;
    bi = (lane < nwarps) ? s_redi[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float oc = __shfl_xor_sync(0xffffffffu, bc, o);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(oc, oi, bc, bi, B.linvels, B.angvels, n_w)) {
        bc = oc;
        bi = oi;
      }
    }
    if (lane == 0) {
      SfwBlockBest bb;
      bb.cost = bc;
      bb.index = bi;
      B.blockbest[(size_t)scene * B.tiles_per_scene + tile] = bb;
      __threadfence();
      const unsigned int done = atomicAdd(&B.counters[scene], 1u);
      s_last = (done == B.tiles_per_scene - 1u);
    }
    __syncwarp();
    if (s_last) {
      __threadfence();
      bc = -1.f;
      bi = 0u;
      const SfwBlockBest *bbp = B.blockbest + (size_t)scene * B.tiles_per_scene;
      for (uint32_t t = lane; t < B.tiles_per_scene; t += 32u) {
        const float oc = __ldcg(&bbp[t].cost);
        const uint32_t oi = __ldcg(&bbp[t].index);
        if (better(oc, oi, bc, bi, B.linvels, B.angvels, n_w)) {
          bc = oc;
          bi = oi;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float oc = __shfl_xor_sync(0xffffffffu, bc, o);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(oc, oi, bc, bi, B.linvels, B.angvels, n_w)) {
          bc = oc;
          bi = oi;
        }
      }
      if (lane == 0) {
        SfwBest r;
        r.valid = (bc >= 0.f) ? 1 : 0;
        r.index = r.valid ? bi : 0u;
        r.cost = r.valid ? bc : 0.f;
        r.reserved0 = 0.f;
        r.v = r.valid ? B.linvels[bi / n_w] : 0.0;
        r.w = r.valid ? B.angvels[bi % n_w] : 0.0;
        B.best[scene] = r;
        if (B.best_host)
          B.best_host[scene] = r;
        B.counters[scene] = 0u; // ready for the next launch
        if (B.xchg.enabled)
          export_best(B.xchg, scene, r);
      }
    }
  }
}

// ================================================================================================
// Trajectory points of one sample (Trajectory::addPoint, src/trajectory.cpp:36-40): pure robot
// kinematics replayed for the number of points the scorer recorded.
// ================================================================================================
extern "C" __global__ void sfw_points_kernel(const __grid_constant__ SfwBatchDev B, uint32_t scene,
                                             uint32_t idx, uint32_t n_points, double *out_xyz) {
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  const SfwSceneDev *scp = B.scenes + scene;
  const double v_s = B.linvels[idx / B.n_w], w_s = B.angvels[idx % B.n_w];
  double x = scp->rx, y = scp->ry, th = scp->rth, vx = scp->rvx, vth = scp->rvth;
  const double vy = scp->rvy, dt = B.dt;
  const double ax_dt = __dmul_rn(B.acc_x, dt), ath_dt = __dmul_rn(B.acc_th, dt);
  for (uint32_t i = 0; i < n_points; ++i) {
    out_xyz[3 * i] = x;
    out_xyz[3 * i + 1] = y;
    out_xyz[3 * i + 2] = th;
    double sn, cs;
    sincos(th, &sn, &cs);
    vx = step_velocity(v_s, vx, ax_dt);
    vth = step_velocity(w_s, vth, ath_dt);
    double lx = __dmul_rn(vx, cs), ly = __dmul_rn(vx, sn);
    if (vy != 0.0) {
      double sn2, cs2;
      sincos(__dadd_rn(1.57079632679489661923, th), &sn2, &cs2);
      lx = __dadd_rn(lx, __dmul_rn(vy, cs2));
      ly = __dadd_rn(ly, __dmul_rn(vy, sn2));
    }
    x = __dadd_rn(x, __dmul_rn(lx, dt));
    y = __dadd_rn(y, __dmul_rn(ly, dt));
    th = __dadd_rn(th, __dmul_rn(vth, dt));
  }
}

// The same replay for many samples at once (RViz markers, reference src/sfw_planner.cpp:366-374): thread t
// handles sample first + t * stride and writes its recorded points to out_xyz[t][max_points][3].
extern "C" __global__ void sfw_marker_points_kernel(const __grid_constant__ SfwBatchDev B, uint32_t scene,
                                                    uint32_t first, uint32_t stride, uint32_t count,
                                                    uint32_t max_points, double *out_xyz, uint16_t *out_n) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count)
    return;
  const uint32_t idx = first + t * stride;
  const SfwSceneDev *scp = B.scenes + scene;
  const uint32_t n_rec = B.npts[(size_t)scene * B.n_v * B.n_w + idx];
  out_n[t] = (uint16_t)n_rec;
  const uint32_t n_points = min(n_rec, max_points);
  const double v_s = B.linvels[idx / B.n_w], w_s = B.angvels[idx % B.n_w];
  double x = scp->rx, y = scp->ry, th = scp->rth, vx = scp->rvx, vth = scp->rvth;
  const double vy = scp->rvy, dt = B.dt;
  const double ax_dt = __dmul_rn(B.acc_x, dt), ath_dt = __dmul_rn(B.acc_th, dt);
  double *o = out_xyz + (size_t)t * max_points * 3u;
  for (uint32_t i = 0; i < n_points; ++i) {
    o[3 * i] = x;
    o[3 * i + 1] = y;
    o[3 * i + 2] = th;
    double sn, cs;
    sincos(th, &sn, &cs);
    vx = step_velocity(v_s, vx, ax_dt);
    vth = step_velocity(w_s, vth, ath_dt);
    double lx = __dmul_rn(vx, cs), ly = __dmul_rn(vx, sn);
    if (vy != 0.0) {
      double sn2, cs2;
      sincos(__dadd_rn(1.57079632679489661923, th), &sn2, &cs2);
      lx = __dadd_rn(lx, __dmul_rn(vy, cs2));
      ly = __dadd_rn(ly, __dmul_rn(vy, sn2));
    }
    x = __dadd_rn(x, __dmul_rn(lx, dt));
    y = __dadd_rn(y, __dmul_rn(ly, dt));
    th = __dadd_rn(th, __dmul_rn(vth, dt));
  }
}

// SFWPlanner::mayIStop (reference src/sfw_planner.cpp:718-765; its only call, :637, is commented out upstream):
// brake at the acceleration limits from (vl_x, vl_y, va) and check the footprint at every pose until the robot
// stands.  Faithful to the reference including its argument mix-up: the position update is handed the ANGULAR
// VELOCITY where the helpers expect the heading (:731-732).  out[0] = 1 can stop / 0 collision, out[1] = steps.
extern "C" __global__ void sfw_may_i_stop_kernel(const __grid_constant__ SfwBatchDev B, uint32_t scene, double vl_x,
                                                 double vl_y, double va, double x, double y, double th, double dt,
                                                 int *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  const SfwSceneDev *scp = B.scenes + scene;
  MapView mv;
  mv.win = nullptr;
  mv.glob = B.maps + scp->map_off;
  mv.ox = scp->origin_x;
  mv.oy = scp->origin_y;
  mv.res = scp->resolution;
  mv.rinv = __ddiv_rn(1.0, scp->resolution);
  mv.sx = scp->size_x;
  mv.sy = scp->size_y;
  mv.pitch = B.map_pitch;
  mv.wx0 = 0;
  mv.wy0 = 0;
  mv.wwp = 0;
  mv.wh = 0;
  mv.free_bits = nullptr;
  mv.fw32 = 0;
  const double2 *fp = B.footprint + scp->fp_off;
  const double ax_dt = __dmul_rn(B.acc_x, dt), ath_dt = __dmul_rn(B.acc_th, dt);
  double lvx = vl_x, lvy = vl_y, av = va, xp = x, yp = y, hp = th;
  int steps = 0, ok = 1;
  while ((lvx > 0.0 || lvy > 0.0) && steps < 1000000) {
    lvx = step_velocity(0.0, lvx, ax_dt);
    lvy = step_velocity(0.0, lvy, ax_dt);
    av = step_velocity(0.0, av, ath_dt);
    double sa, ca, sb, cb; // the helpers get `av` as their theta argument (:731-732)
    sincos(av, &sa, &ca);
    sincos(__dadd_rn(1.57079632679489661923, av), &sb, &cb);
    xp = __dadd_rn(xp, __dmul_rn(__dadd_rn(__dmul_rn(lvx, ca), __dmul_rn(lvy, cb)), dt));
    yp = __dadd_rn(yp, __dmul_rn(__dadd_rn(__dmul_rn(lvx, sa), __dmul_rn(lvy, sb)), dt));
    hp = __dadd_rn(hp, __dmul_rn(av, dt));
    double sn, cs;
    sincos(hp, &sn, &cs);
    const int fc = footprint_cost(mv, fp, (int)scp->n_fp, xp, yp, sn, cs);
    ++steps;
    if (fc < 0) { // footprint_cost folds the reference's "< 0 || >= 254" into -1
      ok = 0;
      break;
    }
  }
  out[0] = ok;
  out[1] = steps;
}

// ================================================================================================
// launch wrappers (called from sfw_abi.cu)
// ================================================================================================
size_t sfw_small_smem_bytes(uint32_t win_wp, uint32_t win_h, uint32_t P, uint32_t M, uint32_t F, uint32_t T) {
  const uint32_t win_bytes = win_wp * win_h;
  const size_t P2 = (P + 1u) / 2u, Mp = sfw_obst_slots(M);
  size_t off = (win_bytes + 127u) & ~127u;
  off += 5u * P2 * 16u;
  off += Mp * 8u;
  off += (size_t)F * 16u;
  off += 16u;
  off += 64u * 4u;
  off += 2u * ((win_bytes ? (size_t)win_h * ((win_wp + 31u) / 32u) * 4u : 0u) + 15u & ~(size_t)15u);
  off += 3u * P2 * T * 16u;
  return off;
}

namespace {
typedef void (*SmallKernel)(const SfwBatchDev, const CUtensorMap);
struct SmallVariant {
  uint32_t maxt;
  SmallKernel fn;
  const char *name;
};
const SmallVariant kSmall[] = {
    {384, sfw_score_small<384, false>, "sfw_score_small<384>"},
    {448, sfw_score_small<448, false>, "sfw_score_small<448>"},
    {512, sfw_score_small<512, false>, "sfw_score_small<512>"},
};
const SmallVariant kSmallShare[] = {
    {384, sfw_score_small<384, true>, "sfw_score_small<384,share>"},
    {448, sfw_score_small<448, true>, "sfw_score_small<448,share>"},
    {512, sfw_score_small<512, true>, "sfw_score_small<512,share>"},
};
const SmallVariant kWarpPath = {SFW_PATH_WARP_THREADS, sfw_score_small<SFW_PATH_WARP_THREADS, 2>, "sfw_score_small<warp-per-path>"};
const SmallVariant &small_variant(uint32_t T, bool share = false) {
  const SmallVariant *tab = share ? kSmallShare : kSmall;
  for (int i = 0; i < 3; ++i)
    if (T <= tab[i].maxt)
      return tab[i];
  return tab[2];
}
} // namespace

const char *sfw_small_kernel_name(uint32_t T, bool share) { return small_variant(T, share).name; }

// Largest dynamic shared memory a block of sfw_score_small may request on the current device
// (opt-in limit minus the kernel's static shared memory); also opts every variant in.
cudaError_t sfw_small_max_dynamic_smem(size_t *bytes) {
  // function attributes are per device: cache (and opt in) per device, not per process
  static size_t cached_dev[64] = {};
  int dev = 0;
  {
    const cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess)
      return e0;
  }
  size_t scratch = 0;
  size_t &cached = (dev >= 0 && dev < 64) ? cached_dev[dev] : scratch;
  if (!cached) {
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess)
      return e;
    size_t dyn = (size_t)optin;
    for (const SmallVariant *tab : {kSmall, kSmallShare})
      for (int i = 0; i < 3; ++i) {
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, tab[i].fn);
        if (e != cudaSuccess)
          return e;
        dyn = std::min(dyn, (size_t)optin - fa.sharedSizeBytes);
      }
    {
      cudaFuncAttributes fa;
      e = cudaFuncGetAttributes(&fa, kWarpPath.fn);
      if (e != cudaSuccess)
        return e;
      dyn = std::min(dyn, (size_t)optin - fa.sharedSizeBytes);
    }
    for (const SmallVariant *tab : {kSmall, kSmallShare})
      for (int i = 0; i < 3; ++i) {
        e = cudaFuncSetAttribute(tab[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess)
          return e;
      }
    e = cudaFuncSetAttribute(kWarpPath.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess)
      return e;
    cached = dyn;
  }
  *bytes = cached;
  return cudaSuccess;
}

namespace {
// dependent = true: programmatic dependent launch behind the previous kernel of the stream (the kernel itself
// calls griddep_wait() before it reads what that kernel wrote)
cudaError_t launch_small_kernel(SmallKernel fn, uint32_t grid, uint32_t T, size_t smem_bytes, cudaStream_t stream,
                                const SfwBatchDev &B, const CUtensorMap &tmap, bool dependent) {
  if (!dependent) {
    fn<<<grid, T, smem_bytes, stream>>>(B, tmap);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, fn, B, tmap);
}
} // namespace

cudaError_t sfw_launch_small(const SfwBatchDev &B, const CUtensorMap &tmap, uint32_t T,
                             size_t smem_bytes, cudaStream_t stream, bool dependent) {
  const uint32_t grid = (B.launch_scenes ? B.launch_scenes : B.n_scenes) * B.tiles_per_scene;
  return launch_small_kernel(small_variant(T, B.share.mode != 0).fn, grid, T, smem_bytes, stream, B, tmap, dependent);
}

// Path writer with one warp per path (B.share.mode 1 or 2; B.tiles_per_scene = blocks of SFW_PATH_WARP_THREADS / 32
// paths per scene).
cudaError_t sfw_launch_warp_paths(const SfwBatchDev &B, const CUtensorMap &tmap, size_t smem_bytes,
                                  cudaStream_t stream, bool dependent) {
  const uint32_t grid = (B.launch_scenes ? B.launch_scenes : B.n_scenes) * B.tiles_per_scene;
  return launch_small_kernel(kWarpPath.fn, grid, SFW_PATH_WARP_THREADS, smem_bytes, stream, B, tmap, dependent);
}

cudaError_t sfw_warp_paths_occupancy(size_t smem_bytes, int *blocks_per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kWarpPath.fn, SFW_PATH_WARP_THREADS, smem_bytes);
}

cudaError_t sfw_small_occupancy(uint32_t T, size_t smem_bytes, int *blocks_per_sm) {
  int a = 0, b = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, small_variant(T, false).fn, (int)T, smem_bytes);
  if (e != cudaSuccess)
    return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, small_variant(T, true).fn, (int)T, smem_bytes);
  *blocks_per_sm = std::min(a, b);
  return e;
}

cudaError_t sfw_launch_points(const SfwBatchDev &B, uint32_t scene, uint32_t idx, uint32_t n_points,
                              double *out_xyz, cudaStream_t stream) {
  sfw_points_kernel<<<1, 32, 0, stream>>>(B, scene, idx, n_points, out_xyz);
  return cudaGetLastError();
}


cudaError_t sfw_launch_marker_points(const SfwBatchDev &B, uint32_t scene, uint32_t first, uint32_t stride,
                                     uint32_t count, uint32_t max_points, double *out_xyz, uint16_t *out_n,
                                     cudaStream_t stream) {
  sfw_marker_points_kernel<<<(count + 127u) / 128u, 128, 0, stream>>>(B, scene, first, stride, count, max_points,
                                                                    out_xyz, out_n);
  return cudaGetLastError();
}

cudaError_t sfw_launch_may_i_stop(const SfwBatchDev &B, uint32_t scene, double vl_x, double vl_y, double va, double x,
                                  double y, double th, double dt, int *out, cudaStream_t stream) {
  sfw_may_i_stop_kernel<<<1, 32, 0, stream>>>(B, scene, vl_x, vl_y, va, x, y, th, dt, out);
  return cudaGetLastError();
}
