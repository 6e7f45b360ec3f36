// sfw_crowd.cu — sm_100a kernel for dense crowds (more pedestrians than the thread-per-trajectory
// kernel can keep in one shared-memory column): ONE BLOCK PER TRAJECTORY.
//
//   * persistent blocks pull (scene, sample) work items from a global counter;
//   * prologue: the robot rollout (FP64, identical arithmetic to sfw_kernels.cu) is computed for all
//     S steps first — it does not depend on the pedestrians — and the footprint of every (step, edge)
//     is rasterised in parallel over the block; a trajectory the costmap rejects at step s only
//     simulates the crowd for the s steps whose points the reference would still have recorded;
//   * crowd step: pedestrians live in shared memory as PAIRS (packed FP32x2 layout).  Thread t owns
//     pair t and evaluates it against the next floor((P2-1)/2) pairs in cyclic order, so every
//     unordered pair of pairs is evaluated exactly once (lightsfm's pair force is antisymmetric) and
//     every thread does the same amount of work.  The reaction force is pushed into a per-WARP
//     accumulator row (lanes of a warp hit consecutive entries, rows are private to the warp, so no
//     atomics), rows are summed after a block barrier and each thread integrates its own pair;
//   * robot social force / social work / collision flag: warp shuffles + one shared-memory pass.
//
// Semantics followed: reference src/sfw_planner.cpp:475-705 (see sfw_forces.cuh for the pieces).
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "sfw_dev.h"
#include "sfw_kernels.h"
#include "sfw_forces.cuh"

namespace {

// Two instantiations: 256 threads (8 warps, two blocks per SM) for crowds, and 128 threads (4 warps, four blocks per
// SM) for small crowds on grids that would otherwise need a second wave of 256-thread blocks (BASELINE's 21 x 21
// tick: 441 trajectories are one wave of 592 slots instead of two of 296).  The host picks one per plan.
// The thread that carries a trajectory's social work: lane 0 of warp 1.  Warp 0 integrates the pairs of a small
// crowd in phase 2; the robot's terms are reduced beside it, not after it (one warp alone runs a dependent chain at
// 7+ cycles per instruction: a 5 x 9 tick is nothing but such chains).
constexpr uint32_t kWorkWarp = 1u, kWorkTid = 32u * kWorkWarp;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The robot's obstacle sum (never cut off) by one whole warp: lane l takes the point pairs l, l + 32, ... of the
// clustered list (4 pairs per cluster behind its 2-slot header), then a butterfly sum.  All 32 lanes must call.
__device__ __forceinline__ void robot_obstacle_sum_warp(const float2 *__restrict__ obs, uint32_t M, float c_obs,
                                                        float px, float py, uint32_t lane, float &sx, float &sy) {
  f2 ax = bc2(0.f), ay = bc2(0.f);
  const f2 eps = bc2(1e-30f);
  const f2 qx = bc2(px * c_obs), qy = bc2(py * c_obs);
  const uint32_t n4 = (M / SFW_OBST_CLUSTER_SLOTS) * (SFW_OBST_CLUSTER / 2);
  for (uint32_t i = lane; i < n4; i += 32u) {
    const uint32_t g = i / (SFW_OBST_CLUSTER / 2), r = i - g * (SFW_OBST_CLUSTER / 2);
    const float4 p = *reinterpret_cast<const float4 *>(obs + g * SFW_OBST_CLUSTER_SLOTS + 2u + 2u * r);
    const f2 dx = sub2(qx, mk2(p.x, p.z)), dy = sub2(qy, mk2(p.y, p.w));
    const f2 d2 = fma2(dx, dx, fma2(dy, dy, eps));
    const f2 rd = rsqrt2(d2);
    float d0, d1;
    un2(mul2(d2, rd), d0, d1);
    const f2 e = mul2(mk2(ex2_approx(-d0), ex2_approx(-d1)), rd);
    ax = fma2(e, dx, ax);
    ay = fma2(e, dy, ay);
  }
  float x0, x1, y0, y1;
  un2(ax, x0, x1);
  un2(ay, y0, y1);
  sx = warp_sum(x0 + x1);
  sy = warp_sum(y0 + y1);
}

struct CrowdSmem {
  float4 *pos, *vel, *goal, *par, *par2; // [P2]
  float4 *frc;                           // [warps][P2]
  float2 *obs;                           // [M]
  double2 *fp;                           // [F]
  double *rx, *ry, *rth, *rsn, *rcs;     // [S + 1] pose before step s (entry S = final pose)
  float *rvx;                            // [S + 1] robot vx after s updates (float view for the SFM)
  int *fcm;                              // [S] footprint max per step (>= 254: lethal, 1000: off map)
  uint8_t *goalflag;                     // [2 * P2]
  float *red;                            // [warps][4]
  int *flags;                            // [0] work item, [1] hit, [2] first bad step
};

__device__ __forceinline__ CrowdSmem carve(unsigned char *base, uint32_t P2, uint32_t M, uint32_t F, uint32_t S,
                                           uint32_t warps) {
  CrowdSmem s;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char *p = base + off;
    off += (bytes + 15u) & ~(size_t)15u;
    return p;
  };
  s.pos = (float4 *)take(16u * P2);
  s.vel = (float4 *)take(16u * P2);
  s.goal = (float4 *)take(16u * P2);
  s.par = (float4 *)take(16u * P2);
  s.par2 = (float4 *)take(16u * P2);
  s.frc = (float4 *)take(16u * P2 * warps);
  s.obs = (float2 *)take(8u * M);
  s.fp = (double2 *)take(16u * F);
  s.rx = (double *)take(8u * (S + 1));
  s.ry = (double *)take(8u * (S + 1));
  s.rth = (double *)take(8u * (S + 1));
  s.rsn = (double *)take(8u * (S + 1));
  s.rcs = (double *)take(8u * (S + 1));
  s.rvx = (float *)take(4u * (S + 1));
  s.fcm = (int *)take(4u * S);
  s.goalflag = (uint8_t *)take(2u * P2);
  s.red = (float *)take(4u * 4u * warps);
  s.flags = (int *)take(16);
  return s;
}

} // namespace

size_t sfw_crowd_smem_bytes(uint32_t P, uint32_t M, uint32_t F, uint32_t S, uint32_t threads) {
  const size_t P2 = (P + 1u) / 2u, Mp = sfw_obst_slots(M), warps = threads / 32u;
  auto r = [](size_t b) { return (b + 15u) & ~(size_t)15u; };
  return 5 * r(16 * P2) + r(16 * P2 * warps) + r(8 * Mp) + r(16 * (size_t)F) + 5 * r(8 * ((size_t)S + 1)) +
         r(4 * ((size_t)S + 1)) + r(4 * (size_t)S) + r(2 * P2) + r(16 * warps) + 16;
}

// ================================================================================================
// Arg-min of one scene's cost vector with the reference's tie-breaks (sfw_planner.cpp:394-414), by one block
// of up to 512 threads.  Called from sfw_argmin_kernel (one block per scene) or, for small batches, from the last block
// of sfw_score_crowd to finish (the costs then come from other SMs: L2 loads).
// ================================================================================================
__device__ __forceinline__ void scene_argmin(const SfwBatchDev &B, uint32_t scene, float *s_c, uint32_t *s_i) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t n_w = B.n_w, n = B.n_v * n_w;
  const float *costs = B.costs + (size_t)scene * n;
  float bc = -1.f;
  uint32_t bi = 0u;
  for (uint32_t i = B.row_begin * n_w + tid; i < B.row_end * n_w; i += blockDim.x) {
    const float c = __ldcg(costs + i);
    if (eligible(c, B.linvels[i / n_w]) && better(c, i, bc, bi, B.linvels, B.angvels, n_w)) {
      bc = c;
      bi = i;
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
    s_c[warp] = bc;
    s_i[warp] = bi;
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t n_warps = blockDim.x >> 5; // at most 16 (s_c / s_i hold 16 entries)
    bc = (lane < n_warps) ? s_c[lane] : -1.f;
    bi = (lane < n_warps) ? s_i[lane] : 0u;
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
      if (B.xchg.enabled)
        export_best(B.xchg, scene, r);
    }
  }
  __syncthreads(); // s_c / s_i may be reused for the next scene
}

__global__ void __launch_bounds__(256) sfw_argmin_kernel(const __grid_constant__ SfwBatchDev B) {
  __shared__ float s_c[16];
  __shared__ uint32_t s_i[16];
  scene_argmin(B, blockIdx.x, s_c, s_i);
}

// ================================================================================================
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
sfw_score_crowd(const __grid_constant__ SfwBatchDev B, unsigned int *__restrict__ work_counter, int fused_argmin) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kCrowdThreads = THREADS, kCrowdWarps = THREADS / 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t n_w = B.n_w;
  // Rollout prefix sharing (SfwShareDev): mode 1 / 2 items are shared paths that write a record per step,
  // mode 3 items are the samples, each starting from the record of its fork point.  A record here is
  // {double social_work; int alive; int steps_done} + pos[P2] + vel[P2] (float4) + goal flags (2 P2 bytes):
  // the robot rollout and the footprint checks are recomputed per item in the prologue anyway.
  const uint32_t share_mode = B.share.mode;
  const bool writer = share_mode == 1u || share_mode == 2u;
  const uint32_t per_scene = share_mode == 1u   ? 4u
                             : share_mode == 2u ? 2u * n_w + 2u * B.n_v
                                                : (B.row_end - B.row_begin) * n_w;
  const uint32_t total = B.n_scenes * per_scene;
  const int S = B.num_steps;
  const SfmConst K = make_sfm_const(B);
  const double dt = B.dt;
  const double ax_dt = __dmul_rn(B.acc_x, dt), ath_dt = __dmul_rn(B.acc_th, dt);
  const float dtf = B.dtf;
  __shared__ int s_item;
  uint32_t staged_scene = 0xffffffffu;

  for (;;) {
    __syncthreads(); // previous item fully retired before shared state is reused
    if (tid == 0)
      s_item = (int)atomicAdd(work_counter, 1u);
    __syncthreads();
    const uint32_t item = (uint32_t)s_item;
    if (item >= total)
      break;
    const uint32_t scene = item / per_scene;
    const uint32_t local = item - scene * per_scene;
    const uint32_t idx = B.row_begin * n_w + local;
    const SfwSceneDev *__restrict__ scp = B.scenes + scene;
    const uint32_t P2 = scp->n_pairs, M = scp->n_obst, F = scp->n_fp;
    const uint32_t n_groups = scp->n_groups;
    const CrowdSmem sm = carve(smem_raw, P2, M, F, (uint32_t)S, (uint32_t)kCrowdWarps);
    double v_s, w_s;
    int s0 = 0;                     // first step this item simulates itself
    const uint8_t *rec_in = nullptr; // record it starts from
    uint8_t *rec_out = nullptr;     // writer items: this path's records [step count]
    if (share_mode) {
      const SfwShareDev &H = B.share;
      const uint16_t *kv = H.kv + (size_t)scene * B.n_v, *kw = H.kw + (size_t)scene * n_w;
      const uint8_t *dirv = H.dirv + (size_t)scene * B.n_v, *dirw = H.dirw + (size_t)scene * n_w;
      uint8_t *base = H.records + (size_t)scene * H.scene_stride;
      const size_t R = H.rec_bytes, K1 = (size_t)H.kmax + 1u;
      size_t parent = 0;
      if (share_mode == 3u) {
        const uint32_t r = idx / n_w, c = idx - r * n_w;
        v_s = B.linvels[r];
        w_s = B.angvels[c];
        const int kvr = kv[r], kwc = kw[c];
        s0 = max(kvr, kwc);
        parent = kvr == kwc  ? (size_t)dirv[r] * 2u + dirw[c]
                 : kvr > kwc ? 4u + (size_t)dirv[r] * n_w + c
                             : 4u + 2u * (size_t)n_w + (size_t)r * 2u + dirw[c];
      } else if (share_mode == 1u) {
        v_s = (local >> 1) ? 1.0e300 : -1.0e300;
        w_s = (local & 1u) ? 1.0e300 : -1.0e300;
        rec_out = base + (size_t)local * K1 * R;
      } else if (local < 2u * n_w) {
        const uint32_t dv = local / n_w, c = local - dv * n_w;
        v_s = dv ? 1.0e300 : -1.0e300;
        w_s = B.angvels[c];
        s0 = kw[c];
        parent = (size_t)dv * 2u + dirw[c];
        rec_out = base + (4u + (size_t)local) * K1 * R;
      } else {
        const uint32_t q = local - 2u * n_w, r = q >> 1, dw = q & 1u;
        v_s = B.linvels[r];
        w_s = dw ? 1.0e300 : -1.0e300;
        s0 = kv[r];
        parent = (size_t)dirv[r] * 2u + dw;
        rec_out = base + (4u + (size_t)local) * K1 * R;
      }
      if (s0 > 0)
        rec_in = base + (parent * K1 + (size_t)s0) * R;
    } else {
      v_s = B.linvels[idx / n_w];
      w_s = B.angvels[idx % n_w];
    }
    const int n_run = writer ? (int)B.share.kmax : S; // steps this item is asked to reach
    const size_t out = (size_t)scene * B.n_v * n_w + idx;
    if (!writer && !B.score_zero && v_s == 0.0 && w_s == 0.0) { // sfw_planner.cpp:349-352
      if (tid == 0) {
        B.costs[out] = SFW_COST_SKIPPED;
        if (B.costs_host)
          B.costs_host[out] = SFW_COST_SKIPPED;
        B.npts[out] = 0;
      }
      continue;
    }

    // ---- stage the scene (static part once per scene, state every item) ----------------------
    const float4 *rec_pv = rec_in ? reinterpret_cast<const float4 *>(rec_in + 16) : nullptr;
    const uint8_t *rec_gf = rec_in ? rec_in + 16 + 32u * (size_t)P2 : nullptr;
    for (uint32_t k = tid; k < P2; k += kCrowdThreads) {
      sm.pos[k] = rec_pv ? rec_pv[k] : B.pedPos[scp->ped_off + k];
      sm.vel[k] = rec_pv ? rec_pv[P2 + k] : B.pedVel[scp->ped_off + k];
      sm.goalflag[2 * k] = rec_gf ? rec_gf[2 * k] : B.goal_bits[2u * (scp->ped_off + k)];
      sm.goalflag[2 * k + 1] = rec_gf ? rec_gf[2 * k + 1] : B.goal_bits[2u * (scp->ped_off + k) + 1u];
      for (uint32_t w = 0; w < (uint32_t)kCrowdWarps; ++w)
        sm.frc[w * P2 + k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (staged_scene != scene) {
      for (uint32_t k = tid; k < P2; k += kCrowdThreads) {
        sm.goal[k] = B.pedGoal[scp->ped_off + k];
        sm.par[k] = B.pedPar[scp->ped_off + k];
        sm.par2[k] = B.pedPar2[scp->ped_off + k];
      }
      for (uint32_t o = tid; o < M; o += kCrowdThreads)
        sm.obs[o] = B.obst[scp->obs_off + o];
      for (uint32_t f = tid; f < F; f += kCrowdThreads)
        sm.fp[f] = B.footprint[scp->fp_off + f];
      staged_scene = scene;
    }

    // ---- prologue A: robot rollout, sfw_planner.cpp:581-588 (identical FP64 ops to the small kernel) ----
    const double vy = scp->rvy;
    if (tid == 0) { // heading and speed first: no trigonometry on this chain
      double th = scp->rth, vx = scp->rvx, vth = scp->rvth;
      sm.rth[0] = th;
      sm.rvx[0] = (float)vx;
      for (int i = 0; i < S; ++i) {
        vx = step_velocity(v_s, vx, ax_dt);
        vth = step_velocity(w_s, vth, ath_dt);
        th = __dadd_rn(th, __dmul_rn(vth, dt));
        sm.rth[i + 1] = th;
        sm.rvx[i + 1] = (float)vx;
        sm.rx[i + 1] = vx; // the new speed is consumed (as a double) when positions are chained below
      }
      sm.flags[1] = 0;
      sm.flags[2] = S;
    }
    __syncthreads();
    for (int i = (int)tid; i <= S; i += kCrowdThreads) {
      double sn, cs;
      sincos(sm.rth[i], &sn, &cs);
      sm.rsn[i] = sn;
      sm.rcs[i] = cs;
      if (i < S)
        sm.fcm[i] = 0;
    }
    __syncthreads();
    if (tid == 0) {
      double x = scp->rx, y = scp->ry;
      sm.rx[0] = x;
      sm.ry[0] = y;
      for (int i = 0; i < S; ++i) {
        const double vx = sm.rx[i + 1];
        const double sn = sm.rsn[i], cs = sm.rcs[i];
        double lx = __dmul_rn(vx, cs), ly = __dmul_rn(vx, sn);
        if (vy != 0.0) {
          double sn2, cs2;
          sincos(__dadd_rn(1.57079632679489661923, sm.rth[i]), &sn2, &cs2);
          lx = __dadd_rn(lx, __dmul_rn(vy, cs2));
          ly = __dadd_rn(ly, __dmul_rn(vy, sn2));
        }
        x = __dadd_rn(x, __dmul_rn(lx, dt));
        y = __dadd_rn(y, __dmul_rn(ly, dt));
        sm.rx[i + 1] = x;
        sm.ry[i + 1] = y;
      }
    }
    __syncthreads();

    // ---- prologue B: footprint of every (step, edge) in parallel (sfw_planner.cpp:545-575) -------
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
    {
      const uint32_t per_step = (F < 3u) ? 1u : F + 1u; // F edges + the centre-in-map test
      for (uint32_t w = tid; w < (uint32_t)S * per_step; w += kCrowdThreads) {
        const uint32_t s = w / per_step, e = w - s * per_step;
        const double x = sm.rx[s], y = sm.ry[s], sn = sm.rsn[s], cs = sm.rcs[s];
        int res = 0;
        if (F < 3u || e == F) {
          int cx, cy;
          if (!world_to_map(mv, x, y, cx, cy))
            res = 1000;
          else if (F < 3u) {
            const int c = (int)cell_cost(mv, cx, cy);
            res = (c >= 253) ? 1000 : c;
          }
        } else {
          const double2 v0 = sm.fp[e], v1 = sm.fp[(e + 1u == F) ? 0u : e + 1u];
          int x0, y0, x1, y1;
          const bool ok0 = world_to_map(mv, __dadd_rn(x, __dsub_rn(__dmul_rn(v0.x, cs), __dmul_rn(v0.y, sn))),
                                        __dadd_rn(y, __dadd_rn(__dmul_rn(v0.x, sn), __dmul_rn(v0.y, cs))), x0, y0);
          const bool ok1 = world_to_map(mv, __dadd_rn(x, __dsub_rn(__dmul_rn(v1.x, cs), __dmul_rn(v1.y, sn))),
                                        __dadd_rn(y, __dadd_rn(__dmul_rn(v1.x, sn), __dmul_rn(v1.y, cs))), x1, y1);
          res = (ok0 && ok1) ? line_max<false>(mv, x0, y0, x1, y1) : 1000;
        }
        if (res > 0)
          atomicMax(&sm.fcm[s], res);
      }
    }
    __syncthreads();
    for (int i = (int)tid; i < S; i += kCrowdThreads)
      if (sm.fcm[i] >= 254)
        atomicMin(&sm.flags[2], i);
    __syncthreads();
    const int S_eff = sm.flags[2]; // steps whose pose is legal; < S => the trajectory is invalid

    // ---- crowd simulation ---------------------------------------------------------------------
    const double base_x = scp->rx, base_y = scp->ry;
    const float rr2 = B.rr2;
    const float a_obs_scale = scp->a_obs_scale;
    double social_work = 0.0; // thread kWorkTid only
    const uint32_t half = P2 ? (P2 - 1u) / 2u : 0u; // full cyclic offsets
    const bool even = P2 >= 2u && (P2 & 1u) == 0u; // + the opposite pair, first half of the ring only
    const uint32_t owned = (P2 + kCrowdThreads - 1u) / kCrowdThreads;
    float4 *myrow = sm.frc + warp * P2;
    // Obstacle sums of the pedestrians: with at most 128 pairs the warps beyond the first `pair_warps` own no pair
    // and would idle through the force phase.  They take the obstacle clusters instead — helper h of pair a sums
    // the clusters h, h + n_help, ... into its own warp's accumulator row — so the sums run beside the pair forces
    // instead of after them (a tick at the reference's 5 x 9 samples is one wave of this kernel: latency is all
    // there is).  The split only depends on the crowd size, never on the batch or on prefix sharing.
    const uint32_t pair_warps = (P2 + 31u) >> 5;
    // Up to 32 pairs (64 pedestrians: every crowd AUTO sends here for a small grid) one warp would own every pair
    // and walk its cyclic offsets one after the other.  Instead the force phase is SPREAD over the block: lane a of
    // every warp stands for pair a; warp 0 takes the robot-pair and the inside-pair forces, warp w = 1 .. 6 the cyclic
    // offsets w, w + 6, ... (all pairs at once) and the obstacle clusters 6 - w, 12 - w, ... (so the warps with the
    // fewest offsets get the clusters); the last warp is left to the robot's obstacle sum.  Every warp adds into its
    // own accumulator row, the rows are summed in warp order after the barrier.  A step then costs about two
    // pair-force latencies instead of 2 + (P2 - 1) / 2 of them.  (A separate body on purpose: folding the two
    // layouts into one loop with run-time strides cost the dense-crowd path 5 %.)
    constexpr uint32_t kSpreadWarps = (uint32_t)kCrowdWarps - 1u; // warps 0 .. 6 share the pedestrians' forces
    const bool spread = P2 >= 4u && P2 <= 32u; // up to 3 pairs the owner layout with obstacle helpers measured as fast or faster
    const uint32_t n_help = (!spread && P2 && 2u * pair_warps <= (uint32_t)kCrowdWarps) ? (uint32_t)kCrowdWarps / pair_warps - 1u : 0u;
    const uint32_t help_idx = n_help ? tid / (32u * pair_warps) : 0u; // 0: the threads that own the pairs
    const uint32_t help_pair = n_help ? tid - help_idx * 32u * pair_warps : 0u;
    int steps_done = 0;
    bool collided = false;
    if (rec_in) { // the shared path's bookkeeping at this item's fork point
      const int alive_in = reinterpret_cast<const int *>(rec_in)[2];
      if (!alive_in) {
        collided = true;
        steps_done = reinterpret_cast<const int *>(rec_in)[3];
      } else {
        social_work = *reinterpret_cast<const double *>(rec_in);
      }
    }
    const int i_end = collided ? 0 : min(S_eff, n_run);
    int written = s0; // writer items: last step count whose record holds a live state
    for (int i = s0; i < i_end; ++i) {
      // robot as the SFM sees it during computeForces of this step (previous pose)
      const float prx = (i == 0) ? scp->ax : (float)(sm.rx[i] - base_x);
      const float pry = (i == 0) ? scp->ay : (float)(sm.ry[i] - base_y);
      const float rvxf = (i == 0) ? scp->avx : sm.rvx[i];
      const float rvyf = (i == 0) ? scp->avy : (float)vy;
      const f2 RX = bc2(prx), RY = bc2(pry), RVX = bc2(rvxf), RVY = bc2(rvyf);
      f2 rfx2 = bc2(0.f), rfy2 = bc2(0.f), wp2 = bc2(0.f);

      // -- phase 0: group forces (lightsfm computeGroupForce), one member per thread, into the warp's row --
      if (n_groups) {
        const uint32_t *gt = B.groups + scp->grp_off;
        const uint32_t *mem0 = gt + n_groups + 1u;
        const uint32_t n_members = gt[n_groups];
        const float *Pf = reinterpret_cast<const float *>(sm.pos);
        float *Ff = reinterpret_cast<float *>(myrow);
        for (uint32_t q = tid; q < n_members; q += kCrowdThreads) {
          uint32_t g = 0;
          while (gt[g + 1u] <= q)
            ++g;
          const uint32_t s0 = gt[g], c = gt[g + 1u] - s0;
          const uint32_t *mem = mem0 + 2u * s0;
          float cx, cy, ddx, ddy, gfx, gfy;
          group_centre(mem, c, Pf, 4u, cx, cy);
          const uint32_t j = mem[2u * (q - s0)];
          desired_direction(Pf, 4u, sm.goal, sm.par, j, sm.goalflag[j] != 0, ddx, ddy);
          group_member_force(mem, c, q - s0, Pf, 4u, cx, cy, ddx, ddy, B.k_gaze, B.k_coh, B.k_rep, gfx, gfy);
          Ff[(j >> 1) * 4u + (j & 1u)] += gfx;
          Ff[(j >> 1) * 4u + 2u + (j & 1u)] += gfy;
        }
        __syncwarp();
      }
      // the robot's sums of this warp (its own social force, the pedestrians' social work terms) -> sm.red
      auto reduce_robot_sums = [&]() {
        float l, h;
        un2(rfx2, l, h);
        const float rfx = warp_sum(l + h);
        un2(rfy2, l, h);
        const float rfy = warp_sum(l + h);
        un2(wp2, l, h);
        const float wp = warp_sum(l + h);
        if (lane == 0) {
          sm.red[warp * 4 + 0] = rfx;
          sm.red[warp * 4 + 1] = rfy;
          sm.red[warp * 4 + 2] = wp;
        }
      };
      // -- phase 1: forces (sfw_planner.cpp:592) --
      if (spread) {
        const uint32_t a = lane;
        const bool act = a < P2;
        float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), va = pa;
        if (act) {
          pa = sm.pos[a];
          va = sm.vel[a];
        }
        f2 FX = bc2(0.f), FY = bc2(0.f);
        if (warp == 0u) {
          if (act) {
            f2 fx, fy, fm;
            pair_force2<true>(K, mk2(pa.x, pa.y), mk2(pa.z, pa.w), mk2(va.x, va.y), mk2(va.z, va.w), RX, RY, RVX,
                              RVY, fx, fy, fm);
            FX = fx;
            FY = fy;
            rfx2 = sub2(rfx2, fx);
            rfy2 = sub2(rfy2, fy);
            wp2 = add2(wp2, fm);
            float gx_, gy_, gm_;
            pair_force<false>(K, pa.x, pa.z, va.x, va.z, pa.y, pa.w, va.y, va.w, gx_, gy_, gm_);
            FX = add2(FX, mk2(gx_, -gx_));
            FY = add2(FY, mk2(gy_, -gy_));
          }
        } else if (warp < kSpreadWarps) {
          const f2 A0X = bc2(pa.x), A0Y = bc2(pa.z), A0VX = bc2(va.x), A0VY = bc2(va.z);
          const f2 A1X = bc2(pa.y), A1Y = bc2(pa.w), A1VX = bc2(va.y), A1VY = bc2(va.w);
          f2 s0x = bc2(0.f), s0y = bc2(0.f), s1x = bc2(0.f), s1y = bc2(0.f);
          const uint32_t n_off = half + (even ? 1u : 0u);
          for (uint32_t off = warp; off <= n_off; off += kSpreadWarps - 1u) {
            const bool go = act && (off <= half || a < P2 / 2u);
            if (go) {
              uint32_t j = a + off;
              if (j >= P2)
                j -= P2;
              const float4 pb = sm.pos[j], vb = sm.vel[j];
              const f2 BX = mk2(pb.x, pb.y), BY = mk2(pb.z, pb.w);
              const f2 BVX = mk2(vb.x, vb.y), BVY = mk2(vb.z, vb.w);
              f2 hx, hy, gx2, gy2, hm;
              pair_force2<false>(K, A0X, A0Y, A0VX, A0VY, BX, BY, BVX, BVY, hx, hy, hm);
              pair_force2<false>(K, A1X, A1Y, A1VX, A1VY, BX, BY, BVX, BVY, gx2, gy2, hm);
              s0x = add2(s0x, hx);
              s0y = add2(s0y, hy);
              s1x = add2(s1x, gx2);
              s1y = add2(s1y, gy2);
              const float4 fb4 = myrow[j]; // within one offset the lanes hit distinct entries
              float bx0, bx1, by0, by1;
              un2(sub2(mk2(fb4.x, fb4.y), add2(hx, gx2)), bx0, bx1);
              un2(sub2(mk2(fb4.z, fb4.w), add2(hy, gy2)), by0, by1);
              myrow[j] = make_float4(bx0, bx1, by0, by1);
            }
            __syncwarp();
          }
          if (act) {
            float l0, h0, l1, h1;
            un2(s0x, l0, h0);
            un2(s1x, l1, h1);
            FX = mk2(l0 + h0, l1 + h1);
            un2(s0y, l0, h0);
            un2(s1y, l1, h1);
            FY = mk2(l0 + h0, l1 + h1);
            if (M) {
              f2 ox, oy;
              obstacle_sum2(sm.obs, (int)M, B.c_obs, mk2(pa.x, pa.y), mk2(pa.z, pa.w), ox, oy,
                            (int)kSpreadWarps - 1 - (int)warp, (int)kSpreadWarps - 1);
              const float4 Pc = sm.par2[a];
              const f2 OS = mk2(Pc.x, Pc.y);
              FX = fma2(OS, ox, FX);
              FY = fma2(OS, oy, FY);
            }
          }
        }
        if (act && warp < kSpreadWarps) {
          const float4 own = myrow[a];
          float x0, x1, y0, y1;
          un2(add2(mk2(own.x, own.y), FX), x0, x1);
          un2(add2(mk2(own.z, own.w), FY), y0, y1);
          myrow[a] = make_float4(x0, x1, y0, y1);
        }
        if (warp == 0u) // the only warp that met the robot in this layout
          reduce_robot_sums();
        else if (lane < 3u)
          sm.red[warp * 4 + lane] = 0.f;
      } else {
      // pass A: every owned pair against the robot and inside itself, added to the pair's own row entry; the robot's
      // sums leave the registers before the ring walk (the walk runs at the register limit: ten live registers
      // fewer let the compiler keep its loop-carried sums in place)
      for (uint32_t m = 0; m < owned; ++m) {
        const uint32_t a = tid + m * kCrowdThreads;
        if (a < P2) {
          const float4 pa = sm.pos[a], va = sm.vel[a];
          f2 fx, fy, fm;
          pair_force2<true>(K, mk2(pa.x, pa.y), mk2(pa.z, pa.w), mk2(va.x, va.y), mk2(va.z, va.w), RX, RY, RVX,
                            RVY, fx, fy, fm);
          rfx2 = sub2(rfx2, fx);
          rfy2 = sub2(rfy2, fy);
          wp2 = add2(wp2, fm);
          float gx_, gy_, gm_;
          pair_force<false>(K, pa.x, pa.z, va.x, va.z, pa.y, pa.w, va.y, va.w, gx_, gy_, gm_);
          const float4 own = myrow[a];
          float x0, x1, y0, y1;
          un2(add2(mk2(own.x, own.y), add2(fx, mk2(gx_, -gx_))), x0, x1);
          un2(add2(mk2(own.z, own.w), add2(fy, mk2(gy_, -gy_))), y0, y1);
          myrow[a] = make_float4(x0, x1, y0, y1);
        }
      }
      reduce_robot_sums();
      __syncwarp();
      // pass B: the cyclic cross-pair walk
      for (uint32_t m = 0; m < owned; ++m) {
        const uint32_t a = tid + m * kCrowdThreads;
        const bool act = a < P2;
        // Lanes past the last pair (the tail of the last warp) walk the ring as pair 0 and never store: the loop
        // below then has no divergent region at all — one basic block per cyclic offset.
        const uint32_t ar = act ? a : 0u;
        const float4 pa = sm.pos[ar], va = sm.vel[ar];
        f2 s0x = bc2(0.f), s0y = bc2(0.f), s1x = bc2(0.f), s1y = bc2(0.f);
        const f2 A0X = bc2(pa.x), A0Y = bc2(pa.z), A0VX = bc2(va.x), A0VY = bc2(va.z);
        const f2 A1X = bc2(pa.y), A1Y = bc2(pa.w), A1VX = bc2(va.y), A1VY = bc2(va.w);
        // one cyclic offset: pair a against pair j (state already in registers), reaction pushed into the warp's row
        // GATED: lanes with `push` false leave no trace (the opposite-pair trip, once per step); otherwise every
        // lane runs the same instructions and only the store of the lanes past the last pair is predicated off
        // (they read their "row entry" from a location nobody writes in this phase).
        auto trip = [&](const float4 pb, const float4 vb, const uint32_t j, const bool push, auto gated) {
          constexpr bool GATED = decltype(gated)::value;
          if (GATED && !push)
            return;
          const float4 *src = (GATED || push) ? myrow + j : sm.pos;
          const float4 fb4 = *src; // within one offset the lanes hit distinct entries
          const f2 BX = mk2(pb.x, pb.y), BY = mk2(pb.z, pb.w);
          const f2 BVX = mk2(vb.x, vb.y), BVY = mk2(vb.z, vb.w);
          f2 hx, hy, gx2, gy2, hm;
          pair_force2<false>(K, A0X, A0Y, A0VX, A0VY, BX, BY, BVX, BVY, hx, hy, hm);
          pair_force2<false>(K, A1X, A1Y, A1VX, A1VY, BX, BY, BVX, BVY, gx2, gy2, hm);
          s0x = add2(s0x, hx);
          s0y = add2(s0y, hy);
          s1x = add2(s1x, gx2);
          s1y = add2(s1y, gy2);
          float bx0, bx1, by0, by1;
          un2(sub2(mk2(fb4.x, fb4.y), add2(hx, gx2)), bx0, bx1);
          un2(sub2(mk2(fb4.z, fb4.w), add2(hy, gy2)), by0, by1);
          if (push)
            myrow[j] = make_float4(bx0, bx1, by0, by1);
        };
        // The ring is walked with the NEXT pair's state fetched one offset ahead: the reaction push of an offset
        // (a shared-memory read-modify-write the compiler must keep in order with every other shared access) would
        // otherwise put the load latency of pos / vel at the head of every trip.
        // (a warp with no pair at all — small crowds leave most of the block without one — skips the walk: its
        // lanes would only take issue slots from the warps that own pairs)
        const bool warp_owns = __any_sync(0xffffffffu, act); // a vote: the compiler sees a uniform branch
        uint32_t j = ar + 1u;
        if (j >= P2)
          j -= P2;
        float4 pb = make_float4(0.f, 0.f, 0.f, 0.f), vb = pb;
        if (warp_owns) {
          pb = sm.pos[j];
          vb = sm.vel[j];
          for (uint32_t off = 1; off <= half; ++off) {
            uint32_t jn = j + 1u;
            if (jn >= P2)
              jn -= P2;
            const float4 pbn = sm.pos[jn], vbn = sm.vel[jn];
            trip(pb, vb, j, act, std::false_type{});
            __syncwarp(); // lane t+1's push to entry j lands before lane t reaches it one offset later
            j = jn;
            pb = pbn;
            vb = vbn;
          }
        }
        if (even) { // the opposite pair of the ring: first half of the ring only
          trip(pb, vb, j, act && a < P2 / 2u, std::true_type{});
          __syncwarp();
        }
        if (act) {
          float l0, h0, l1, h1;
          un2(s0x, l0, h0);
          un2(s1x, l1, h1);
          f2 FX = mk2(l0 + h0, l1 + h1);
          un2(s0y, l0, h0);
          un2(s1y, l1, h1);
          f2 FY = mk2(l0 + h0, l1 + h1);
          if (!n_help) {
            f2 ox, oy;
            obstacle_sum2(sm.obs, (int)M, B.c_obs, mk2(pa.x, pa.y), mk2(pa.z, pa.w), ox, oy);
            const float4 Pc = sm.par2[a];
            const f2 OS = mk2(Pc.x, Pc.y);
            FX = fma2(OS, ox, FX);
            FY = fma2(OS, oy, FY);
          }
          const float4 own = myrow[a];
          float x0, x1, y0, y1;
          un2(add2(mk2(own.x, own.y), FX), x0, x1);
          un2(add2(mk2(own.z, own.w), FY), y0, y1);
          myrow[a] = make_float4(x0, x1, y0, y1);
        }
        __syncwarp();
      }
      }
      if (help_idx >= 1u && help_idx <= n_help && help_pair < P2 && M) {
        const float4 pa = sm.pos[help_pair];
        f2 ox, oy;
        obstacle_sum2(sm.obs, (int)M, B.c_obs, mk2(pa.x, pa.y), mk2(pa.z, pa.w), ox, oy, (int)help_idx - 1, (int)n_help);
        const float4 Pc = sm.par2[help_pair];
        const f2 OS = mk2(Pc.x, Pc.y);
        const float4 own = myrow[help_pair];
        float x0, x1, y0, y1;
        un2(fma2(OS, ox, mk2(own.x, own.y)), x0, x1);
        un2(fma2(OS, oy, mk2(own.z, own.w)), y0, y1);
        myrow[help_pair] = make_float4(x0, x1, y0, y1);
      }
      if (warp == (uint32_t)kCrowdWarps - 1u) { // the robot's obstacle force: off the critical path of phase 2
        __syncwarp();
        float rox, roy;
        const float qrx = (i == 0) ? scp->ax : (float)(sm.rx[i] - base_x), qry = (i == 0) ? scp->ay : (float)(sm.ry[i] - base_y);
        robot_obstacle_sum_warp(sm.obs, M, B.c_obs, qrx, qry, lane, rox, roy);
        if (lane == 0u) {
          sm.red[3] = rox * a_obs_scale;
          sm.red[7] = roy * a_obs_scale;
        }
      }
      __syncthreads();

      // -- phase 2: updatePosition (:594), collision (:613-627), social work (:629) --
      bool hit = false;
      const f2 DT = bc2(dtf);
      // (computed here, not at the top of the step: nothing that the ring walk does not need stays live across it)
      const float nrx = (float)(sm.rx[i + 1] - base_x), nry = (float)(sm.ry[i + 1] - base_y);
      for (uint32_t m = 0; m < owned; ++m) {
        const uint32_t a = tid + m * kCrowdThreads;
        if (a >= P2)
          break;
        // One warp alone walks this chain in a small crowd, so its length is the step's: all rows are loaded before
        // the first is cleared (a store between two loads pins their order), and — lanes are different PEDESTRIANS
        // here — the conditional rsqrt / goal-flag updates are selects, so that no data-dependent branch can split
        // the warp.
        f2 FX = bc2(0.f), FY = bc2(0.f);
        {
          float4 f[kCrowdWarps];
#pragma unroll
          for (uint32_t w = 0; w < (uint32_t)kCrowdWarps; ++w)
            f[w] = sm.frc[w * P2 + a];
#pragma unroll
          for (uint32_t w = 0; w < (uint32_t)kCrowdWarps; ++w) {
            FX = add2(FX, mk2(f[w].x, f[w].y));
            FY = add2(FY, mk2(f[w].z, f[w].w));
            sm.frc[w * P2 + a] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        const float4 pa = sm.pos[a], va = sm.vel[a];
        const f2 AX = mk2(pa.x, pa.y), AY = mk2(pa.z, pa.w), AVX = mk2(va.x, va.y), AVY = mk2(va.z, va.w);
        const float4 G = sm.goal[a], Pp = sm.par[a], Pc = sm.par2[a];
        const f2 gdx = sub2(mk2(G.x, G.y), AX), gdy = sub2(mk2(G.z, G.w), AY);
        float g20, g21;
        un2(fma2(gdx, gdx, mul2(gdy, gdy)), g20, g21);
        const bool hg0 = sm.goalflag[2 * a] != 0, hg1 = sm.goalflag[2 * a + 1] != 0;
        const bool go0 = hg0 && g20 > Pp.x, go1 = hg1 && g21 > Pp.y;
        const float rg0 = rsqrt_approx(g20) * Pp.z, rg1 = rsqrt_approx(g21) * Pp.w; // (inf / NaN at g2 == 0: not selected)
        const f2 gs = mk2(go0 ? rg0 : 0.f, go1 ? rg1 : 0.f);
        const f2 c1 = mk2(go0 ? B.kd_tau : B.inv_tau, go1 ? B.kd_tau : B.inv_tau);
        const f2 Fx = add2(mul2(c1, sub2(mul2(gdx, gs), AVX)), FX);
        const f2 Fy = add2(mul2(c1, sub2(mul2(gdy, gs), AVY)), FY);
        f2 nvx = fma2(Fx, DT, AVX), nvy = fma2(Fy, DT, AVY);
        float v20, v21;
        un2(fma2(nvx, nvx, mul2(nvy, nvy)), v20, v21);
        const float rs0 = Pp.z * rsqrt_approx(v20), rs1 = Pp.w * rsqrt_approx(v21);
        const f2 sc = mk2(v20 > Pc.z ? rs0 : 1.0f, v21 > Pc.w ? rs1 : 1.0f);
        nvx = mul2(nvx, sc);
        nvy = mul2(nvy, sc);
        const f2 npx = fma2(nvx, DT, AX), npy = fma2(nvy, DT, AY);
        float px0, px1, py0, py1, vx0, vx1, vy0, vy1;
        un2(npx, px0, px1);
        un2(npy, py0, py1);
        un2(nvx, vx0, vx1);
        un2(nvy, vy0, vy1);
        sm.pos[a] = make_float4(px0, px1, py0, py1);
        sm.vel[a] = make_float4(vx0, vx1, vy0, vy1);
        {
          const f2 hx = sub2(mk2(G.x, G.y), npx), hy = sub2(mk2(G.z, G.w), npy);
          float h0, h1;
          un2(fma2(hx, hx, mul2(hy, hy)), h0, h1);
          // goal reached -> pop: both flags of the pair in one 16-bit store, whatever they turn out to be
          const uint32_t f0 = (hg0 && !(h0 <= Pp.x)) ? 1u : 0u, f1 = (hg1 && !(h1 <= Pp.y)) ? 1u : 0u;
          *reinterpret_cast<uint16_t *>(sm.goalflag + 2 * a) = (uint16_t)(f0 | (f1 << 8));
        }
        {
          const f2 cx = sub2(bc2(nrx), npx), cy = sub2(bc2(nry), npy);
          float c0, c1_;
          un2(fma2(cx, cx, mul2(cy, cy)), c0, c1_);
          hit |= (c0 <= rr2) | (c1_ <= rr2);
        }
      }
      if (hit)
        sm.flags[1] = 1;
      if (warp == kWorkWarp) { // robot terms: wr from the forces of :592, wp belongs to the previous step
        float rfx = (lane < (uint32_t)kCrowdWarps) ? sm.red[lane * 4 + 0] : 0.f;
        float rfy = (lane < (uint32_t)kCrowdWarps) ? sm.red[lane * 4 + 1] : 0.f;
        float wp = (lane < (uint32_t)kCrowdWarps) ? sm.red[lane * 4 + 2] : 0.f;
        rfx = warp_sum(rfx);
        rfy = warp_sum(rfy);
        wp = warp_sum(wp);
        if (lane == 0) {
          const float rox = sm.red[3], roy = sm.red[7]; // summed by the last thread during phase 1
          const float wr = sqrt_approx(fmaf(rfx, rfx, rfy * rfy)) + sqrt_approx(fmaf(rox, rox, roy * roy));
          social_work += (double)(wr + ((i > 0) ? wp : 0.f));
        }
      }
      __syncthreads();
      steps_done = i + 1;
      if (sm.flags[1]) {
        collided = true;
        break;
      }
      if (writer) { // the path's state after i + 1 steps
        uint8_t *rec = rec_out + (size_t)(i + 1) * B.share.rec_bytes;
        float4 *pv = reinterpret_cast<float4 *>(rec + 16);
        uint8_t *gf = rec + 16 + 32u * (size_t)P2;
        for (uint32_t k = tid; k < P2; k += kCrowdThreads) {
          pv[k] = sm.pos[k];
          pv[P2 + k] = sm.vel[k];
          gf[2 * k] = sm.goalflag[2 * k];
          gf[2 * k + 1] = sm.goalflag[2 * k + 1];
        }
        if (tid == kWorkTid) {
          *reinterpret_cast<double *>(rec) = social_work;
          reinterpret_cast<int *>(rec)[2] = 1;
          reinterpret_cast<int *>(rec)[3] = i + 1;
        }
        written = i + 1;
      }
    }
    if (writer) {
      // step counts the path did not reach alive: a collision kills every descendant; an illegal pose is seen
      // by the descendants' own footprint check (same poses), they only must not read a stale record
      if (tid == kWorkTid)
        for (int k = written + 1; k <= (int)B.share.kmax; ++k) {
          uint8_t *rec = rec_out + (size_t)k * B.share.rec_bytes;
          *reinterpret_cast<double *>(rec) = social_work;
          reinterpret_cast<int *>(rec)[2] = collided ? 0 : 1;
          reinterpret_cast<int *>(rec)[3] = collided ? steps_done : S_eff;
        }
      continue; // shared paths have no cost of their own
    }

    // ---- terminal costs (sfw_planner.cpp:643-675) -----------------------------------------------
    const bool valid = !collided && S_eff == S;
    if (valid) {
      // computeSocialWork's pedestrian term of the last step (updated states, robot at final pose)
      const f2 RX = bc2((float)(sm.rx[S] - base_x)), RY = bc2((float)(sm.ry[S] - base_y));
      const f2 RVX = bc2(sm.rvx[S]), RVY = bc2((float)vy);
      f2 wp2 = bc2(0.f);
      for (uint32_t a = tid; a < P2; a += kCrowdThreads) {
        const float4 pa = sm.pos[a], va = sm.vel[a];
        f2 fx, fy, fm;
        pair_force2<true>(K, mk2(pa.x, pa.y), mk2(pa.z, pa.w), mk2(va.x, va.y), mk2(va.z, va.w), RX, RY, RVX, RVY,
                          fx, fy, fm);
        wp2 = add2(wp2, fm);
      }
      float l, h;
      un2(wp2, l, h);
      const float wp = warp_sum(l + h);
      if (lane == 0)
        sm.red[warp * 4 + 3] = wp;
    }
    __syncthreads();
    if (tid == kWorkTid) {
      float cost = SFW_COST_INVALID;
      // points recorded: one per legal pose until the first violation (sfw_planner.cpp:578)
      int npts = collided ? steps_done : (S_eff < S ? S_eff : S);
      if (valid) {
        float wp = 0.f;
        for (int w = 0; w < kCrowdWarps; ++w)
          wp += sm.red[w * 4 + 3];
        social_work += (double)wp;
        double costmap_sum = 0.0;
        for (int i = 0; i < S; ++i)
          costmap_sum = __dadd_rn(costmap_sum, __ddiv_rn((double)sm.fcm[i], 255.0));
        const double x = sm.rx[S], y = sm.ry[S], th = sm.rth[S];
        const double v_end = B.linvels[idx / n_w]; // (a sample, not a shared path: reloaded, not kept live)
        double vx = scp->rvx;
        for (int i = 0; i < S; ++i)
          vx = step_velocity(v_end, vx, ax_dt);
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
      B.costs[out] = cost;
      if (B.costs_host)
        B.costs_host[out] = cost;
      B.npts[out] = (uint16_t)npts;
    }
  }
  // ---- epilogue: the last block to run out of work re-arms the two counters (work_counter[0]: next item,
  // [1]: blocks done) for the next launch and, for small batches, reduces every scene's winner itself: no memset
  // before and no arg-min launch after the kernel (a 5 x 9 tick is 40 x 3 us of kernel: two extra stream
  // operations are 5 % of it).
  __shared__ int s_last;
  if (tid == 0) {
    __threadfence(); // this block's costs before its arrival
    const unsigned int prev = atomicAdd(work_counter + 1, 1u);
    s_last = (prev == gridDim.x - 1u) ? 1 : 0;
    if (s_last) {
      __threadfence();
      work_counter[0] = 0u;
      work_counter[1] = 0u;
    }
  }
  __syncthreads();
  if (s_last && fused_argmin) {
    float *s_c = reinterpret_cast<float *>(smem_raw);
    uint32_t *s_i = reinterpret_cast<uint32_t *>(smem_raw + 64);
    for (uint32_t scene = 0; scene < B.n_scenes; ++scene)
      scene_argmin(B, scene, s_c, s_i);
  }
}

// ================================================================================================
cudaError_t sfw_crowd_prepare(uint32_t threads, size_t smem_bytes, int *blocks_per_sm) {
  auto prep = [&](auto kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess)
      return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, (int)threads, smem_bytes);
  };
  return threads == SFW_CROWD_THREADS_SMALL ? prep(sfw_score_crowd<SFW_CROWD_THREADS_SMALL>)
                                            : prep(sfw_score_crowd<SFW_CROWD_THREADS>);
}

bool sfw_crowd_fuses_argmin(const SfwBatchDev &B) {
  const uint64_t values = (uint64_t)B.n_scenes * (B.row_end - B.row_begin) * B.n_w;
  return values <= 8192u && B.n_scenes <= 64u;
}

// `work_counter`: two words that are 0 before the first launch (sfw_upload clears them) — the kernel re-arms them.
// Small batches (sfw_crowd_fuses_argmin: at most 8192 cost values in at most 64 scenes) are reduced by the kernel's last block, larger ones by
// sfw_argmin_kernel with one block per scene.
cudaError_t sfw_launch_crowd(const SfwBatchDev &B, unsigned int *work_counter, uint32_t grid, uint32_t threads,
                             size_t smem_bytes, cudaStream_t stream, bool with_argmin) {
  const bool fused = with_argmin && sfw_crowd_fuses_argmin(B);
  if (threads == SFW_CROWD_THREADS_SMALL)
    sfw_score_crowd<SFW_CROWD_THREADS_SMALL><<<grid, threads, smem_bytes, stream>>>(B, work_counter, fused ? 1 : 0);
  else
    sfw_score_crowd<SFW_CROWD_THREADS><<<grid, threads, smem_bytes, stream>>>(B, work_counter, fused ? 1 : 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || !with_argmin || fused)
    return e;
  sfw_argmin_kernel<<<B.n_scenes, 256, 0, stream>>>(B);
  return cudaGetLastError();
}
