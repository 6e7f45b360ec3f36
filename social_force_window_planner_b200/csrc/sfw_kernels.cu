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
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "sfw_dev.h"
#include "sfw_kernels.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

// ------------------------------------------------------------------------------------------------
// PTX helpers: fast MUFU ops, mbarrier, TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map); bytes % 16 == 0, 16 B aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 3-D tensor tile global -> shared: the costmap window (x0, y0, scene)
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, int x, int y, int z,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
      "%3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// lightsfm pair social force in FP32.  Force on agent a from agent b, already
// scaled by forceFactorSocial.  e_ang is the magnitude of the angular term, used for |F|.
// ------------------------------------------------------------------------------------------------
struct SfmConst {
  float lambda, c_d, g2, c_np, c_n, k_soc; // g2 = gamma^2
};

// atan(t) for t in [0,1]: odd minimax polynomial, |err| < 1e-7 in FP32 (fit in DESIGN.md)
__device__ __forceinline__ float atan01(float t) {
  const float u = t * t;
  float p = 0.0024567015934735537f;
  p = fmaf(p, u, -0.01440125796943903f);
  p = fmaf(p, u, 0.03978104144334793f);
  p = fmaf(p, u, -0.07234840095043182f);
  p = fmaf(p, u, 0.10498936474323273f);
  p = fmaf(p, u, -0.14161226153373718f);
  p = fmaf(p, u, 0.19985906779766083f);
  p = fmaf(p, u, -0.33332598209381104f);
  p = fmaf(p, u, 0.9999998807907104f);
  return p * t;
}

// WITH_MAG: also return |F| (only the robot-pedestrian pairs need it, for the social work).
template <bool WITH_MAG>
__device__ __forceinline__ void pair_force(const SfmConst &K, float ax, float ay, float avx,
                                           float avy, float bx, float by, float bvx, float bvy,
                                           float &fx, float &fy, float &fmag) {
  const float dx = bx - ax, dy = by - ay;
  const float d2 = fmaf(dx, dx, fmaf(dy, dy, 1e-30f));    // +eps: coincident agents give 0, not NaN
  const float rd = rsqrt_approx(d2);
  const float ex = dx * rd, ey = dy * rd;                 // diffDirection
  const float ix = fmaf(K.lambda, avx - bvx, ex);         // interactionVector
  const float iy = fmaf(K.lambda, avy - bvy, ey);
  const float L2 = fmaf(ix, ix, fmaf(iy, iy, 1e-30f));
  const float rL = rsqrt_approx(L2);                      // 1 / interactionLength
  // theta = angle from interactionDirection to diffDirection = atan2(i x e, i . e).  The cross
  // product is a difference of two ROUNDED products (no FMA) so that equal velocities (i == e
  // bit for bit) give exactly theta = 0 and no angular term, as lightsfm's Angle::sign() does.
  const float sn = __fsub_rn(__fmul_rn(ix, ey), __fmul_rn(iy, ex));
  const float cs = fmaf(ix, ex, iy * ey);
  const float asn = fabsf(sn), acs = fabsf(cs);
  const float mx = fmaxf(asn, acs), mn = fminf(asn, acs);
  float th = atan01(mn * rcp_approx(fmaxf(mx, 1e-30f)));
  th = (asn > acs) ? (1.5707963267948966f - th) : th;
  th = (cs < 0.0f) ? (3.14159265358979f - th) : th;       // |theta| in [0, pi]
  const float q = (K.g2 * L2) * (th * th);                // (B theta)^2, B = gamma * interactionLength
  const float t = -((d2 * rd) * rL) * K.c_d;              // -|diff| / B  (in log2 units)
  const float e_vel = ex2_approx(fmaf(-K.c_np, q, t));    // exp(-d/B - (n' B theta)^2)
  const float e_ang = ex2_approx(fmaf(-K.c_n, q, t));     // exp(-d/B - (n  B theta)^2)
  const float rLk = rL * K.k_soc;
  const float fv = e_vel * rLk;                           // -(force along interactionVector)
  // forceAngle = -sign(theta) * e_ang along the left normal.  sign(theta) is 0 only for theta == 0
  // exactly; theta == pi counts as positive (lightsfm Angle wraps to (-pi, pi]).
  float fa = __uint_as_float(__float_as_uint(e_ang * rLk) | (__float_as_uint(sn) & 0x80000000u));
  if (sn == 0.0f)
    fa = (cs < 0.0f) ? fabsf(fa) : 0.0f;                  // here fa = +sign(theta) * magnitude
  // F = -fv * i - fa * leftNormal(i), leftNormal(i) = (-iy, ix)
  fx = fmaf(fa, iy, -(fv * ix));
  fy = -fmaf(fa, ix, fv * iy);
  if (WITH_MAG) {
    const float ea = (sn == 0.0f && cs >= 0.0f) ? 0.0f : e_ang;
    fmag = K.k_soc * sqrt_approx(fmaf(e_vel, e_vel, ea * ea));
  }
}

// lightsfm obstacle force sum (unscaled): sum_o exp(-|p-o|/sigma) (p-o)/|p-o|.  Obstacle points are
// stored pre-multiplied by c_obs = log2(e)/sigma, and so is the query point: then |p'-o'| is the
// exponent in log2 units and the unit vector is unchanged.
__device__ __forceinline__ void obstacle_sum(const float2 *__restrict__ obs, int M, float c_obs,
                                             float px, float py, float &sx, float &sy) {
  float ax = 0.f, ay = 0.f;
  const float qx = px * c_obs, qy = py * c_obs;
#pragma unroll 4
  for (int o = 0; o < M; ++o) {
    const float2 p = obs[o];
    const float dx = qx - p.x, dy = qy - p.y;
    const float d2 = fmaf(dx, dx, fmaf(dy, dy, 1e-30f));
    const float rd = rsqrt_approx(d2);
    const float e = ex2_approx(-(d2 * rd)) * rd;
    ax = fmaf(e, dx, ax);
    ay = fmaf(e, dy, ay);
  }
  sx = ax;
  sy = ay;
}

// ------------------------------------------------------------------------------------------------
// Costmap access + footprint rasterisation (bit-faithful integer work)
// ------------------------------------------------------------------------------------------------
struct MapView {
  const uint8_t *win;   // staged window in shared memory (or nullptr)
  const uint8_t *glob;  // this scene's costmap slot in HBM
  double ox, oy, res;
  uint32_t sx, sy, pitch;
  int32_t wx0, wy0;
  uint32_t wwp, wh;     // window pitch / rows
};

__device__ __forceinline__ uint32_t cell_cost(const MapView &m, int cx, int cy) {
  const uint32_t lx = (uint32_t)(cx - m.wx0), ly = (uint32_t)(cy - m.wy0);
  if (lx < m.wwp && ly < m.wh)
    return m.win[ly * m.wwp + lx];
  return __ldg(m.glob + (size_t)cy * m.pitch + cx);
}

// nav2 Costmap2D::worldToMap [external, SURVEY.md App. C]; no FMA contraction, true division
__device__ __forceinline__ bool world_to_map(const MapView &m, double wx, double wy, int &mx, int &my) {
  if (wx < m.ox || wy < m.oy)
    return false;
  const unsigned int ux = (unsigned int)__ddiv_rn(__dsub_rn(wx, m.ox), m.res);
  const unsigned int uy = (unsigned int)__ddiv_rn(__dsub_rn(wy, m.oy), m.res);
  mx = (int)ux;
  my = (int)uy;
  return ux < m.sx && uy < m.sy;
}

// CostmapModel::lineCost over LineIterator (costmap_model.cpp:95-110, line_iterator.hpp:37-97).
// Returns the max cell cost, or -1 when a 254/255 cell is met.
__device__ __forceinline__ int line_cost(const MapView &m, int x0, int y0, int x1, int y1) {
  const int deltax = abs(x1 - x0), deltay = abs(y1 - y0);
  int xinc1 = (x1 >= x0) ? 1 : -1, xinc2 = xinc1;
  int yinc1 = (y1 >= y0) ? 1 : -1, yinc2 = yinc1;
  int den, num, numadd, numpixels;
  if (deltax >= deltay) {
    xinc1 = 0;
    yinc2 = 0;
    den = deltax;
    num = deltax / 2;
    numadd = deltay;
    numpixels = deltax;
  } else {
    xinc2 = 0;
    yinc1 = 0;
    den = deltay;
    num = deltay / 2;
    numadd = deltax;
    numpixels = deltay;
  }
  int x = x0, y = y0, worst = 0;
  for (int cur = 0; cur <= numpixels; ++cur) {
    const int c = (int)cell_cost(m, x, y);
    if (c >= 254)
      return -1;
    worst = max(worst, c);
    num += numadd;
    if (num >= den) {
      num -= den;
      x += xinc1;
      y += yinc1;
    }
    x += xinc2;
    y += yinc2;
  }
  return worst;
}

// WorldModel::footprintCost(x,y,theta,spec) + CostmapModel::footprintCost
// (world_model.hpp:45-75, costmap_model.cpp:21-92).  Returns 0..253, or -1 for any of the
// reference's negative codes (-1/-2/-3 are all "invalid" to scoreTrajectory, cpp:555-573).
__device__ __forceinline__ int footprint_cost(const MapView &m, const double2 *__restrict__ fp, int F,
                                              double x, double y, double sn, double cs) {
  int cx, cy;
  if (!world_to_map(m, x, y, cx, cy))
    return -1;
  if (F < 3) {
    const int c = (int)cell_cost(m, cx, cy);
    return (c >= 253) ? -1 : c;
  }
  int worst = 0;
  int fx0 = 0, fy0 = 0, px = 0, py = 0;
  for (int i = 0; i < F; ++i) {
    const double2 v = fp[i];
    // world_model.hpp:56-59, evaluated without contraction
    const double wx = __dadd_rn(x, __dsub_rn(__dmul_rn(v.x, cs), __dmul_rn(v.y, sn)));
    const double wy = __dadd_rn(y, __dadd_rn(__dmul_rn(v.x, sn), __dmul_rn(v.y, cs)));
    int mx, my;
    if (!world_to_map(m, wx, wy, mx, my))
      return -1;
    if (i == 0) {
      fx0 = mx;
      fy0 = my;
    } else {
      const int lc = line_cost(m, px, py, mx, my);
      if (lc < 0)
        return -1;
      worst = max(worst, lc);
    }
    px = mx;
    py = my;
  }
  const int lc = line_cost(m, px, py, fx0, fy0); // closing edge last -> first
  if (lc < 0)
    return -1;
  return max(worst, lc);
}

// sfw_planner.hpp:457-463
__device__ __forceinline__ double step_velocity(double vg, double vi, double a_dt) {
  if (__dsub_rn(vg, vi) >= 0.0)
    return fmin(vg, __dadd_rn(vi, a_dt));
  return fmax(vg, __dsub_rn(vi, a_dt));
}

// sfw_planner.hpp:399-407 (all-float arithmetic, fmodf)
__device__ __forceinline__ float normalize_angle_f(float val, float mn, float mx) {
  if (val >= mn)
    return mn + fmodf(val - mn, mx - mn);
  return mx - fmodf(mn - val, mx - mn);
}

// Total order of the reference's sequential best-update (sfw_planner.cpp:394-414):
// lower cost, then higher linvel, then lower |angvel|, then later index.
__device__ __forceinline__ bool better(float ca, uint32_t ia, float cb, uint32_t ib,
                                       const double *__restrict__ lin,
                                       const double *__restrict__ ang, uint32_t n_w) {
  if (cb < 0.f)
    return ca >= 0.f;
  if (ca < 0.f)
    return false;
  if (ca != cb)
    return ca < cb;
  const double la = lin[ia / n_w], lb = lin[ib / n_w];
  if (la != lb)
    return la > lb;
  const double wa = fabs(ang[ia % n_w]), wb = fabs(ang[ib % n_w]);
  if (wa != wb)
    return wa < wb;
  return ia > ib;
}

// A cost can only become "best" if 0 <= cost <= 10000; == 10000 needs linvel > 0 because the
// initial best is (10000, xv = 0, thetav = 0) (sfw_planner.cpp:338-344,394-407).
__device__ __forceinline__ bool eligible(float c, double linvel) {
  return c >= 0.f && (c < 10000.f || (c == 10000.f && linvel > 0.0));
}

} // namespace

// ================================================================================================
// Kernel: one thread per trajectory
// ================================================================================================
extern "C" __global__ void __launch_bounds__(SFW_MAX_BLOCK_SMALL, 1)
sfw_score_small(const __grid_constant__ SfwBatchDev B, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t T = blockDim.x;
  const uint32_t tid = threadIdx.x;
  const uint32_t scene = blockIdx.x / B.tiles_per_scene;
  const uint32_t tile = blockIdx.x - scene * B.tiles_per_scene;
  const SfwSceneDev *__restrict__ scp = B.scenes + scene;
  const uint32_t P = scp->n_peds, M = scp->n_obst, F = scp->n_fp;

  // ---- shared memory carve-up -------------------------------------------------------------
  // [window][pedA P][pedB P][pedC P][obst M (+pad)][footprint F][mbar][tile best][state P*T][force P*T]
  const uint32_t win_bytes = B.win_wp * B.win_h; // multiple of 16
  uint8_t *s_win = smem_raw;
  uint32_t off = (win_bytes + 127u) & ~127u;
  float4 *s_pedA = reinterpret_cast<float4 *>(smem_raw + off);
  off += P * 16u;
  float4 *s_pedB = reinterpret_cast<float4 *>(smem_raw + off);
  off += P * 16u;
  float4 *s_pedC = reinterpret_cast<float4 *>(smem_raw + off);
  off += P * 16u;
  float2 *s_obs = reinterpret_cast<float2 *>(smem_raw + off);
  off += ((M * 8u) + 15u) & ~15u;
  double2 *s_fp = reinterpret_cast<double2 *>(smem_raw + off);
  off += F * 16u;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + off);
  off += 16u;
  float *s_redc = reinterpret_cast<float *>(smem_raw + off);
  uint32_t *s_redi = reinterpret_cast<uint32_t *>(smem_raw + off + 32u * 4u);
  off += 64u * 4u;
  float4 *s_state = reinterpret_cast<float4 *>(smem_raw + off);
  off += P * T * 16u;
  float2 *s_force = reinterpret_cast<float2 *>(smem_raw + off);

  __shared__ bool s_last;

  // ---- stage the scene with the TMA engine ---------------------------------------------------
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
    const uint32_t obs_bytes = ((M * 8u) + 15u) & ~15u;
    const uint32_t tx = win_bytes + 3u * P * 16u + obs_bytes + F * 16u;
    mbar_expect_tx(s_bar, tx);
    if (win_bytes)
      tma_load_3d(s_win, &tmap, scp->win_x0, scp->win_y0, (int)scene, s_bar);
    if (P) {
      bulk_g2s(s_pedA, B.pedA + scp->ped_off, P * 16u, s_bar);
      bulk_g2s(s_pedB, B.pedB + scp->ped_off, P * 16u, s_bar);
      bulk_g2s(s_pedC, B.pedC + scp->ped_off, P * 16u, s_bar);
    }
    if (obs_bytes)
      bulk_g2s(s_obs, B.obst + scp->obs_off, obs_bytes, s_bar);
    if (F)
      bulk_g2s(s_fp, B.footprint + scp->fp_off, F * 16u, s_bar);
  }
  __syncthreads(); // barrier init visible to the waiters
  mbar_wait(s_bar, 0);

  MapView mv;
  mv.win = win_bytes ? s_win : nullptr;
  mv.glob = B.maps + scp->map_off;
  mv.ox = scp->origin_x;
  mv.oy = scp->origin_y;
  mv.res = scp->resolution;
  mv.sx = scp->size_x;
  mv.sy = scp->size_y;
  mv.pitch = B.map_pitch;
  mv.wx0 = scp->win_x0;
  mv.wy0 = scp->win_y0;
  mv.wwp = win_bytes ? B.win_wp : 0u;
  mv.wh = win_bytes ? B.win_h : 0u;

  const SfmConst K = {B.lambda, B.c_d, B.gamma * B.gamma, B.c_np, B.c_n, B.k_soc};

  // ---- which trajectory is mine ----------------------------------------------------------------
  const uint32_t n_w = B.n_w;
  const uint32_t first = B.row_begin * n_w, last = B.row_end * n_w;
  const uint32_t idx = first + tile * T + tid;
  const bool in_range = idx < last;
  double v_s = 0.0, w_s = 0.0;
  if (in_range) {
    v_s = B.linvels[idx / n_w];
    w_s = B.angvels[idx % n_w];
  }
  const bool skipped = in_range && (v_s == 0.0 && w_s == 0.0); // sfw_planner.cpp:349-352
  bool alive = in_range && !skipped;

  // ---- per-thread state ------------------------------------------------------------------------
  float4 *st = s_state + tid;  // st[a*T]
  float2 *fr = s_force + tid;  // fr[a*T]
  for (uint32_t a = 0; a < P; ++a) {
    st[a * T] = s_pedA[a];
    fr[a * T] = make_float2(0.f, 0.f);
  }
  uint64_t goalmask = 0;
  for (uint32_t a = 0; a < P; ++a)
    if (s_pedC[a].y != 0.f)
      goalmask |= (1ull << a);

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

  for (int i = 0; i < S; ++i) {
    if (!__any_sync(0xffffffffu, alive))
      break;
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

        // -- social force step (sfw_planner.cpp:592-629) --
        float rfx = 0.f, rfy = 0.f, wp = 0.f;
        bool hit = false;
        for (uint32_t a = 0; a < P; ++a) {
          const float4 A = st[a * T];
          float2 fa = fr[a * T];
          float fx, fy, fm;
          // pedestrian a <- robot (robot at the pose the previous step left it)
          pair_force<true>(K, A.x, A.y, A.z, A.w, prx, pry, rvxf, rvyf, fx, fy, fm);
          fa.x += fx;
          fa.y += fy;
          rfx -= fx;
          rfy -= fy;
          wp += fm; // = computeSocialWork's per-pedestrian term of the PREVIOUS step
#pragma unroll 2
          for (uint32_t b = a + 1; b < P; ++b) {
            const float4 Bs = st[b * T];
            float gx_, gy_, gm_;
            pair_force<false>(K, A.x, A.y, A.z, A.w, Bs.x, Bs.y, Bs.z, Bs.w, gx_, gy_, gm_);
            fa.x += gx_;
            fa.y += gy_;
            float2 fb = fr[b * T];
            fb.x -= gx_;
            fb.y -= gy_;
            fr[b * T] = fb;
          }
          // obstacle force
          float ox, oy;
          obstacle_sum(s_obs, (int)M, B.c_obs, A.x, A.y, ox, oy);
          const float4 Bp = s_pedB[a];
          const float4 Cp = s_pedC[a];
          // desired force (App. B-1)
          float dfx, dfy;
          const float gx = Bp.x - A.x, gy = Bp.y - A.y;
          const float g2 = fmaf(gx, gx, gy * gy);
          const bool has_goal = (goalmask >> a) & 1ull;
          if (has_goal && g2 > Bp.z) {
            const float rg = rsqrt_approx(g2) * Bp.w;
            dfx = B.kd_tau * fmaf(gx, rg, -A.z);
            dfy = B.kd_tau * fmaf(gy, rg, -A.w);
          } else {
            dfx = -A.z * B.inv_tau;
            dfy = -A.w * B.inv_tau;
          }
          const float Fx = dfx + fa.x + Cp.x * ox;
          const float Fy = dfy + fa.y + Cp.x * oy;
          // updatePosition (App. B-5)
          float nvx = fmaf(Fx, dtf, A.z), nvy = fmaf(Fy, dtf, A.w);
          const float v2 = fmaf(nvx, nvx, nvy * nvy);
          if (v2 > Cp.z) {
            const float sc = Bp.w * rsqrt_approx(v2);
            nvx *= sc;
            nvy *= sc;
          }
          const float npx = fmaf(nvx, dtf, A.x), npy = fmaf(nvy, dtf, A.y);
          st[a * T] = make_float4(npx, npy, nvx, nvy);
          fr[a * T] = make_float2(0.f, 0.f);
          if (has_goal) {
            const float hx = Bp.x - npx, hy = Bp.y - npy;
            if (fmaf(hx, hx, hy * hy) <= Bp.z)
              goalmask &= ~(1ull << a);
          }
          // robot / pedestrian collision with the NEW robot pose (sfw_planner.cpp:613-627)
          const float cx = nrx - npx, cy = nry - npy;
          hit |= (fmaf(cx, cx, cy * cy) <= rr2);
        }
        // robot's own obstacle force at the pose the forces were evaluated at
        float rox, roy;
        obstacle_sum(s_obs, (int)M, B.c_obs, prx, pry, rox, roy);
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
  }

  // ---- terminal costs (sfw_planner.cpp:643-675) -----------------------------------------------
  float cost = in_range ? (skipped ? SFW_COST_SKIPPED : SFW_COST_INVALID) : SFW_COST_SKIPPED;
  if (alive) {
    // computeSocialWork's pedestrian term of the last step (updated states, robot at final pose)
    float wp = 0.f;
    for (uint32_t a = 0; a < P; ++a) {
      const float4 A = st[a * T];
      float fx, fy, fm;
      pair_force<true>(K, A.x, A.y, A.z, A.w, prx, pry, rvxf, rvyf, fx, fy, fm);
      wp += fm;
    }
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
        B.counters[scene] = 0u; // ready for the next launch
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

// ================================================================================================
// launch wrappers (called from sfw_abi.cu)
// ================================================================================================
size_t sfw_small_smem_bytes(uint32_t win_bytes, uint32_t P, uint32_t M, uint32_t F, uint32_t T) {
  size_t off = (win_bytes + 127u) & ~127u;
  off += 3u * (size_t)P * 16u;
  off += ((M * 8u) + 15u) & ~15u;
  off += (size_t)F * 16u;
  off += 16u;
  off += 64u * 4u;
  off += (size_t)P * T * 16u;
  off += (size_t)P * T * 8u;
  return off;
}

// Largest dynamic shared memory a block of sfw_score_small may request on the current device
// (opt-in limit minus the kernel's static shared memory); also opts the kernel in.
cudaError_t sfw_small_max_dynamic_smem(size_t *bytes) {
  static size_t cached = 0;
  if (!cached) {
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
      return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess)
      return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, sfw_score_small);
    if (e != cudaSuccess)
      return e;
    const size_t dyn = (size_t)optin - fa.sharedSizeBytes;
    e = cudaFuncSetAttribute(sfw_score_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess)
      return e;
    cached = dyn;
  }
  *bytes = cached;
  return cudaSuccess;
}

cudaError_t sfw_launch_small(const SfwBatchDev &B, const CUtensorMap &tmap, uint32_t T,
                             size_t smem_bytes, cudaStream_t stream) {
  const uint32_t grid = B.n_scenes * B.tiles_per_scene;
  sfw_score_small<<<grid, T, smem_bytes, stream>>>(B, tmap);
  return cudaGetLastError();
}

cudaError_t sfw_launch_points(const SfwBatchDev &B, uint32_t scene, uint32_t idx, uint32_t n_points,
                              double *out_xyz, cudaStream_t stream) {
  sfw_points_kernel<<<1, 32, 0, stream>>>(B, scene, idx, n_points, out_xyz);
  return cudaGetLastError();
}

cudaError_t sfw_small_occupancy(uint32_t T, size_t smem_bytes, int *blocks_per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, sfw_score_small, (int)T,
                                                       smem_bytes);
}
