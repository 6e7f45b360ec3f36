// sfw_forces.cuh — device functions shared by the scorer kernels (sfw_kernels.cu: thread per
// trajectory; sfw_crowd.cu: block per trajectory): PTX helpers (MUFU, mbarrier, TMA), packed FP32x2
// arithmetic, the lightsfm force evaluators, costmap access / footprint rasterisation, kinematics and
// the arg-min order.  Reference semantics (paths relative to the reference repo):
//   src/sfw_planner.cpp:338-417, :475-705, include/.../sfw_planner.hpp:399-463,
//   include/.../world_model.hpp:45-75, src/costmap_model.cpp:21-121,
//   include/.../line_iterator.hpp:37-124, lightsfm.
#ifndef SFW_FORCES_CUH
#define SFW_FORCES_CUH
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "sfw_dev.h"

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
  // try_wait suspends in hardware for a bounded time; back off between tries so that the spinning warps of a
  // block leave the issue slots to the thread that is still setting the copies up
  while (!mbar_try_wait(bar, parity))
    __nanosleep(32);
}
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream still runs; griddep_wait() blocks until the predecessor has completed
// and its writes are visible (no-op for a normal launch), griddep_launch_dependents() lets the successor's blocks be
// scheduled as soon as this grid's blocks leave room.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
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
// Packed FP32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): two independent force evaluations
// ride in the two halves of a 64-bit register pair.  Measured on B200 (scripts/dbg/ffma2_bench.cu):
// FFMA2 sustains the same 128 FMA/clk/SM as FFMA with HALF the issue slots, which is what this
// issue-bound kernel needs.  MUFU, min/max and selects stay scalar per half.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 mk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 bc2(float a) { return mk2(a, a); }
__device__ __forceinline__ void un2(f2 a, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2 rsqrt2(f2 a) {
  float lo, hi;
  un2(a, lo, hi);
  return mk2(rsqrt_approx(lo), rsqrt_approx(hi));
}

// ------------------------------------------------------------------------------------------------
// lightsfm pair social force in FP32.  Force on agent a from agent b, already
// scaled by forceFactorSocial.
//
// theta (angle from interactionDirection i to diffDirection e) is needed only as theta^2 plus its
// sign.  With I = lambda*vd + e (unnormalised), L = |I|:  I x e = lambda * (vd x e)  exactly, so the
// sine is formed from the velocity difference — no cancellation, and EXACTLY zero when the two
// velocities are equal, where lightsfm gets theta == 0 and switches the angular term off.
// |theta| = asin(min(|sin|,|cos|)) folded back by quadrant; asin on [0, 1/sqrt 2] is a 7-term odd
// minimax polynomial (|err| < 1.3e-7 in FP32), so no division and one MUFU less than atan2.
// ------------------------------------------------------------------------------------------------
struct SfmConst {
  float lambda, c_d, g2, c_np, c_n, k_soc; // g2 = gamma^2
  float kq_v, kq_a;                        // c_np * g2, c_n * g2: exponent = -c_d |d|/L - kq (L theta)^2
};
__device__ __forceinline__ SfmConst make_sfm_const(const SfwBatchDev &B) {
  SfmConst K;
  K.lambda = B.lambda;
  K.c_d = B.c_d;
  K.g2 = B.gamma * B.gamma;
  K.c_np = B.c_np;
  K.c_n = B.c_n;
  K.k_soc = B.k_soc;
  K.kq_v = B.c_np * K.g2;
  K.kq_a = B.c_n * K.g2;
  return K;
}

// asin(m)/m on m^2 in [0, 1/2]: 7-coefficient minimax (|err| of asin < 8e-8 exact, 1.3e-7 in FP32 Horner)
#define SFW_ASIN_C0 1.0000001192092896f
#define SFW_ASIN_C1 0.1666467934846878f
#define SFW_ASIN_C2 0.0755898728966713f
#define SFW_ASIN_C3 0.03815845027565956f
#define SFW_ASIN_C4 0.06345300376415253f
#define SFW_ASIN_C5 -0.05950368940830231f
#define SFW_ASIN_C6 0.10390560328960419f

// quadrant fold: phi = asin(min(|s|,|c|)) in [0, pi/4] -> |theta| in [0, pi]
__device__ __forceinline__ float fold_theta(float phi, float asn, float acs, float cs) {
  float th = (asn > acs) ? (1.5707963267948966f - phi) : phi;
  return (cs < 0.0f) ? (3.14159265358979f - th) : th;
}
// +sign(theta) * |mag| with lightsfm's Angle::sign(): 0 only for theta == 0, +1 for theta == pi
// (mag may carry any sign; only its magnitude is used)
__device__ __forceinline__ float signed_angle_term(float mag, float sn, float cs) {
  float fa = __uint_as_float((__float_as_uint(mag) & 0x7fffffffu) | (__float_as_uint(sn) & 0x80000000u));
  if (sn == 0.0f)
    fa = (cs < 0.0f) ? fabsf(mag) : 0.0f;
  return fa;
}
__device__ __forceinline__ float flip_sign(float v) { return __uint_as_float(__float_as_uint(v) ^ 0x80000000u); }

// scalar version (diagonal pairs, last-step pass)
template <bool WITH_MAG>
__device__ __forceinline__ void pair_force(const SfmConst &K, float ax, float ay, float avx,
                                           float avy, float bx, float by, float bvx, float bvy,
                                           float &fx, float &fy, float &fmag) {
  const float dx = bx - ax, dy = by - ay;
  const float d2 = fmaf(dx, dx, fmaf(dy, dy, 1e-30f));    // +eps: coincident agents give 0, not NaN
  const float rd = rsqrt_approx(d2);
  const float ex = dx * rd, ey = dy * rd;                 // diffDirection
  const float vdx = avx - bvx, nvdy = bvy - avy;          // velocity difference (y negated)
  const float ix = fmaf(K.lambda, vdx, ex);               // interactionVector
  const float iy = fmaf(-K.lambda, nvdy, ey);
  const float L2 = fmaf(ix, ix, fmaf(iy, iy, 1e-30f));
  const float rL = rsqrt_approx(L2);                      // 1 / interactionLength
  const float sn = K.lambda * fmaf(vdx, ey, nvdy * ex);   // I x e = lambda (vd x e): exactly 0 for vd == 0
  const float cs = fmaf(ix, ex, iy * ey);                 // I . e
  const float asn = fabsf(sn), acs = fabsf(cs);
  const float m = fminf(asn, acs) * rL;
  const float u = m * m;
  float p = SFW_ASIN_C6;
  p = fmaf(p, u, SFW_ASIN_C5);
  p = fmaf(p, u, SFW_ASIN_C4);
  p = fmaf(p, u, SFW_ASIN_C3);
  p = fmaf(p, u, SFW_ASIN_C2);
  p = fmaf(p, u, SFW_ASIN_C1);
  p = fmaf(p, u, SFW_ASIN_C0);
  const float th = fold_theta(p * m, asn, acs, cs);
  const float q = L2 * (th * th);                         // (L theta)^2
  const float t = ((d2 * rd) * rL) * -K.c_d;              // -|diff| / B  (in log2 units)
  const float e_vel = ex2_approx(fmaf(-K.kq_v, q, t));    // exp(-d/B - (n' B theta)^2)
  const float e_ang = ex2_approx(fmaf(-K.kq_a, q, t));    // exp(-d/B - (n  B theta)^2)
  const float rLkn = rL * -K.k_soc;
  const float fvn = e_vel * rLkn;                         // force along interactionVector (negative)
  const float fa = signed_angle_term(e_ang * rLkn, sn, cs);
  // F = fvn * I - fa * leftNormal(I), leftNormal(I) = (-iy, ix)
  fx = fmaf(fa, iy, fvn * ix);
  fy = fmaf(-fa, ix, fvn * iy);
  if (WITH_MAG) {
    const float ea = (sn == 0.0f && cs >= 0.0f) ? 0.0f : e_ang;
    fmag = K.k_soc * sqrt_approx(fmaf(e_vel, e_vel, ea * ea));
  }
}

// packed version: two (a, b) evaluations at once, one per register half (41 packed FP32x2 operations,
// 8 MUFU; the rest is per-half compare / select work on the ALU pipe)
template <bool WITH_MAG>
__device__ __forceinline__ void pair_force2(const SfmConst &K, f2 ax, f2 ay, f2 avx, f2 avy, f2 bx,
                                            f2 by, f2 bvx, f2 bvy, f2 &fx, f2 &fy, f2 &fmag) {
  const f2 eps = bc2(1e-30f);
  const f2 dx = sub2(bx, ax), dy = sub2(by, ay);
  const f2 d2 = fma2(dx, dx, fma2(dy, dy, eps));
  const f2 rd = rsqrt2(d2);
  const f2 ex = mul2(dx, rd), ey = mul2(dy, rd);
  const f2 vdx = sub2(avx, bvx), nvdy = sub2(bvy, avy);
  const f2 lam = bc2(K.lambda);
  const f2 ix = fma2(lam, vdx, ex), iy = fma2(bc2(-K.lambda), nvdy, ey);
  const f2 L2 = fma2(ix, ix, fma2(iy, iy, eps));
  const f2 rL = rsqrt2(L2);
  const f2 sn = mul2(lam, fma2(vdx, ey, mul2(nvdy, ex)));
  const f2 cs = fma2(ix, ex, mul2(iy, ey));
  float sn0, sn1, cs0, cs1;
  un2(sn, sn0, sn1);
  un2(cs, cs0, cs1);
  const float asn0 = fabsf(sn0), asn1 = fabsf(sn1), acs0 = fabsf(cs0), acs1 = fabsf(cs1);
  const f2 m = mul2(mk2(fminf(asn0, acs0), fminf(asn1, acs1)), rL);
  const f2 u = mul2(m, m);
  f2 p = bc2(SFW_ASIN_C6);
  p = fma2(p, u, bc2(SFW_ASIN_C5));
  p = fma2(p, u, bc2(SFW_ASIN_C4));
  p = fma2(p, u, bc2(SFW_ASIN_C3));
  p = fma2(p, u, bc2(SFW_ASIN_C2));
  p = fma2(p, u, bc2(SFW_ASIN_C1));
  p = fma2(p, u, bc2(SFW_ASIN_C0));
  float ph0, ph1;
  un2(mul2(p, m), ph0, ph1);
  const f2 th = mk2(fold_theta(ph0, asn0, acs0, cs0), fold_theta(ph1, asn1, acs1, cs1));
  const f2 q = mul2(L2, mul2(th, th));
  const f2 t = mul2(mul2(mul2(d2, rd), rL), bc2(-K.c_d));
  float a0, a1, b0, b1;
  un2(fma2(bc2(-K.kq_v), q, t), a0, a1);
  un2(fma2(bc2(-K.kq_a), q, t), b0, b1);
  const float ev0 = ex2_approx(a0), ev1 = ex2_approx(a1);
  const float ea0 = ex2_approx(b0), ea1 = ex2_approx(b1);
  const f2 rLkn = mul2(rL, bc2(-K.k_soc));
  const f2 fvn = mul2(mk2(ev0, ev1), rLkn);
  float m0, m1;
  un2(mul2(mk2(ea0, ea1), rLkn), m0, m1);
  const float fa0 = signed_angle_term(m0, sn0, cs0), fa1 = signed_angle_term(m1, sn1, cs1);
  fx = fma2(mk2(fa0, fa1), iy, mul2(fvn, ix));
  fy = fma2(mk2(flip_sign(fa0), flip_sign(fa1)), ix, mul2(fvn, iy));
  if (WITH_MAG) {
    const float z0 = (sn0 == 0.0f && cs0 >= 0.0f) ? 0.0f : ea0;
    const float z1 = (sn1 == 0.0f && cs1 >= 0.0f) ? 0.0f : ea1;
    fmag = mk2(K.k_soc * sqrt_approx(fmaf(ev0, ev0, z0 * z0)), K.k_soc * sqrt_approx(fmaf(ev1, ev1, z1 * z1)));
  }
}

// lightsfm obstacle force sum (unscaled): sum_o exp(-|p-o|/sigma) (p-o)/|p-o|.  Obstacle points are
// stored pre-multiplied by c_obs = log2(e)/sigma, and so is the query point: then |p'-o'| is the
// exponent in log2 units and the unit vector is unchanged.  The list comes as clusters of 8 points behind
// a 2-slot header {centre, reach^2} (sfw_dev.h); the last cluster is padded with points 1e15 away, whose
// term is exactly 0 (ex2 underflows).  `M` counts float2 SLOTS (10 per cluster).
// one obstacle point against the two pedestrians of a pair
__device__ __forceinline__ void obstacle_term2(float2 p, f2 qx, f2 qy, f2 &ax, f2 &ay) {
  const f2 dx = sub2(qx, bc2(p.x)), dy = sub2(qy, bc2(p.y));
  const f2 d2 = fma2(dx, dx, fma2(dy, dy, bc2(1e-30f)));
  const f2 rd = rsqrt2(d2);
  float d0, d1;
  un2(mul2(d2, rd), d0, d1);
  const f2 e = mul2(mk2(ex2_approx(-d0), ex2_approx(-d1)), rd);
  ax = fma2(e, dx, ax);
  ay = fma2(e, dy, ay);
}

// (a) two query points (a pedestrian pair) against every obstacle, one cluster of 8 points per trip (the
// loop is instruction bound: a software exponential on the FMA pipe for some of the points made it slower,
// DESIGN.md 4.4).  A cluster out of reach of BOTH pedestrians is skipped.  The test only reads the thread's
// own pair, so the result does not depend on which trajectories share a warp (prefix sharing and the
// warp-per-path writers stay bit-identical); pedestrians are within millimetres of each other across the
// trajectories of a warp, so the branch is uniform in practice.
// `first` / `stride` (in clusters): the block-per-trajectory kernel deals the clusters of a pair over helper threads.
__device__ __forceinline__ void obstacle_sum2(const float2 *__restrict__ obs, int M, float c_obs, f2 px,
                                              f2 py, f2 &sx, f2 &sy, int first = 0, int stride = 1) {
  f2 ax = bc2(0.f), ay = bc2(0.f);
  const f2 qx = mul2(px, bc2(c_obs)), qy = mul2(py, bc2(c_obs));
  for (int o = first * SFW_OBST_CLUSTER_SLOTS; o < M; o += stride * SFW_OBST_CLUSTER_SLOTS) {
    const float4 hd = *reinterpret_cast<const float4 *>(obs + o);
    const f2 cx = sub2(qx, bc2(hd.x)), cy = sub2(qy, bc2(hd.y));
    float c0, c1;
    un2(fma2(cx, cx, mul2(cy, cy)), c0, c1);
    if (fminf(c0, c1) <= hd.z) {
#pragma unroll
      for (int i = 0; i < SFW_OBST_CLUSTER; ++i)
        obstacle_term2(obs[o + 2 + i], qx, qy, ax, ay);
    }
  }
  sx = ax;
  sy = ay;
}
// (b) one query point (the robot) against two obstacles per iteration; never skips a cluster (the robot's
// obstacle force goes into the social work as it is)
__device__ __forceinline__ void obstacle_sum1(const float2 *__restrict__ obs, int M, float c_obs,
                                              float px, float py, float &sx, float &sy) {
  f2 ax = bc2(0.f), ay = bc2(0.f);
  const f2 eps = bc2(1e-30f);
  const f2 qx = bc2(px * c_obs), qy = bc2(py * c_obs);
  for (int o = 0; o < M; o += SFW_OBST_CLUSTER_SLOTS) {
    const float4 *__restrict__ obs4 = reinterpret_cast<const float4 *>(obs + o + 2);
#pragma unroll
    for (int i = 0; i < SFW_OBST_CLUSTER / 2; ++i) {
      const float4 p = obs4[i];
      const f2 dx = sub2(qx, mk2(p.x, p.z)), dy = sub2(qy, mk2(p.y, p.w));
      const f2 d2 = fma2(dx, dx, fma2(dy, dy, eps));
      const f2 rd = rsqrt2(d2);
      float d0, d1;
      un2(mul2(d2, rd), d0, d1);
      const f2 e = mul2(mk2(ex2_approx(-d0), ex2_approx(-d1)), rd);
      ax = fma2(e, dx, ax);
      ay = fma2(e, dy, ay);
    }
  }
  float x0, x1, y0, y1;
  un2(ax, x0, x1);
  un2(ay, y0, y1);
  sx = x0 + x1;
  sy = y0 + y1;
}

// ------------------------------------------------------------------------------------------------
// lightsfm group forces (lightsfm computeGroupForce, reached through the
// computeForces call at reference src/sfw_planner.cpp:592) for ONE member of one group, FP32.
//   gaze:       k_gaze * (dd . rel) * dd  when the centre of mass of the OTHER members lies behind the
//               desired direction dd (angle > 90 deg <=> dd . rel < 0; dd == 0 without a live goal -> none)
//   coherence:  k_coh * (tanh(dist - (c-1)/2) + 1)/2 * (centre - p) = k_coh * sigmoid(2 x) * (centre - p)
//   repulsion:  k_rep * sum over members closer than r_a + r_b of (p_a - p_b)
// Positions are read through (pair k, half h): P[k * stride + h] = x, P[k * stride + 2 + h] = y with
// `stride` floats between consecutive pairs (4 for a plain float4 array, 4*T for the per-thread columns
// of the thread-per-trajectory kernel).  `mem` points at the group's member records (ped index, radius).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_member_force(const uint32_t *__restrict__ mem, uint32_t c, uint32_t self,
                                                   const float *P, uint32_t stride, float cx, float cy,
                                                   float ddx, float ddy, float k_gaze, float k_coh, float k_rep,
                                                   float &fx, float &fy) {
  const uint32_t j = mem[2u * self];
  const float rad = __uint_as_float(mem[2u * self + 1u]);
  const uint32_t k = j >> 1, h = j & 1u;
  const float px = P[k * stride + h], py = P[k * stride + 2u + h];
  const float fc = (float)c;
  // gaze: centre of mass of the others = (c * centre - p) / (c - 1)
  const float inv_cm1 = 1.0f / (fc - 1.0f);
  const float relx = (fc * cx - px) * inv_cm1 - px, rely = (fc * cy - py) * inv_cm1 - py;
  const float ep = fmaf(ddx, relx, ddy * rely);
  const float gz = (ep < 0.0f) ? k_gaze * ep : 0.0f;
  fx = gz * ddx;
  fy = gz * ddy;
  // coherence
  const float rx = cx - px, ry = cy - py;
  const float d2 = fmaf(rx, rx, ry * ry);
  const float dist = d2 * rsqrt_approx(fmaxf(d2, 1e-30f));
  const float x = dist - 0.5f * (fc - 1.0f);
  const float soft = k_coh * rcp_approx(1.0f + ex2_approx(-2.885390081777927f * x)); // sigmoid(2x)
  fx = fmaf(rx, soft, fx);
  fy = fmaf(ry, soft, fy);
  // repulsion
  float rpx = 0.f, rpy = 0.f;
  for (uint32_t m = 0; m < c; ++m) {
    if (m == self)
      continue;
    const uint32_t j2 = mem[2u * m];
    const float rr = rad + __uint_as_float(mem[2u * m + 1u]);
    const uint32_t k2 = j2 >> 1, h2 = j2 & 1u;
    const float dx = px - P[k2 * stride + h2], dy = py - P[k2 * stride + 2u + h2];
    if (fmaf(dx, dx, dy * dy) < rr * rr) {
      rpx += dx;
      rpy += dy;
    }
  }
  fx = fmaf(k_rep, rpx, fx);
  fy = fmaf(k_rep, rpy, fy);
}

// centre of a group (mean of the member positions)
__device__ __forceinline__ void group_centre(const uint32_t *__restrict__ mem, uint32_t c, const float *P,
                                             uint32_t stride, float &cx, float &cy) {
  float sx = 0.f, sy = 0.f;
  for (uint32_t m = 0; m < c; ++m) {
    const uint32_t j = mem[2u * m];
    const uint32_t k = j >> 1, h = j & 1u;
    sx += P[k * stride + h];
    sy += P[k * stride + 2u + h];
  }
  const float inv = 1.0f / (float)c;
  cx = sx * inv;
  cy = sy * inv;
}

// desired direction of pedestrian j (computeDesiredForce's return value): unit vector to a live goal, else 0
__device__ __forceinline__ void desired_direction(const float *P, uint32_t stride, const float4 *__restrict__ goal,
                                                  const float4 *__restrict__ par, uint32_t j, bool has_goal,
                                                  float &ddx, float &ddy) {
  const uint32_t k = j >> 1, h = j & 1u;
  const float *G = reinterpret_cast<const float *>(goal + k);
  const float *R = reinterpret_cast<const float *>(par + k);
  const float gx = G[h] - P[k * stride + h], gy = G[2u + h] - P[k * stride + 2u + h];
  const float g2 = fmaf(gx, gx, gy * gy);
  const bool on = has_goal && g2 > R[h];
  const float rg = on ? rsqrt_approx(g2) : 0.0f;
  ddx = gx * rg;
  ddy = gy * rg;
}

// ------------------------------------------------------------------------------------------------
// Costmap access + footprint rasterisation (bit-faithful integer work)
// ------------------------------------------------------------------------------------------------
struct MapView {
  const uint8_t *win;   // staged window in shared memory (or nullptr)
  const uint8_t *glob;  // this scene's costmap slot in HBM
  double ox, oy, res, rinv; // rinv = fl(1 / res)
  uint32_t sx, sy, pitch;
  int32_t wx0, wy0;
  uint32_t wwp, wh;     // window pitch / rows
  const uint32_t *free_bits; // [wh][fw32] bit = whole (2 fp_rc + 1)^2 neighbourhood is in the map and free (or nullptr)
  uint32_t fw32;
};

__device__ __forceinline__ uint32_t cell_cost(const MapView &m, int cx, int cy) {
  const uint32_t lx = (uint32_t)(cx - m.wx0), ly = (uint32_t)(cy - m.wy0);
  if (lx < m.wwp && ly < m.wh)
    return m.win[ly * m.wwp + lx];
  return __ldg(m.glob + (size_t)cy * m.pitch + cx);
}

// trunc(fl(a / res)) exactly as the reference's division produces it, without dividing:
// q = a * fl(1/res) is within |q| * 2^-51 of fl(a / res); unless q sits that close to an integer
// the truncations agree.  The (measure ~2^-19) near-integer case and huge quotients take the
// correctly rounded division.
__device__ __forceinline__ unsigned int cell_index(double a, double res, double rinv) {
  const double q = __dmul_rn(a, rinv);
  if (q < 1073741824.0) {
    const int i = __double2int_rz(q);
    const double f = __dsub_rn(q, (double)i);
    if (f > 9.5367431640625e-07 && f < 0.99999904632568359375)
      return (unsigned int)i;
  }
  const double e = __ddiv_rn(a, res);
  return e < 4294967296.0 ? (unsigned int)e : 0xffffffffu;
}

// nav2 Costmap2D::worldToMap [external, SURVEY.md App. C]; no FMA contraction
__device__ __forceinline__ bool world_to_map(const MapView &m, double wx, double wy, int &mx, int &my) {
  if (wx < m.ox || wy < m.oy)
    return false;
  const unsigned int ux = cell_index(__dsub_rn(wx, m.ox), m.res, m.rinv);
  const unsigned int uy = cell_index(__dsub_rn(wy, m.oy), m.res, m.rinv);
  mx = (int)ux;
  my = (int)uy;
  return ux < m.sx && uy < m.sy;
}

// CostmapModel::lineCost over LineIterator (costmap_model.cpp:95-110, line_iterator.hpp:37-97):
// the max cell cost along the Bresenham line.  The reference stops at the first 254/255 cell; those
// are the two largest values, so "max >= 254" carries the same information without a per-cell branch.
// IN_WINDOW: both end cells (hence the whole line) lie in the staged shared-memory window.
template <bool IN_WINDOW>
__device__ __forceinline__ int line_max(const MapView &m, int x0, int y0, int x1, int y1) {
  const int dx = x1 - x0, dy = y1 - y0;
  const int adx = abs(dx), ady = abs(dy);
  const int sx = (dx >= 0) ? 1 : -1, sy = (dy >= 0) ? 1 : -1; // line_iterator.hpp:43-61
  const bool xmajor = adx >= ady;                              // :63
  const int den = xmajor ? adx : ady, numadd = xmajor ? ady : adx;
  int num = den >> 1, worst = 0;
  if (IN_WINDOW) {
    const int pitch = (int)m.wwp;
    const int step_always = xmajor ? sx : sy * pitch; // (xinc2, yinc2)
    const int step_carry = xmajor ? sy * pitch : sx;  // (xinc1, yinc1)
    int idx = (y0 - m.wy0) * pitch + (x0 - m.wx0);
    for (int cur = 0; cur <= den; ++cur) {
      worst = max(worst, (int)m.win[idx]);
      num += numadd;
      if (num >= den) {
        num -= den;
        idx += step_carry;
      }
      idx += step_always;
    }
  } else {
    int x = x0, y = y0;
    for (int cur = 0; cur <= den; ++cur) {
      worst = max(worst, (int)cell_cost(m, x, y));
      num += numadd;
      if (num >= den) {
        num -= den;
        x += xmajor ? 0 : sx;
        y += xmajor ? sy : 0;
      }
      x += xmajor ? sx : 0;
      y += xmajor ? 0 : sy;
    }
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
  // Free-space shortcut (exact): every vertex cell and every Bresenham cell between two of them lies in
  // the square of +-fp_rc cells around the centre cell.  If that whole square is inside the map and
  // holds only cost 0, every vertex maps and the maximum over any subset of its cells is 0.
  if (m.free_bits) {
    const uint32_t lx = (uint32_t)(cx - m.wx0), ly = (uint32_t)(cy - m.wy0);
    if (lx < m.wwp && ly < m.wh && ((m.free_bits[ly * m.fw32 + (lx >> 5)] >> (lx & 31u)) & 1u))
      return 0;
  }
  int worst = 0;
  int fx0 = 0, fy0 = 0, px = 0, py = 0;
  bool pin = false, fin = false;
  for (int i = 0; i <= F; ++i) {
    int mx, my;
    bool in;
    if (i < F) {
      const double2 v = fp[i];
      // world_model.hpp:56-59, evaluated without contraction
      const double wx = __dadd_rn(x, __dsub_rn(__dmul_rn(v.x, cs), __dmul_rn(v.y, sn)));
      const double wy = __dadd_rn(y, __dadd_rn(__dmul_rn(v.x, sn), __dmul_rn(v.y, cs)));
      if (!world_to_map(m, wx, wy, mx, my))
        return -1;
      in = (uint32_t)(mx - m.wx0) < m.wwp && (uint32_t)(my - m.wy0) < m.wh;
    } else { // closing edge last -> first
      mx = fx0;
      my = fy0;
      in = fin;
    }
    if (i == 0) {
      fx0 = mx;
      fy0 = my;
      fin = in;
    } else {
      const int lc = (pin && in) ? line_max<true>(m, px, py, mx, my) : line_max<false>(m, px, py, mx, my);
      worst = max(worst, lc);
    }
    px = mx;
    py = my;
    pin = in;
  }
  return (worst >= 254) ? -1 : worst;
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

// Fused winner exchange: store one scene's record into every rank's gather buffer (peer memory over NVLink),
// make the stores visible system-wide, then signal each rank's arrival counter.  Called by ONE thread.
__device__ __forceinline__ void export_best(const SfwExchangeDev &X, uint32_t scene, const SfwBest &r) {
  const int4 lo = reinterpret_cast<const int4 *>(&r)[0], hi = reinterpret_cast<const int4 *>(&r)[1];
  const size_t at = ((size_t)X.slot * X.world + X.rank) * X.max_scenes + scene;
  for (uint32_t q = 0; q < X.world; ++q) {
    int4 *dst = reinterpret_cast<int4 *>(X.peer_best[q] + at);
    dst[0] = lo;
    dst[1] = hi;
  }
  __threadfence_system();
  for (uint32_t q = 0; q < X.world; ++q)
    atomicAdd_system(X.peer_arrived[q] + X.rank, 1u);
}

// A cost can only become "best" if 0 <= cost <= 10000; == 10000 needs linvel > 0 because the
// initial best is (10000, xv = 0, thetav = 0) (sfw_planner.cpp:338-344,394-407).
__device__ __forceinline__ bool eligible(float c, double linvel) {
  return c >= 0.f && (c < 10000.f || (c == 10000.f && linvel > 0.0));
}

} // namespace
#endif
