// sfw_sensor.cu — the step before the scoring path: laser scan -> obstacle points.
//
// Device restatement of SFMSensorInterface::laserCb (reference src/sensor_interface.cpp:103-229):
//   beam filter (:120-122), polar -> cartesian in float with the scan angle ACCUMULATED in float
//   (:118,124-127), optional planar transform into the controller frame (:143-170), removal of the points
//   that lie within person_radius of a detected person (float hypot, :210-225), order preserved.
// One block per scan.  The float angle recurrence is a serial chain (float addition does not
// associate), so one thread replays it into shared memory; everything else is beam-parallel, and the
// survivors are compacted in beam order with a ballot / block scan.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "sfw_ctx.h"

namespace {

constexpr int kThreads = 256;

struct ScanDev {
  uint32_t range_off, n_ranges; // into the packed ranges array
  uint32_t people_off, n_people; // into the packed people array (pairs)
  float angle_min, angle_inc;
  int32_t has_tf, pad;
  double tx, ty, cs, sn; // cos / sin of tf_yaw evaluated on the host in double
};

__global__ void __launch_bounds__(kThreads)
sfw_laser_kernel(const ScanDev *__restrict__ scans, const float *__restrict__ ranges,
                 const double2 *__restrict__ people, float max_dist, float person_radius,
                 double2 *__restrict__ out, uint32_t slot, uint32_t *__restrict__ n_out,
                 double2 *__restrict__ compact, uint32_t *__restrict__ compact_off,
                 unsigned int *__restrict__ compact_fill) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ScanDev sc = scans[blockIdx.x];
  float *s_angle = reinterpret_cast<float *>(smem);                                      // [n_ranges]
  double2 *s_people = reinterpret_cast<double2 *>(smem + ((sc.n_ranges * 4u + 15u) & ~15u)); // [n_people]
  __shared__ uint32_t s_warp[kThreads / 32];
  __shared__ uint32_t s_base;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) {
    float a = sc.angle_min; // cpp:118; advanced after every beam, valid or not (:127)
    for (uint32_t i = 0; i < sc.n_ranges; ++i) {
      s_angle[i] = a;
      a = __fadd_rn(a, sc.angle_inc);
    }
    s_base = 0;
  }
  for (uint32_t i = tid; i < sc.n_people; i += kThreads)
    s_people[i] = people[sc.people_off + i];
  __syncthreads();
  double2 *dst = out + (size_t)blockIdx.x * slot;
  for (uint32_t i0 = 0; i0 < sc.n_ranges; i0 += kThreads) {
    const uint32_t i = i0 + tid;
    bool keep = false;
    double px = 0.0, py = 0.0;
    if (i < sc.n_ranges) {
      const float r = ranges[sc.range_off + i];
      if (!isnan(r) && isfinite(r) && r < max_dist) { // cpp:120-122
        // ranges[i] * cos(angle) with float operands (math.h's float overloads, hpp:52): the correctly
        // rounded float cosine, taken from the double routine
        const float a = s_angle[i];
        const float cf = (float)cos((double)a), sf = (float)sin((double)a);
        px = (double)__fmul_rn(r, cf);
        py = (double)__fmul_rn(r, sf);
        if (sc.has_tf) { // cpp:143-170 (planar rigid transform)
          const double qx = __dadd_rn(__dsub_rn(__dmul_rn(sc.cs, px), __dmul_rn(sc.sn, py)), sc.tx);
          const double qy = __dadd_rn(__dadd_rn(__dmul_rn(sc.sn, px), __dmul_rn(sc.cs, py)), sc.ty);
          px = qx;
          py = qy;
        }
        keep = true;
        for (uint32_t q = 0; q < sc.n_people; ++q) { // cpp:210-225
          const float dx = (float)__dsub_rn(px, s_people[q].x), dy = (float)__dsub_rn(py, s_people[q].y);
          // glibc hypotf: sqrt of the exact double sum of squares, rounded once
          const float d = (float)sqrt(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)));
          if (d <= person_radius) {
            keep = false;
            break;
          }
        }
      }
    }
    __syncwarp();
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0)
      s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = s_base;
    for (uint32_t w = 0; w < warp; ++w)
      before += s_warp[w];
    if (keep)
      dst[before + __popc(bal & ((1u << lane) - 1u))] = make_double2(px, py);
    __syncthreads();
    if (tid == 0) {
      uint32_t t = 0;
      for (int w = 0; w < kThreads / 32; ++w)
        t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  // Second copy of the survivors into one dense buffer (a range reserved with one atomic per scan), so the
  // host fetches sum(kept) points with a single D2H instead of one copy per scan.
  const uint32_t kept = s_base;
  if (tid == 0) {
    n_out[blockIdx.x] = kept;
    s_warp[0] = kept ? atomicAdd(compact_fill, kept) : 0u;
    compact_off[blockIdx.x] = s_warp[0];
  }
  __syncthreads();
  const uint32_t base = s_warp[0];
  for (uint32_t i = tid; i < kept; i += kThreads)
    compact[base + i] = dst[i];
}

} // namespace

extern "C" int sfw_laser_obstacles(sfw_ctx *c, const SfwLaserScan *scans, uint32_t n_scans, float max_obstacle_dist,
                                   float person_radius, double *points_xy_out, uint32_t max_points_per_scan,
                                   uint32_t *n_points_out) {
  if (!c)
    return SFW_ERR_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!scans || !n_scans || !n_points_out || (!points_xy_out && max_points_per_scan))
    return sfw_fail(c, SFW_ERR_ARG, "sfw_laser_obstacles: null/empty argument");
  uint64_t tot_r = 0, tot_p = 0;
  uint32_t max_r = 0, max_p = 0;
  for (uint32_t s = 0; s < n_scans; ++s) {
    if ((scans[s].n_ranges && !scans[s].ranges) || (scans[s].n_people && !scans[s].people_xy))
      return sfw_fail(c, SFW_ERR_ARG, "scan %u: null array with non-zero count", s);
    if (scans[s].n_ranges > max_points_per_scan)
      return sfw_fail(c, SFW_ERR_ARG, "scan %u: %u beams but only %u output points per scan", s, scans[s].n_ranges,
                      max_points_per_scan);
    tot_r += scans[s].n_ranges;
    tot_p += scans[s].n_people;
    max_r = std::max(max_r, scans[s].n_ranges);
    max_p = std::max(max_p, scans[s].n_people);
  }
  const size_t smem = ((size_t)max_r * 4 + 15) / 16 * 16 + (size_t)max_p * 16;
  if (smem > 200 * 1024)
    return sfw_fail(c, SFW_ERR_UNSUPPORTED, "scan too large for one block (%u beams, %u people)", max_r, max_p);
  SFW_CK(c, cudaSetDevice(c->device));
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_scan = 0, o_rng = up(sizeof(ScanDev) * n_scans), o_ppl = o_rng + up(4 * tot_r);
  const size_t in_bytes = o_ppl + up(16 * tot_p);
  const size_t slot = max_points_per_scan;
  // out arena: counts | dense offsets | fill counter | dense points | per-scan slots
  const size_t o_cnt = 0, o_off = up(4 * (size_t)n_scans), o_fill = o_off + up(4 * (size_t)n_scans);
  const size_t o_dense = o_fill + 256, o_pts = o_dense + up(16 * tot_r), out_bytes = o_pts + 16 * slot * n_scans;
  int rc = sfw_arena_reserve(c, c->sensor_in, in_bytes);
  if (rc != SFW_OK)
    return rc;
  rc = sfw_arena_reserve(c, c->sensor_out, out_bytes);
  if (rc != SFW_OK)
    return rc;
  uint8_t *h = c->sensor_in.host;
  ScanDev *hs = reinterpret_cast<ScanDev *>(h + o_scan);
  float *hr = reinterpret_cast<float *>(h + o_rng);
  double *hp = reinterpret_cast<double *>(h + o_ppl);
  uint32_t pr = 0, pp = 0;
  for (uint32_t s = 0; s < n_scans; ++s) {
    const SfwLaserScan &L = scans[s];
    ScanDev &d = hs[s];
    d.range_off = pr;
    d.n_ranges = L.n_ranges;
    d.people_off = pp;
    d.n_people = L.n_people;
    d.angle_min = L.angle_min;
    d.angle_inc = L.angle_increment;
    d.has_tf = L.has_tf;
    d.pad = 0;
    d.tx = L.tf_x;
    d.ty = L.tf_y;
    d.cs = std::cos(L.tf_yaw);
    d.sn = std::sin(L.tf_yaw);
    if (L.n_ranges)
      memcpy(hr + pr, L.ranges, 4 * (size_t)L.n_ranges);
    if (L.n_people)
      memcpy(hp + 2 * (size_t)pp, L.people_xy, 16 * (size_t)L.n_people);
    pr += L.n_ranges;
    pp += L.n_people;
  }
  SFW_CK(c, cudaMemcpyAsync(c->sensor_in.dev, h, in_bytes, cudaMemcpyHostToDevice, c->stream));
  if (!c->laser_attr_set) { // per context: a process may drive several devices
    SFW_CK(c, cudaFuncSetAttribute(sfw_laser_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    c->laser_attr_set = true;
  }
  uint8_t *dv = c->sensor_in.dev, *dout = c->sensor_out.dev;
  SFW_CK(c, cudaMemsetAsync(dout + o_fill, 0, 4, c->stream));
  sfw_laser_kernel<<<n_scans, kThreads, smem, c->stream>>>(
      reinterpret_cast<const ScanDev *>(dv + o_scan), reinterpret_cast<const float *>(dv + o_rng),
      reinterpret_cast<const double2 *>(dv + o_ppl), max_obstacle_dist, person_radius,
      reinterpret_cast<double2 *>(dout + o_pts), (uint32_t)slot, reinterpret_cast<uint32_t *>(dout + o_cnt),
      reinterpret_cast<double2 *>(dout + o_dense), reinterpret_cast<uint32_t *>(dout + o_off),
      reinterpret_cast<unsigned int *>(dout + o_fill));
  SFW_CK(c, cudaGetLastError());
  c->launches += 1;
  c->last_kernel = "sfw_laser_kernel";
  // counts + dense offsets + total first, then the dense points in one copy
  SFW_CK(c, cudaMemcpyAsync(c->sensor_out.host + o_cnt, dout + o_cnt, o_dense, cudaMemcpyDeviceToHost, c->stream));
  SFW_CK(c, cudaStreamSynchronize(c->stream));
  const uint32_t *cnt = reinterpret_cast<const uint32_t *>(c->sensor_out.host + o_cnt);
  const uint32_t *offs = reinterpret_cast<const uint32_t *>(c->sensor_out.host + o_off);
  const uint32_t total = *reinterpret_cast<const uint32_t *>(c->sensor_out.host + o_fill);
  if (total) {
    SFW_CK(c, cudaMemcpyAsync(c->sensor_out.host + o_dense, dout + o_dense, 16 * (size_t)total, cudaMemcpyDeviceToHost,
                              c->stream));
    SFW_CK(c, cudaStreamSynchronize(c->stream));
  }
  for (uint32_t s = 0; s < n_scans; ++s) {
    n_points_out[s] = cnt[s];
    if (cnt[s])
      memcpy(points_xy_out + 2 * slot * s, c->sensor_out.host + o_dense + 16 * (size_t)offs[s], 16 * (size_t)cnt[s]);
  }
  return SFW_OK;
}
