// sfw_dev.h — device-side data layout shared by the host packer (sfw_abi.cu) and the kernels
// (sfw_kernels.cu).  Everything here is OUR layout in HBM (DESIGN.md "Data layout"); the caller
// facing types live in include/sfw_b200.h.
//
// Frame convention: pedestrians, obstacle points and the robot's SFM position are stored in FP32
// relative to the rollout start pose (robot.x, robot.y) of their scene, computed in FP64 on the
// host before narrowing.  The social-force model only ever uses position differences, so the
// translation is exact in the model and keeps FP32 resolution (~1e-7 m) independent of where the
// odom frame's origin is.  The robot rollout itself and every costmap cell index stay in FP64
// world coordinates so that cell choice is bit-identical to the reference.
#ifndef SFW_DEV_H
#define SFW_DEV_H

#include <stdint.h>

#include "../../include/sfw_b200.h"

#define SFW_MAX_PEDS_SMALL 64 /* thread-per-trajectory kernel: goal flags live in one 64-bit mask */
#define SFW_FAR_AWAY 1.0e15f  /* padding pedestrian / obstacle: every force term underflows to exactly 0 */
#define SFW_MAX_FOOTPRINT 64
#define SFW_CROWD_THREADS 256       /* block-per-trajectory kernel (sfw_crowd.cu) */
#define SFW_CROWD_THREADS_SMALL 128 /* its second instantiation: small crowds, four blocks per SM */
/* Block-per-trajectory kernel: the whole crowd of one trajectory lives in one block's shared memory, 208 B per
 * pedestrian PAIR (position, velocity, goal, 2 parameter words, 8 per-warp reaction rows) + the rollout arrays:
 * about 2 100 pedestrians at 128 steps on a B200 (227 KB per block).  2048 is the guaranteed limit; between 2048
 * and the shared-memory ceiling sfw_upload still answers SFW_ERR_UNSUPPORTED with the byte count. */
#define SFW_MAX_PEDS_CROWD 2048
#define SFW_MAX_BLOCK_SMALL 512 /* launch bound of the thread-per-trajectory kernel (128 regs/thread) */
#define SFW_PATH_WARP_THREADS 128 /* block of the warp-per-path record writer: 4 paths */

// Obstacle points are stored as spatially compact CLUSTERS of 8 (sfw_abi.cu: pack_obstacles): 10 float2 slots
// per cluster = {centre x, centre y}, {reach^2, unused}, then the 8 points (the last cluster padded with
// SFW_FAR_AWAY points).  All in the obstacle_sum2 units (metres * log2(e)/sigma).  A pedestrian whose squared
// distance to the centre exceeds reach^2 = (cluster radius + largest pedestrian radius + cutoff)^2 skips the
// cluster: every term of it is below 2^-cutoff of the force factor.
#define SFW_OBST_CLUSTER 8
#define SFW_OBST_CLUSTER_SLOTS 10
#define SFW_OBST_CUTOFF_LOG2 24.0 /* default cutoff: terms below 2^-24 (one FP32 ulp of the factor) */
static inline uint32_t sfw_obst_slots(uint32_t n_points) {
  return SFW_OBST_CLUSTER_SLOTS * ((n_points + SFW_OBST_CLUSTER - 1u) / SFW_OBST_CLUSTER);
}

struct SfwSceneDev {
  // robot rollout start (FP64 world frame) — SfwRobot
  double rx, ry, rth, rvx, rvy, rvth;
  double wpx, wpy;
  // costmap geometry
  double origin_x, origin_y, resolution;
  // agents[0] as the sensor interface saw it, scene frame FP32
  float ax, ay, avx, avy;
  float a_obs_scale; // (k_obs / M) * exp(agent_radius / sigma)
  uint32_t size_x, size_y;
  int32_t win_x0, win_y0; // first cell of the staged window (may be negative / beyond the map)
  uint32_t n_peds, n_obst, n_fp; // n_obst: float2 SLOTS of the clustered obstacle list (sfw_obst_slots)
  uint32_t ped_off, obs_off, fp_off; // element offsets into the packed arrays (ped_off in PAIRS)
  uint32_t n_pairs;                  // ceil(n_peds / 2); an odd crowd is padded with a far-away agent
  uint64_t map_off; // byte offset of this scene's costmap slot
  uint64_t goal_mask; // bit j: pedestrian j has a goal (sfm::Agent::goals non-empty)
  // every cell the footprint can touch lies within fp_rc cells (Chebyshev) of the robot's own cell:
  // floor(circumscribed radius / resolution) + 2; 0 switches the free-space shortcut off
  uint32_t fp_rc;
  // pedestrian groups (lightsfm computeGroupForce): table at SfwBatchDev::groups + grp_off, laid out as
  // start[0..n_groups] (member index of each group's first member), then 2 words per member:
  // pedestrian index, radius (float bits).  Only groups with >= 2 members in the scene are listed.
  uint32_t n_groups, grp_off, pad1;
};

struct SfwBlockBest {
  float cost; // < 0 => no valid trajectory in the tile
  uint32_t index;
};

#define SFW_DEVSTAT_PATH_WAIT 1u /* merged path launch: a record this path continues from never appeared */
#define SFW_MAX_RANKS 8
// Multi-GPU winner exchange fused into the scorer's epilogue (csrc/sfw_exchange.cu): every rank owns a gather
// buffer [2 epochs][world][max_scenes] of SfwBest plus one arrival counter per source rank; the block that
// reduces a scene's winner stores the record straight into EVERY rank's buffer over NVLink (peer mappings of
// cudaIpc handles) and then bumps that rank's counter with a system-scope release.
struct SfwExchangeDev {
  SfwBest *peer_best[SFW_MAX_RANKS];         // peer q's gather buffer (q == rank: the local one)
  unsigned int *peer_arrived[SFW_MAX_RANKS]; // peer q's arrival counters [world]
  uint32_t rank, world, max_scenes, enabled;
  uint32_t slot, pad0;                       // epoch parity of this launch
};

// Rollout prefix sharing (thread-per-trajectory kernel, DESIGN.md 4.1 "prefix sharing").  The accel-limited
// unicycle makes many samples identical for their first steps: while a sample's linear (angular) velocity is
// still ramping at the full +-a*dt per step it is indistinguishable from every other sample that ramps the same
// way.  kv[r] / kw[c] = number of leading fully saturated updates of row r / column c, dirv / dirw = their
// direction.  Sample (r, c) is bit-identical to a SHARED path for its first max(kv, kw) steps:
//   steps <= min(kv, kw): one of the 4 doubly saturated paths              (records region A)
//   then, until max:      the path "v saturated, w as column c" (region B) or "v as row r, w saturated" (C)
// Launch 1 simulates A, launch 2 B and C (each starting from an A record), launch 3 the samples themselves,
// each from the record of its own fork point.  A record = SfwCkptHdr + (pos, vel) float4 per pedestrian pair.
struct SfwCkptHdr {
  double x, y, th, vx, vth, social_work, costmap_sum;
  float prx, pry, rvxf, rvyf;
  uint64_t goalmask;
  int32_t npts, alive;
  uint32_t epoch; // merged path launch (mode 4): the record is complete for run `epoch` (release / acquire)
  uint32_t pad;
};
struct SfwShareDev {
  uint8_t *records;         // [scene][4 + 2 n_w + 2 n_v paths][kmax + 1 step counts][rec_bytes]
  const uint16_t *kv, *kw;  // [scene][n_v], [scene][n_w]
  const uint8_t *dirv, *dirw; // 1 = ramping up
  // launch 3 walks every scene's grid in its own fork-step order:
  // thread position `pos` -> the pos-th sample when all are sorted by max(kv, kw).  No per-sample table: with
  // rows sorted by kv (row_perm), columns by kw (col_perm) and lvl_rows[k] / lvl_cols[k] = how many rows / columns
  // fork before step k, the samples that fork before step k are the lvl_rows[k] x lvl_cols[k] corner of the sorted
  // grid, and level k itself is (rows of level k) x (columns up to level k) followed by (rows before level k) x
  // (columns of level k).  A warp is therefore 32 samples that fork together whatever the aspect of the grid.
  const uint32_t *col_perm, *row_perm; // [scene][n_w], [scene][n_v]
  const uint32_t *lvl_rows, *lvl_cols; // [scene][kmax + 2]
  // one-wave launches: chunk_map[idx / 32] = which 32 sorted samples warp idx / 32 takes (deals long and short
  // warps evenly over blocks and schedulers); nullptr = identity
  const uint32_t *chunk_map;
  uint64_t scene_stride;
  uint32_t rec_bytes, kmax;
  uint32_t mode, epoch;     // 0 off, 1 / 2 path writers, 3 reader, 4 both path stages in one launch of the
                            // warp-per-path writer (tile 0 of a scene = the 4 doubly saturated paths; the others
                            // wait for the record they continue from: flag = SfwCkptHdr::epoch == epoch)
};

struct SfwBatchDev {
  const SfwSceneDev *scenes;
  // pedestrians are stored as PAIRS (2k, 2k+1), one float4 per pair and quantity, so that the two
  // halves of a packed FP32x2 register pair load with one LDS.128 (scene frame, FP32):
  const float4 *pedPos;  // x0, x1, y0, y1
  const float4 *pedVel;  // vx0, vx1, vy0, vy1
  const float4 *pedGoal; // gx0, gx1, gy0, gy1
  const float4 *pedPar;  // goal_r^2 (0,1), desired_velocity (0,1)
  const float4 *pedPar2; // obs_scale (0,1), desired_velocity^2 (0,1)
  const uint8_t *goal_bits; // [2 * pairs] 1 = pedestrian has a goal (same flags as SfwSceneDev::goal_mask)
  const float2 *obst; // obstacle points (scene frame) * log2(e)/sigma
  const uint32_t *groups; // per-scene group tables (SfwSceneDev::grp_off)
  const double2 *footprint;
  const uint8_t *maps; // costmap slots: map_rows rows of map_pitch bytes each
  const double *linvels;
  const double *angvels;
  float *costs;        // [n_scenes][n_v*n_w]
  uint16_t *npts;      // [n_scenes][n_v*n_w] trajectory points recorded before the rollout stopped
  SfwBest *best;       // [n_scenes]
  // Small results are ALSO stored straight into the pinned host landing buffer (mapped, zero-copy): the stores ride
  // over PCIe while the kernel runs and sfw_download has nothing left to copy across the bus.  nullptr: off.
  float *costs_host;
  SfwBest *best_host;
  SfwBlockBest *blockbest; // [n_scenes][tiles_per_scene]
  unsigned int *counters;  // [n_scenes] tiles finished (self-resetting)
  uint32_t map_pitch, map_rows;
  uint32_t n_scenes, n_v, n_w;
  uint32_t scene_base, launch_scenes; // thread-per-trajectory launches may cover scenes [scene_base, scene_base +
                                      // launch_scenes) only (a batch is scored piece by piece while the rest of it is
                                      // still on its way to the device); launch_scenes == 0: all n_scenes
  uint32_t row_begin, row_end; // linvel rows scored by this launch
  uint32_t tiles_per_scene;
  uint32_t win_wp, win_h; // staged window box (padded width, rows); 0 => read the map from global
  int32_t num_steps;
  uint32_t score_zero; // 1: the (0,0) sample is scored like any other (single-sample calls, sfw_set_zero_sample)
  unsigned int *status; // mapped pinned word [0]: a kernel that has to give up says why (SFW_DEVSTAT_*)
  double dt;
  // ControllerParams on the path
  double max_vel_x, acc_x, acc_th;
  double w_vel, w_dist, w_ang, w_map, w_soc;
  float rr2; // robot_radius * robot_radius evaluated in float (reference sfw_planner.cpp:617)
  // lightsfm constants folded for the FP32 force evaluator
  float lambda;     // lambda
  float gamma;      // gamma
  float c_d;        // log2(e) / gamma            : -d/B * log2e = -d * rL * c_d
  float c_np;       // n'^2 * log2(e)
  float c_n;        // n^2 * log2(e)
  float k_soc;      // forceFactorSocial
  float kd_tau;     // forceFactorDesired / relaxationTime
  float inv_tau;    // 1 / relaxationTime
  float c_obs;      // log2(e) / sigma
  float dtf;        // (float)dt
  float k_gaze, k_coh, k_rep; // forceFactorGroupGaze / Coherence / Repulsion
  float pad2;
  SfwExchangeDev xchg;
  SfwShareDev share;
};

#endif
