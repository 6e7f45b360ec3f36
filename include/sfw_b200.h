/*
 * sfw_b200.h — C ABI of the B200-native DWA + social-force trajectory sampler/scorer.
 *
 * This is the drop-in boundary for ONE path of robotics-upo/social_force_window_planner: the
 * (v, w) sampling / scoring loop of SFWPlanner::findBestAction
 * (reference src/sfw_planner.cpp:338-417) together with everything it calls per sample:
 * SFWPlanner::scoreTrajectory (:475-676), SFWPlanner::computeSocialWork (:678-705),
 * WorldModel/CostmapModel::footprintCost (include/.../world_model.hpp:45-75,
 * src/costmap_model.cpp:21-121, include/.../line_iterator.hpp:37-124) and the lightsfm
 * calls sfm::SFM.computeForces / updatePosition (:592, :594, :697).
 *
 * Plain C: pointers + sizes only, no C++/torch types.  Every entry point returns an int status
 * (SFW_OK == 0, negative on error) and never throws.  There is NO CPU fallback: if no CUDA
 * device is usable sfw_create fails with SFW_ERR_CUDA and nothing else can be called.
 *
 * Threading: one in-flight call per context (the reference holds configuration_mutex_ for the
 * whole findBestAction, src/sfw_planner.cpp:123-454).  Contexts are independent.
 */
#ifndef SFW_B200_H
#define SFW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFW_ABI_VERSION 4

/* status codes */
#define SFW_OK 0
#define SFW_ERR_ARG (-1)         /* null pointer / inconsistent sizes */
#define SFW_ERR_CUDA (-2)        /* CUDA runtime/driver failure; see sfw_last_error */
#define SFW_ERR_UNSUPPORTED (-3) /* input outside what the kernels implement (see sfw_last_error) */
#define SFW_ERR_STATE (-4)       /* call order violated (e.g. sfw_run before sfw_upload) */

/* per-trajectory cost codes written to the cost vector */
#define SFW_COST_INVALID (-1.0f) /* reference returns -1.0 on any violation (sfw_planner.cpp:546-573,624) */
#define SFW_COST_SKIPPED (-2.0f) /* the (0,0) sample is never scored (sfw_planner.cpp:349-352) */

/*
 * ControllerParams fields that the scoring path reads (reference
 * include/social_force_window_planner/sfw_planner.hpp:55-227; used at
 * src/sfw_planner.cpp:356-358,519,527,617,654,663-667).  Types follow the reference
 * (robot_radius_ is a float there, hpp:208, and is squared in float at cpp:617).
 */
typedef struct SfwParams {
  double max_vel_x;       /* max_trans_vel  (0.7) */
  double max_trans_acc;   /* max_trans_acc  (1.0) -> acc_x; acc_y is always 0 (cpp:357) */
  double max_rot_acc;     /* max_rot_acc    (1.0) -> acc_theta */
  double sim_time;        /* sim_time       (1.0) */
  double sim_granularity; /* sim_granularity (0.025): num_steps = int(sim_time/gran + 0.5) (cpp:519) */
  float robot_radius;     /* robot_radius   (0.35f): robot-pedestrian collision radius (cpp:617) */
  float reserved0;
  double social_weight;   /* 1.2 */
  double costmap_weight;  /* 2.0 */
  double angle_weight;    /* 0.7 */
  double distance_weight; /* 1.0 */
  double vel_weight;      /* velocity_weight 1.0 */
} SfwParams;

/*
 * lightsfm sfm::Parameters (un-vendored dependency; defaults restated in SURVEY.md Appendix B).
 * The reference never changes them, so one set applies to every agent of a call.
 * Pass NULL wherever this struct is accepted to get the defaults below.
 */
typedef struct SfwSfmParams {
  double force_factor_desired;         /* 2.0  */
  double force_factor_obstacle;        /* 10.0 */
  double force_sigma_obstacle;         /* 0.2  */
  double force_factor_social;          /* 2.1  */
  double force_factor_group_gaze;      /* 3.0  */
  double force_factor_group_coherence; /* 2.0  */
  double force_factor_group_repulsion; /* 1.0  */
  double lambda;                       /* 2.0  */
  double gamma;                        /* 0.35 */
  double n;                            /* 2.0  */
  double n_prime;                      /* 3.0  */
  double relaxation_time;              /* 0.5  */
} SfwSfmParams;

/*
 * Robot inputs of one scoreTrajectory sweep (src/sfw_planner.cpp:475-481).
 * x..vtheta are the values findBestAction passes after narrowing to float (cpp:145-152): the
 * CALLER narrows (the host shim does), the library uses them as given.
 * agent_* is agents[0] exactly as SFMSensorInterface::getAgents returned it
 * (src/sensor_interface.cpp:553-579,618-631): odom pose, ROBOT-frame velocity, radius.  It is the
 * robot state the first computeForces sees (cpp:592, i == 0) before the rollout overwrites it.
 */
typedef struct SfwRobot {
  double x, y, theta;             /* start pose of the rollout */
  double vx, vy, vtheta;          /* current velocity (vy never changes: acc_y == 0) */
  double wpx, wpy;                /* waypoint scored against (cpp:643-652) */
  double agent_x, agent_y;        /* agents[0].position */
  double agent_vx, agent_vy;      /* agents[0].velocity (robot frame, sensor_interface.cpp:575) */
  double agent_radius;            /* agents[0].radius (sensor_interface.cpp:34) */
} SfwRobot;

/*
 * One pedestrian = agents[1..P] as built by SFMSensorInterface::peopleCb
 * (src/sensor_interface.cpp:447-504): position, velocity, one naive goal, desired speed, radius,
 * group id.  has_goal == 0 means an empty goal list (lightsfm then brakes the agent).
 */
typedef struct SfwPed {
  double x, y;
  double vx, vy;
  double goal_x, goal_y;
  double goal_radius;
  double desired_velocity;
  double radius;
  int32_t has_goal;
  int32_t group_id; /* -1 = none (sensor_interface.cpp:449) */
  int32_t id;       /* people tag id.  Carried for the caller's bookkeeping only: lightsfm's computeForces(agent,
                     * others) skips an `other` whose id equals the agent's, and the reference never sets the
                     * robot's id (sfm::Agent::id is uninitialised upstream), so whether a
                     * pedestrian is skipped there is undefined; here every pedestrian always counts. */
  int32_t reserved0;
} SfwPed;

/*
 * One planning scene: robot + costmap + pedestrians + the obstacle points shared by all agents
 * (sensor_interface.cpp:513-524) + footprint polygon (robot frame).  All pointers are caller
 * owned host memory and are only read during the call.
 * costmap: row-major uint8, cost(mx,my) = costmap[my*size_x + mx] (nav2 Costmap2D layout).
 */
typedef struct SfwScene {
  SfwRobot robot;
  const uint8_t *costmap;
  uint32_t size_x, size_y;
  double resolution, origin_x, origin_y;
  const SfwPed *peds;
  uint32_t n_peds;
  uint32_t n_obstacles;
  const double *obstacles_xy; /* n_obstacles pairs (x,y) */
  const double *footprint_xy; /* n_footprint pairs (x,y), robot frame; <3 => centre-cell mode */
  uint32_t n_footprint;
  uint32_t reserved0;
} SfwScene;

/* Winner of one scene (reference: best_traj / vx, vt at src/sfw_planner.cpp:394-414,426-468). */
typedef struct SfwBest {
  int32_t valid;  /* 0 => no valid trajectory: findBestAction returns false, zero twist (cpp:456-468) */
  uint32_t index; /* i = i_v * n_w + i_w (cpp:342-416) */
  float cost;
  float reserved0;
  double v, w;    /* chosen (linvel, angvel); 0,0 when !valid */
} SfwBest;

/* Limits used to pre-size device buffers; all may be exceeded later (buffers grow). */
typedef struct SfwLimits {
  uint32_t max_scenes;
  uint32_t max_samples; /* n_v * n_w */
  uint32_t max_peds;
  uint32_t max_obstacles;
  uint32_t max_cells; /* size_x * size_y */
} SfwLimits;

typedef struct sfw_ctx sfw_ctx;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int sfw_abi_version(void);
/* device: CUDA ordinal.  stream: a cudaStream_t to run on, or NULL for a private non-blocking
 * stream owned by the context.  limits may be NULL. */
int sfw_create(sfw_ctx **out, int device, void *stream, const SfwLimits *limits);
int sfw_destroy(sfw_ctx *ctx);
const char *sfw_last_error(const sfw_ctx *ctx); /* ctx may be NULL: last sfw_create error */
void sfw_default_params(SfwParams *p);          /* reference header defaults (sfw_planner.hpp:56-66) */
void sfw_default_sfm_params(SfwSfmParams *p);   /* lightsfm defaults */

/* ---- the hot path --------------------------------------------------------------------------
 * sfw_score: one scene, the whole (linvels x angvels) grid; replaces the double loop of
 * SFWPlanner::findBestAction (src/sfw_planner.cpp:345-417).  costs_out (n_v*n_w floats, may be
 * NULL) receives the per-trajectory cost vector, best_out the arg-min with the reference's
 * tie-breaks.  Synchronous: host buffers are filled on return.
 * sfw_score_batch: n_scenes independent scenes sharing params, sample arrays and array sizes
 * (size_x/size_y/n_peds/n_obstacles/n_footprint may differ per scene up to the first scene's
 * values being an upper bound is NOT required; the library sizes for the maximum). */
int sfw_score(sfw_ctx *ctx, const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
              const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
              float *costs_out, SfwBest *best_out);
int sfw_score_batch(sfw_ctx *ctx, const SfwParams *params, const SfwSfmParams *sfm,
                    const SfwScene *scenes, uint32_t n_scenes, const double *linvels, uint32_t n_v,
                    const double *angvels, uint32_t n_w, float *costs_out, SfwBest *best_out);

/* ---- split form of the same call (staging / launch / fetch), for callers that keep scenes
 * resident on the device across ticks and for benchmarking the kernel alone ------------------ */
int sfw_upload(sfw_ctx *ctx, const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scenes,
               uint32_t n_scenes, const double *linvels, uint32_t n_v, const double *angvels,
               uint32_t n_w);          /* pack + async H2D on the context stream */
int sfw_run(sfw_ctx *ctx);             /* launch the scorer on the staged scenes (async) */
int sfw_download(sfw_ctx *ctx, float *costs_out, SfwBest *best_out); /* D2H + stream sync */
int sfw_sync(sfw_ctx *ctx);            /* cudaStreamSynchronize on the context stream */

/* Kernel selection.  SFW_POLICY_AUTO (default): crowds of more than 64 pedestrians use the block-per-
 * trajectory kernel, smaller ones the thread-per-trajectory kernel — except SMALL GRIDS, which are latency
 * bound there (a tick costs what one warp costs) and go to the block-per-trajectory kernel while that takes
 * few waves (a 5 x 9 tick with 20 pedestrians: 0.23 ms instead of 0.90 ms).  The two kernels evaluate the same
 * model with different summation orders (costs agree to ~1e-6 relative), so callers that need bit-identical
 * cost vectors whatever the batch size or rank count pin SFW_POLICY_THROUGHPUT (never the small-grid switch);
 * SFW_POLICY_LATENCY always takes the block-per-trajectory kernel.  Applies from the next sfw_upload. */
#define SFW_POLICY_AUTO 0
#define SFW_POLICY_THROUGHPUT 1
#define SFW_POLICY_LATENCY 2
int sfw_set_policy(sfw_ctx *ctx, int policy);

/* Host workers.  A batch of scenes (sfw_upload / sfw_score_batch with >= 8 scenes) is packed into the staging
 * arena by n_threads host threads, one scene per work item, and goes to the device in pieces while the next piece
 * is being packed; big cost-vector downloads are copied out by the same workers.  A single scene (the control
 * tick) is always packed on the calling thread.  Default: min(16, cores / visible devices); 1 = no workers. */
int sfw_set_host_threads(sfw_ctx *ctx, int n_threads);

/* The (0,0) sample.  The reference skips (linvel, angvel) == (0,0) INSIDE its grid loop only
 * (src/sfw_planner.cpp:349-352); its single scoreTrajectory calls — rotate in place (:225-245), approach the goal
 * (:298-330) — always evaluate, whatever the velocities.  score_it = 1 scores the (0,0) sample like any other
 * (set it for single-sample calls; a rotate-in-place with min_in_place_vel_th == 0 is such a sample); 0 (default)
 * marks it SFW_COST_SKIPPED.  Applies from the next sfw_upload. */
int sfw_set_zero_sample(sfw_ctx *ctx, int score_it);

/* Rollout prefix sharing (on by default; applies from the next sfw_upload).  With acceleration limits, samples
 * whose velocity is still ramping at the full +-a*dt per step are identical for their first steps; on dense
 * grids (one wave or many) the library simulates those shared prefixes once and starts every sample from the state of
 * its fork point.  The cost vector is bit-identical either way (same arithmetic, same order).
 * on: 0 = never, 1 = when the library's cost model says it pays (default), 2 = whenever the staged batch allows
 * it (grids of >= 1024 samples, >= 8 steps, some saturated ramp; for tests and experiments). */
int sfw_set_prefix_sharing(sfw_ctx *ctx, int on);

/* Far-field cutoff of the pedestrians' obstacle force (applies from the next sfw_upload).  lightsfm sums
 * k/M * exp(-(|p - o| - r)/sigma) over every obstacle point o (reached through the
 * computeForces call at reference src/sfw_planner.cpp:592).  The library stores the points as spatially compact
 * clusters of 8 and a pedestrian skips a cluster all of whose terms are below 2^-cutoff_log2 of the force factor
 * k/M.  Default 24 (one FP32 ulp: with sigma = 0.2 m, r = 0.35 m that is > 3.7 m away); <= 0 switches the
 * cutoff off (every term is summed).  The robot's own obstacle force, which enters the social work directly,
 * always sums every point. */
int sfw_set_obstacle_cutoff(sfw_ctx *ctx, double cutoff_log2);

/* The obstacle layout sfw_upload builds for one scene, computed on the host (no context, no GPU): the points
 * scaled by log2(e)/sigma relative to (ref_x, ref_y) and grouped into clusters of 8 behind a 2-slot header
 * {centre x, centre y}, {reach^2, 0} — 10 float2 slots per cluster, ceil(n / 8) clusters, the last one padded with
 * points 1e15 away.  r_max = largest agent radius of the scene (metres).  For tests, and for integrators who want
 * to see what the cutoff will skip.  Returns the number of float2 slots; writes min(slots, slots_cap) of them
 * (2 floats each) to slots_out (may be NULL to only query the count). */
uint32_t sfw_obstacle_layout(const double *obstacles_xy, uint32_t n, double ref_x, double ref_y, double sigma,
                             double r_max, double cutoff_log2, float *slots_out, uint32_t slots_cap);

/* Restrict the next sfw_run calls to linvel rows [row_begin, row_end) of every staged scene
 * (multi-GPU sharding of a single scene across ranks: each rank scores a slab and the winners
 * are all-gathered by the caller).  Rows outside the slab get SFW_COST_SKIPPED. */
int sfw_set_row_slab(sfw_ctx *ctx, uint32_t row_begin, uint32_t row_end);

/* Winner trajectory points (Trajectory::x_pts_/y_pts_/th_pts_, src/trajectory.cpp:36-40) for one
 * sample of one staged scene: fills up to max_points (x,y,theta) triples, returns the number of
 * points the rollout recorded in *n_points (stops at the first illegal pose like the reference). */
int sfw_trajectory_points(sfw_ctx *ctx, uint32_t scene, uint32_t sample_index, double *xyz_out,
                          uint32_t max_points, uint32_t *n_points);

/* The same for many samples in one launch — the points of the RViz markers the reference fills for EVERY
 * evaluated sample (src/sfw_planner.cpp:366-374): samples first, first + stride, ... (count of them) of one
 * staged scene.  xyz_out: count slots of max_points (x,y,theta) triples; n_points_out[count]: points each
 * rollout recorded (may exceed max_points; only min(n, max_points) triples are written). */
int sfw_marker_points(sfw_ctx *ctx, uint32_t scene, uint32_t first, uint32_t stride, uint32_t count,
                      double *xyz_out, uint32_t max_points, uint16_t *n_points_out);

/* SFWPlanner::mayIStop (src/sfw_planner.cpp:718-765; unreachable upstream — its call at :637 is commented out —
 * provided for completeness): brake from (vl_x, vl_y, va) at the acceleration limits of the last sfw_upload and
 * check the footprint on that scene's costmap at every pose.  *can_stop = 1 when the robot comes to rest without
 * an illegal footprint, *steps = poses checked. */
int sfw_may_i_stop(sfw_ctx *ctx, uint32_t scene, double vl_x, double vl_y, double va, double x, double y, double th,
                   double dt, int32_t *can_stop, uint32_t *steps);

/* ---- the step before the path: laser scan -> obstacle points ----------------------------------
 * SFMSensorInterface::laserCb (reference src/sensor_interface.cpp:103-229): keep the beams that are
 * finite and closer than max_obstacle_dist (:120-122), polar -> cartesian in FLOAT (:124-125; the
 * scan angle is accumulated in float, :127), move them to the controller frame when the scan's frame
 * differs (:143-170, the planar transform tf2 would apply is passed in), drop every point within
 * person_radius (float hypot, :213-218) of a detected person, keep the beam order.  The surviving
 * points are the obstacles1 list every agent of the scene carries (:513-524), i.e. SfwScene::
 * obstacles_xy.  One block per scan; n_scans scans (one per scene of a batch) in one launch. */
typedef struct SfwLaserScan {
  const float *ranges;     /* sensor_msgs/LaserScan::ranges */
  uint32_t n_ranges;
  float angle_min, angle_increment;
  int32_t has_tf;          /* 0: scan already in the controller frame */
  double tf_x, tf_y, tf_yaw; /* laser frame -> controller frame */
  const double *people_xy; /* n_people (x,y) pairs, controller frame (cpp:176-209) */
  uint32_t n_people;
  uint32_t reserved0;
} SfwLaserScan;

/* points_xy_out: n_scans slots of max_points_per_scan (x,y) pairs; n_points_out[n_scans] = points kept
 * (never more than n_ranges; SFW_ERR_ARG if a slot is too small).  Synchronous. */
int sfw_laser_obstacles(sfw_ctx *ctx, const SfwLaserScan *scans, uint32_t n_scans, float max_obstacle_dist,
                        float person_radius, double *points_xy_out, uint32_t max_points_per_scan,
                        uint32_t *n_points_out);

/* ---- multi-GPU: winner exchange fused into the scorer's epilogue -------------------------------
 * One process per GPU.  After sfw_exchange_connect, each sfw_run stores its SfwBest records directly into every
 * rank's gather buffer over NVLink (peer mappings of cudaIpc handles) from the kernel that reduces them; no
 * collective follows the kernel.  Two sharding modes (reference loop being sharded: src/sfw_planner.cpp:345-417):
 *   scene batch (BASELINE configs[3])   rank q scores its own block of scenes; sfw_exchange_fetch returns all of
 *                                       them, rank after rank = global scene order of a block partition
 *   row slabs of one batch (configs[1], every rank stages the SAME scenes and scores its linvel rows
 *   configs[4] on several GPUs)         (sfw_set_row_slab); sfw_exchange_merge reduces the world slab winners of
 *                                       every scene on the device with the reference's tie-break order
 *                                       (:394-414), so every rank holds the full grid's winner
 *   sfw_exchange_export   allocate this rank's gather buffer (max_scenes per rank), write its 64-byte
 *                         cudaIpcMemHandle_t to handle_out; the caller ships the handles to all ranks
 *   sfw_exchange_connect  handles = world x 64 bytes in rank order (own slot ignored); world <= 8
 *   sfw_exchange_connect_local  the same for `world` contexts of ONE process (one per GPU, or several on one
 *                         GPU): ctxs[r] becomes rank r; peers are reached through direct / peer-enabled pointers
 *   sfw_exchange_expect   scenes_per_rank[world]: how many scenes each rank stages per tick from now on (they may
 *                         differ); NULL / never called = every rank stages what this rank stages
 *   sfw_exchange_set_timeout  bound of the device-side arrival wait (default 10 s)
 *   sfw_exchange_sync     enqueue a device-side wait on the context stream until every rank's records of the
 *                         latest sfw_run have arrived (asynchronous for the host)
 *   sfw_exchange_fetch    wait (as above) + copy the gathered records to all_best_out[sum of scenes_per_rank]
 *   sfw_exchange_merge    wait + device-side merge of the slab winners; merged_out[n_scenes] (NULL: leave them
 *                         on the device, sfw_exchange_merged_device, and do not synchronise)
 * A peer that dies, skips a tick or stages fewer scenes than announced does not hang the others: the wait gives
 * up after the timeout and fetch / merge return SFW_ERR_STATE naming the rank that did not deliver.  An empty
 * row slab still delivers (invalid) records. */
int sfw_exchange_export(sfw_ctx *ctx, uint32_t max_scenes, void *handle_out);
int sfw_exchange_connect(sfw_ctx *ctx, uint32_t rank, uint32_t world, const void *handles);
int sfw_exchange_connect_local(sfw_ctx *const *ctxs, uint32_t world);
int sfw_exchange_expect(sfw_ctx *ctx, const uint32_t *scenes_per_rank);
int sfw_exchange_set_timeout(sfw_ctx *ctx, double seconds);
int sfw_exchange_sync(sfw_ctx *ctx);
int sfw_exchange_fetch(sfw_ctx *ctx, SfwBest *all_best_out);
int sfw_exchange_merge(sfw_ctx *ctx, SfwBest *merged_out);
const void *sfw_exchange_device_buffer(sfw_ctx *ctx); /* device [world][max_scenes] SfwBest of the latest run */
const void *sfw_exchange_merged_device(sfw_ctx *ctx); /* device [max_scenes] SfwBest of the latest merge */

/* ---- introspection (benchmark / interop) ---------------------------------------------------- */
void *sfw_stream(sfw_ctx *ctx);                 /* cudaStream_t the context launches on */
const float *sfw_device_costs(sfw_ctx *ctx);    /* device cost vector of the last run */
const void *sfw_device_best(sfw_ctx *ctx);      /* device SfwBest[n_scenes] of the last run */
uint64_t sfw_kernel_launches(const sfw_ctx *ctx); /* kernels launched by this context so far */
/* algorithmic bytes of the staged batch (SURVEY.md section 8d formula) */
uint64_t sfw_algorithmic_bytes(const sfw_ctx *ctx);
/* bytes the last sfw_upload copied host->device, and bytes sfw_download copies back when both the
 * cost vector and the winners are requested (bench.py's e2e accounting) */
uint64_t sfw_h2d_bytes(const sfw_ctx *ctx);
uint64_t sfw_d2h_bytes(const sfw_ctx *ctx);
/* name of the kernel variant the last sfw_run dispatched to (for logs/profiles) */
const char *sfw_last_kernel(const sfw_ctx *ctx);
/* threads per block of the plan of the staged batch (the thread-per-trajectory kernel's tile, or the block size of the
 * block-per-trajectory kernel: 256, or 128 for a small crowd on a grid where that saves a wave); 0 before sfw_upload */
uint32_t sfw_block_threads(const sfw_ctx *ctx);
/* rollout prefix sharing of the staged batch: mean number of leading steps a sample takes from a shared path
 * instead of simulating them itself (0 when sharing is off for this batch) */
double sfw_shared_prefix_steps(const sfw_ctx *ctx);
/* obstacle far-field cutoff of the staged batch: fraction of (pedestrian, obstacle cluster) combinations that
 * are out of reach at the pedestrians' START positions (an estimate of the skipped share of the obstacle sums;
 * 0 when the cutoff is off or there are no obstacles) */
double sfw_obstacle_skip_fraction(const sfw_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SFW_B200_H */
