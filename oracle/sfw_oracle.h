/*
 * sfw_oracle.h — CPU restatement (double precision, plain C) of the reference's (v,w) scoring path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker or the timed CPU baseline.  The product (libsfw_b200.so) never links,
 * loads or calls it.
 *
 * PARITY PINNING: the rollout / footprint / cost / arg-min logic is checked bit-for-bit against
 * the reference's own sources compiled unmodified (oracle/_ref, built by oracle/Makefile from
 * /root/reference/src/{sfw_planner,costmap_model,trajectory}.cpp).  The lightsfm arithmetic
 * (desired / obstacle / social / group force, position update) is a restatement of an
 * UN-VENDORED dependency with no pinned version (reference package.xml:31): at that boundary
 * parity is UNPINNED — both oracle and oracle/_ref use the same restated formulas.
 */
#ifndef SFW_ORACLE_H
#define SFW_ORACLE_H

#include "../include/sfw_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-trajectory distance to the nearest discontinuity of the model, so parity tests can tell a
 * float-vs-double rounding difference from a branch flip (goal pop, collision, sign(theta)). */
typedef struct SfwOracleMargins {
  double goal;      /* min over steps/peds of | |goal - p| - goal_radius |  (metres) */
  double collision; /* min over steps/peds of | |robot - ped| - robot_radius | and, for grouped pedestrians,
                     * of | |p_a - p_b| - (r_a + r_b) | (group repulsion switches on at contact) (metres) */
  double theta;     /* min over non-negligible pair evaluations of |theta| (radians) */
  double cell;      /* min distance (metres) of any rasterised world point to a cell boundary */
} SfwOracleMargins;

/* Branch probe.  The model is discontinuous at a handful of decisions: a pedestrian's goal pops when it comes
 * within the goal radius (lightsfm updatePosition / computeDesiredForce), the rollout dies when the robot touches
 * a pedestrian (reference src/sfw_planner.cpp:613-627), lightsfm's angular interaction term carries
 * sign(theta) (jumps at theta = 0 and |theta| = pi), group repulsion switches on at contact.  A float evaluator
 * whose state is a few ulps away may legitimately take such a decision the other way when it is taken within
 * rounding distance of its switching surface.  The probe (i) records every decision taken within the given
 * margins and (ii) re-runs a trajectory with listed decisions FORCED to a given value, so that a parity test
 * can require: GPU cost within 1e-4 of the oracle's cost on the branch the GPU took, validity equal to that
 * branch's — instead of loosening the tolerance near a discontinuity. */
#define SFW_EV_GOAL 1      /* a = agent index (1..P), b = 0; step = the step whose force pass sees the decision;
                              decision 1 = goal reached */
#define SFW_EV_COLLISION 2 /* a = 0, b = pedestrian agent index; decision 1 = hit (rollout invalid) */
#define SFW_EV_THETA 3     /* a < b agent indices (0 = robot); decision = sign(theta) used by BOTH directions */
#define SFW_EV_GROUP 4     /* a < b agent indices; decision 1 = touching (repulsion on) */
typedef struct SfwOracleEvent {
  int32_t kind, step, a, b;
  int32_t decision;
  int32_t reserved0;
  double margin; /* distance to the switching surface: metres (goal, collision, group) or radians (theta) */
  double weight; /* theta: magnitude of the angular force term (the jump is twice that); group: distance */
} SfwOracleEvent;

typedef struct SfwOracleProbe {
  double goal_margin, collision_margin, theta_margin; /* record decisions closer than this */
  double theta_min_weight;                            /* ... theta ones only when the term is at least this big */
  const SfwOracleEvent *flips; /* decisions to force: (kind, step, a, b) -> decision */
  uint32_t n_flips;
  uint32_t max_events;
  SfwOracleEvent *events; /* out: first max_events recorded decisions, in order of occurrence */
  uint32_t n_events;      /* out: how many were met (may exceed max_events) */
  uint32_t reserved0;
} SfwOracleProbe;

/* scoreTrajectory with the probe on. */
double sfw_oracle_score_trajectory_probe(const SfwParams *params, const SfwSfmParams *sfm,
                                         const SfwScene *scene, double vx_samp, double vy_samp,
                                         double vtheta_samp, double acc_x, double acc_y, double acc_theta,
                                         SfwOracleProbe *probe);

/* Samples [first, first + count) of the (linvels x angvels) grid with the probe on, farmed over n_threads:
 * costs_out[count], events_out[count][cfg->max_events], n_events_out[count]. */
int sfw_oracle_score_probe(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                           const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                           uint32_t first, uint32_t count, const SfwOracleProbe *cfg, double *costs_out,
                           SfwOracleEvent *events_out, uint32_t *n_events_out, int n_threads);

/* One scoreTrajectory call (reference src/sfw_planner.cpp:475-676).  pts_xyz (nullable) receives
 * the recorded (x,y,theta) points, *n_pts their count.  Returns the cost or -1.0. */
double sfw_oracle_score_trajectory(const SfwParams *params, const SfwSfmParams *sfm,
                                   const SfwScene *scene, double vx_samp, double vy_samp,
                                   double vtheta_samp, double acc_x, double acc_y, double acc_theta,
                                   double *pts_xyz, uint32_t max_pts, uint32_t *n_pts,
                                   SfwOracleMargins *margins);

/* The double loop + arg-min of findBestAction (reference src/sfw_planner.cpp:338-417,426-468).
 * costs_out: n_v*n_w doubles (-1 invalid, -2 skipped (0,0)); margins_out nullable. */
int sfw_oracle_score(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                     const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                     double *costs_out, SfwBest *best_out, SfwOracleMargins *margins_out);

/* Same cost vector, samples farmed over n_threads pthreads (timing baseline; the arg-min is
 * re-run serially over the cost vector so the result is identical to sfw_oracle_score). */
int sfw_oracle_score_mt(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                        const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                        double *costs_out, SfwBest *best_out, int n_threads);

/* Arg-min of a cost vector with the reference's sequential tie-break semantics
 * (src/sfw_planner.cpp:344,394-414). */
void sfw_oracle_argmin(const double *costs, const double *linvels, uint32_t n_v,
                       const double *angvels, uint32_t n_w, SfwBest *best_out);

/* WorldModel::footprintCost(x,y,theta,spec) -> CostmapModel::footprintCost
 * (world_model.hpp:45-75, costmap_model.cpp:21-92).  Returns -3/-2/-1 or the max cell cost. */
double sfw_oracle_footprint_cost(const SfwScene *scene, double x, double y, double theta,
                                 double *cell_margin);

/* Bresenham cells of LineIterator (line_iterator.hpp:37-124): writes up to max_cells (x,y) pairs,
 * returns the number of cells of the line. */
int sfw_oracle_line_cells(int x0, int y0, int x1, int y1, int *cells_xy, int max_cells);

/* lightsfm pair social force on `me` from `other` (restated).  Inputs are
 * (px,py,vx,vy) each; out_fxy receives the force already scaled by forceFactorSocial. */
void sfw_oracle_pair_force(const SfwSfmParams *sfm, const double me[4], const double other[4],
                           double out_fxy[2], double *theta_out);

/* lightsfm obstacle force on an agent at (px,py) with radius r from the scene obstacle list. */
void sfw_oracle_obstacle_force(const SfwSfmParams *sfm, double px, double py, double radius,
                               const double *obstacles_xy, uint32_t n_obstacles, double out_fxy[2]);

/* SFMSensorInterface::laserCb (reference src/sensor_interface.cpp:103-229): beams -> obstacle points of
 * one scan, beam order kept.  points_xy_out must hold n_ranges pairs.  Returns the number kept.
 * (The tf2 transform of :143-170 is external; it is restated as the planar rigid transform.) */
uint32_t sfw_oracle_laser_obstacles(const SfwLaserScan *scan, float max_obstacle_dist, float person_radius,
                                    double *points_xy_out);

#ifdef __cplusplus
}
#endif
#endif
