// ref_harness.cpp — C entry points around the reference's OWN planner core, compiled unmodified
// from /root/reference/src/{sfw_planner,costmap_model,trajectory}.cpp (see oracle/Makefile).
//
// TEST INFRASTRUCTURE ONLY (oracle/_ref): used to pin oracle/sfw_oracle.c and, optionally, as the
// "reference" CPU baseline of bench.py.  Never linked into or called by the product library.
//
// What is the reference's own object code here: the (v,w) double loop, tie-breaks, scoreTrajectory
// step ordering, float narrowing, computeSocialWork, footprint rasterisation, Bresenham.
// What is restated (oracle/stubs): ROS/nav2 message + costmap shims and lightsfm (un-vendored).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include <lightsfm/sfm.hpp>
#include <nav2_costmap_2d/costmap_2d.hpp>
#include <nav2_costmap_2d/footprint.hpp>
#include <rclcpp_lifecycle/lifecycle_node.hpp>
#include <social_force_window_planner/sensor_interface.hpp>

// reach scoreTrajectory / linvels_ / angvels_ / footprintCost without touching the reference
#define private public
#include <social_force_window_planner/sfw_planner.hpp>
#undef private

#include "../include/sfw_b200.h"

using social_force_window_planner::SFMSensorInterface;
using social_force_window_planner::SFWPlanner;
using social_force_window_planner::Trajectory;

namespace {

const char *kName = "FollowPath";

struct Rig {
  rclcpp_lifecycle::LifecycleNode::SharedPtr node;
  std::shared_ptr<SFMSensorInterface> iface;
  std::unique_ptr<nav2_costmap_2d::Costmap2D> costmap;
  std::unique_ptr<SFWPlanner> planner;
  std::vector<sfm::Agent> agents;
};

void set_params(rclcpp_lifecycle::LifecycleNode &n, const SfwParams &p) {
  std::string b = std::string(kName) + ".";
  n.set_parameter(b + "max_trans_vel", rclcpp::ParameterValue(p.max_vel_x));
  n.set_parameter(b + "max_trans_acc", rclcpp::ParameterValue(p.max_trans_acc));
  n.set_parameter(b + "max_rot_acc", rclcpp::ParameterValue(p.max_rot_acc));
  n.set_parameter(b + "sim_time", rclcpp::ParameterValue(p.sim_time));
  n.set_parameter(b + "sim_granularity", rclcpp::ParameterValue(p.sim_granularity));
  n.set_parameter(b + "robot_radius", rclcpp::ParameterValue((double)p.robot_radius));
  n.set_parameter(b + "social_weight", rclcpp::ParameterValue(p.social_weight));
  n.set_parameter(b + "costmap_weight", rclcpp::ParameterValue(p.costmap_weight));
  n.set_parameter(b + "angle_weight", rclcpp::ParameterValue(p.angle_weight));
  n.set_parameter(b + "distance_weight", rclcpp::ParameterValue(p.distance_weight));
  n.set_parameter(b + "velocity_weight", rclcpp::ParameterValue(p.vel_weight));
}

void apply_sfm(sfm::Agent &a, const SfwSfmParams *s) {
  if (!s)
    return;
  a.params.forceFactorDesired = s->force_factor_desired;
  a.params.forceFactorObstacle = s->force_factor_obstacle;
  a.params.forceSigmaObstacle = s->force_sigma_obstacle;
  a.params.forceFactorSocial = s->force_factor_social;
  a.params.forceFactorGroupGaze = s->force_factor_group_gaze;
  a.params.forceFactorGroupCoherence = s->force_factor_group_coherence;
  a.params.forceFactorGroupRepulsion = s->force_factor_group_repulsion;
  a.params.lambda = s->lambda;
  a.params.gamma = s->gamma;
  a.params.n = s->n;
  a.params.nPrime = s->n_prime;
  a.params.relaxationTime = s->relaxation_time;
}

// agents[] exactly as SFMSensorInterface builds them (reference src/sensor_interface.cpp:32-37,
// 447-504, 513-524, 553-579): [0] robot (teleoperated, no goals), [1..P] pedestrians with one
// goal each, every agent carrying the same obstacle list.
std::vector<sfm::Agent> make_agents(const SfwScene &sc, const SfwSfmParams *s) {
  std::vector<utils::Vector2d> obs;
  for (uint32_t i = 0; i < sc.n_obstacles; ++i)
    obs.emplace_back(sc.obstacles_xy[2 * i], sc.obstacles_xy[2 * i + 1]);
  std::vector<sfm::Agent> ag(sc.n_peds + 1);
  sfm::Agent &r = ag[0];
  r.id = -1; // indeterminate upstream (never set); fixed to -1 here and in the oracle
  r.position.set(sc.robot.agent_x, sc.robot.agent_y);
  r.velocity.set(sc.robot.agent_vx, sc.robot.agent_vy);
  r.linearVelocity = std::sqrt(sc.robot.agent_vx * sc.robot.agent_vx + sc.robot.agent_vy * sc.robot.agent_vy);
  r.radius = sc.robot.agent_radius;
  r.teleoperated = true;
  r.cyclicGoals = false;
  r.groupId = -1;
  r.obstacles1 = obs;
  apply_sfm(r, s);
  for (uint32_t j = 0; j < sc.n_peds; ++j) {
    const SfwPed &p = sc.peds[j];
    sfm::Agent &a = ag[j + 1];
    a.id = p.id;
    a.groupId = p.group_id;
    a.position.set(p.x, p.y);
    a.velocity.set(p.vx, p.vy);
    a.linearVelocity = a.velocity.norm();
    a.yaw = utils::Angle::fromRadian(std::atan2(p.vy, p.vx));
    a.radius = p.radius;
    a.teleoperated = false;
    a.desiredVelocity = p.desired_velocity;
    if (p.has_goal) {
      sfm::Goal g;
      g.center.set(p.goal_x, p.goal_y);
      g.radius = p.goal_radius;
      a.goals.push_back(g);
    }
    a.obstacles1 = obs;
    apply_sfm(a, s);
  }
  return ag;
}

std::unique_ptr<Rig> make_rig(const SfwParams &p, const SfwSfmParams *s, const SfwScene &sc) {
  std::unique_ptr<Rig> rig(new Rig);
  rig->node = std::make_shared<rclcpp_lifecycle::LifecycleNode>();
  set_params(*rig->node, p);
  rig->iface = std::make_shared<SFMSensorInterface>();
  rig->agents = make_agents(sc, s);
  rig->iface->setAgents(rig->agents);
  rig->costmap.reset(new nav2_costmap_2d::Costmap2D(sc.size_x, sc.size_y, sc.resolution, sc.origin_x,
                                                    sc.origin_y, sc.costmap));
  std::vector<geometry_msgs::msg::Point> fp(sc.n_footprint);
  for (uint32_t i = 0; i < sc.n_footprint; ++i) {
    fp[i].x = sc.footprint_xy[2 * i];
    fp[i].y = sc.footprint_xy[2 * i + 1];
  }
  rig->planner.reset(new SFWPlanner(rig->node, kName, rig->iface, *rig->costmap, fp));
  return rig;
}

geometry_msgs::msg::PoseStamped pose_of(double x, double y, double yaw) {
  geometry_msgs::msg::PoseStamped ps;
  ps.pose.position.x = x;
  ps.pose.position.y = y;
  ps.pose.orientation.z = std::sin(yaw * 0.5);
  ps.pose.orientation.w = std::cos(yaw * 0.5);
  return ps;
}

} // namespace

extern "C" {

// One SFWPlanner::scoreTrajectory call of the reference (src/sfw_planner.cpp:475-676).
double sfw_ref_score_trajectory(const SfwParams *params, const SfwSfmParams *sfm,
                                const SfwScene *scene, double vx_samp, double vy_samp,
                                double vtheta_samp, double acc_x, double acc_y, double acc_theta,
                                double *pts_xyz, uint32_t max_pts, uint32_t *n_pts) {
  auto rig = make_rig(*params, sfm, *scene);
  rig->planner->params_.get(rig->node.get(), kName); // findBestAction does this each tick (:125)
  const SfwRobot &R = scene->robot;
  Trajectory t;
  double c = rig->planner->scoreTrajectory(R.x, R.y, R.theta, R.vx, R.vy, R.vtheta, vx_samp, vy_samp,
                                           vtheta_samp, acc_x, acc_y, acc_theta, R.wpx, R.wpy,
                                           rig->agents, t);
  if (n_pts) {
    *n_pts = t.getPointsSize();
    for (uint32_t i = 0; pts_xyz && i < t.getPointsSize() && i < max_pts; ++i)
      t.getPoint(i, pts_xyz[3 * i], pts_xyz[3 * i + 1], pts_xyz[3 * i + 2]);
  }
  return c;
}

// Cost vector: the reference's scoreTrajectory for every sample, arguments as the double loop of
// findBestAction passes them (src/sfw_planner.cpp:345-358).  best_out (nullable) is produced by
// the reference's findBestAction itself, run on a one-pose plan at the waypoint with the sample
// sets overridden — its arg-min/tie-break code is therefore the reference's, not a restatement.
// best_out->index/cost are recovered from (v,w) since findBestAction only returns the twist.
int sfw_ref_score(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                  const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                  double *costs_out, SfwBest *best_out) {
  if (!params || !scene || !linvels || !angvels)
    return SFW_ERR_ARG;
  auto rig = make_rig(*params, sfm, *scene);
  SFWPlanner &pl = *rig->planner;
  pl.params_.get(rig->node.get(), kName);
  pl.linvels_.assign(linvels, linvels + n_v);
  pl.angvels_.assign(angvels, angvels + n_w);
  pl.initializeMarkers();
  const SfwRobot &R = scene->robot;
  if (costs_out) {
    uint32_t i = 0;
    for (uint32_t a = 0; a < n_v; ++a)
      for (uint32_t b = 0; b < n_w; ++b, ++i) {
        if (linvels[a] == 0.0 && angvels[b] == 0.0) {
          costs_out[i] = -2.0;
          continue;
        }
        Trajectory t;
        costs_out[i] = pl.scoreTrajectory(R.x, R.y, R.theta, R.vx, R.vy, R.vtheta, linvels[a], 0.0,
                                          angvels[b], pl.params_.max_trans_acc_, 0.0,
                                          pl.params_.max_rot_acc_, R.wpx, R.wpy, rig->agents, t);
      }
  }
  if (best_out) {
    std::memset(best_out, 0, sizeof(*best_out));
    std::vector<geometry_msgs::msg::PoseStamped> plan;
    plan.push_back(pose_of(R.wpx, R.wpy, 0.0));
    double d2 = (R.x - R.wpx) * (R.x - R.wpx) + (R.y - R.wpy) * (R.y - R.wpy);
    if (d2 < 1.5 * 1.5 + 1e-9) // keep findBestAction out of its approach / goal branches
      plan.push_back(pose_of(R.wpx + 100.0, R.wpy, 0.0));
    pl.updatePlan(plan);
    geometry_msgs::msg::Twist vel, cmd;
    vel.linear.x = R.vx;
    vel.linear.y = R.vy;
    vel.angular.z = R.vtheta;
    bool ok = pl.findBestAction(pose_of(R.x, R.y, R.theta), vel, cmd);
    best_out->valid = ok ? 1 : 0;
    if (ok) {
      best_out->v = cmd.linear.x;
      best_out->w = cmd.angular.z;
      // index of the marker the reference painted green (:435-441) = best_i
      uint32_t bi = 0;
      for (uint32_t i = 0; i < pl.markers_.markers.size(); ++i)
        if (pl.markers_.markers[i].color.g == 1.0f && pl.markers_.markers[i].color.a == 1.0f)
          bi = i;
      best_out->index = bi;
      best_out->cost = costs_out ? (float)costs_out[bi] : 0.0f;
    }
  }
  return SFW_OK;
}

// The MarkerArray the reference's findBestAction leaves behind after one grid tick (:345-417,435-441):
// per sample its colour (r,g,b,a), number of points and the points (x,y,z).  Same rig as sfw_ref_score.
int sfw_ref_markers(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                    const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w, float *rgba_out,
                    uint32_t *npts_out, double *xyz_out, uint32_t max_pts) {
  auto rig = make_rig(*params, sfm, *scene);
  SFWPlanner &pl = *rig->planner;
  pl.params_.get(rig->node.get(), kName);
  pl.linvels_.assign(linvels, linvels + n_v);
  pl.angvels_.assign(angvels, angvels + n_w);
  pl.initializeMarkers();
  const SfwRobot &R = scene->robot;
  std::vector<geometry_msgs::msg::PoseStamped> plan;
  plan.push_back(pose_of(R.wpx, R.wpy, 0.0));
  double d2 = (R.x - R.wpx) * (R.x - R.wpx) + (R.y - R.wpy) * (R.y - R.wpy);
  if (d2 < 1.5 * 1.5 + 1e-9)
    plan.push_back(pose_of(R.wpx + 100.0, R.wpy, 0.0));
  pl.updatePlan(plan);
  geometry_msgs::msg::Twist vel, cmd;
  vel.linear.x = R.vx;
  vel.linear.y = R.vy;
  vel.angular.z = R.vtheta;
  bool ok = pl.findBestAction(pose_of(R.x, R.y, R.theta), vel, cmd);
  const auto &mk = pl.getMarkers().markers;
  for (uint32_t i = 0; i < mk.size() && i < n_v * n_w; ++i) {
    rgba_out[4 * i] = mk[i].color.r;
    rgba_out[4 * i + 1] = mk[i].color.g;
    rgba_out[4 * i + 2] = mk[i].color.b;
    rgba_out[4 * i + 3] = mk[i].color.a;
    npts_out[i] = (uint32_t)mk[i].points.size();
    for (uint32_t k = 0; k < mk[i].points.size() && k < max_pts; ++k) {
      double *o = xyz_out + ((size_t)i * max_pts + k) * 3;
      o[0] = mk[i].points[k].x;
      o[1] = mk[i].points[k].y;
      o[2] = mk[i].points[k].z;
    }
  }
  return ok ? 1 : 0;
}

// SFWPlanner::mayIStop of the reference (:718-765; private, unreachable upstream).  Returns 1 / 0.
int sfw_ref_may_i_stop(const SfwParams *params, const SfwScene *scene, double vl_x, double vl_y, double va, double x,
                       double y, double th, double dt) {
  SfwScene sc = *scene;
  sc.n_peds = 0;
  sc.n_obstacles = 0;
  auto rig = make_rig(*params, nullptr, sc);
  rig->planner->params_.get(rig->node.get(), kName);
  return rig->planner->mayIStop(vl_x, vl_y, va, x, y, th, dt) ? 1 : 0;
}

// WorldModel::footprintCost(x,y,theta,spec) of the reference via SFWPlanner::footprintCost (:709).
double sfw_ref_footprint_cost(const SfwScene *scene, double x, double y, double theta) {
  SfwParams p;
  std::memset(&p, 0, sizeof(p));
  p.max_vel_x = 0.7; p.max_trans_acc = 1.0; p.max_rot_acc = 1.0; p.sim_time = 1.0;
  p.sim_granularity = 0.025; p.robot_radius = 0.35f; p.social_weight = 1.2; p.costmap_weight = 2.0;
  p.angle_weight = 0.7; p.distance_weight = 1.0; p.vel_weight = 1.0;
  SfwScene sc = *scene;
  sc.n_peds = 0;
  sc.n_obstacles = 0;
  auto rig = make_rig(p, nullptr, sc);
  return rig->planner->footprintCost(x, y, theta);
}

// The reference's findBestAction on an explicit plan (host-logic parity: waypoint selection,
// goal / approach branches, src/sfw_planner.cpp:117-469).  plan_xyt: n_plan (x,y,yaw) triples.
// linvels/angvels may be NULL to keep the reference's own 5x9 sample sets (:65-85).
// ext: {min_vel_x, max_vel_th, min_vel_th, min_in_place_vel_th, yaw_goal_tolerance,
//       xy_goal_tolerance, wp_tolerance, is_circular}
int sfw_ref_find_best_action(const SfwParams *params, const double *ext, const SfwSfmParams *sfm,
                             const SfwScene *scene, const double *plan_xyt, uint32_t n_plan,
                             const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                             double *cmd_vxvyvt, int *wp_index_out, int *running_out) {
  auto rig = make_rig(*params, sfm, *scene);
  std::string b = std::string(kName) + ".";
  if (ext) {
    rig->node->set_parameter(b + "min_trans_vel", rclcpp::ParameterValue(ext[0]));
    rig->node->set_parameter(b + "max_rot_vel", rclcpp::ParameterValue(ext[1]));
    rig->node->set_parameter(b + "min_rot_vel", rclcpp::ParameterValue(ext[2]));
    rig->node->set_parameter(b + "min_in_place_rot_vel", rclcpp::ParameterValue(ext[3]));
    rig->node->set_parameter(b + "yaw_goal_tolerance", rclcpp::ParameterValue(ext[4]));
    rig->node->set_parameter(b + "xy_goal_tolerance", rclcpp::ParameterValue(ext[5]));
    rig->node->set_parameter(b + "wp_tolerance", rclcpp::ParameterValue(ext[6]));
    rig->node->set_parameter(b + "is_circular", rclcpp::ParameterValue(ext[7] != 0.0));
  }
  // rebuild so the constructor derives its sample sets from the final parameters
  std::vector<geometry_msgs::msg::Point> fp = rig->planner->getFootprint();
  rig->planner.reset(new SFWPlanner(rig->node, kName, rig->iface, *rig->costmap, fp));
  SFWPlanner &pl = *rig->planner;
  if (linvels && angvels) {
    pl.linvels_.assign(linvels, linvels + n_v);
    pl.angvels_.assign(angvels, angvels + n_w);
    pl.initializeMarkers();
  }
  std::vector<geometry_msgs::msg::PoseStamped> plan;
  for (uint32_t i = 0; i < n_plan; ++i)
    plan.push_back(pose_of(plan_xyt[3 * i], plan_xyt[3 * i + 1], plan_xyt[3 * i + 2]));
  pl.updatePlan(plan);
  const SfwRobot &R = scene->robot;
  geometry_msgs::msg::Twist vel, cmd;
  vel.linear.x = R.vx;
  vel.linear.y = R.vy;
  vel.angular.z = R.vtheta;
  bool ok = pl.findBestAction(pose_of(R.x, R.y, R.theta), vel, cmd);
  cmd_vxvyvt[0] = cmd.linear.x;
  cmd_vxvyvt[1] = cmd.linear.y;
  cmd_vxvyvt[2] = cmd.angular.z;
  if (wp_index_out)
    *wp_index_out = pl.wp_index_;
  if (running_out)
    *running_out = pl.running_ ? 1 : 0;
  return ok ? 1 : 0;
}

// The reference's default sample sets (src/sfw_planner.cpp:65-85) for given max velocities.
int sfw_ref_default_samples(double max_vel_x, double max_vel_th, double *linvels5, double *angvels9) {
  SfwParams p;
  std::memset(&p, 0, sizeof(p));
  p.max_vel_x = max_vel_x; p.max_trans_acc = 1.0; p.max_rot_acc = 1.0; p.sim_time = 1.0;
  p.sim_granularity = 0.025; p.robot_radius = 0.35f;
  static const unsigned char cell = 0;
  static const double fpz[2] = {0, 0};
  SfwScene sc;
  std::memset(&sc, 0, sizeof(sc));
  sc.costmap = &cell; sc.size_x = 1; sc.size_y = 1; sc.resolution = 1.0; sc.footprint_xy = fpz;
  auto rig = make_rig(p, nullptr, sc);
  rig->node->set_parameter(std::string(kName) + ".max_rot_vel", rclcpp::ParameterValue(max_vel_th));
  rig->planner.reset(new SFWPlanner(rig->node, kName, rig->iface, *rig->costmap, {}));
  for (int i = 0; i < 5; ++i) linvels5[i] = rig->planner->linvels_[i];
  for (int i = 0; i < 9; ++i) angvels9[i] = rig->planner->angvels_[i];
  return 0;
}

} // extern "C"
