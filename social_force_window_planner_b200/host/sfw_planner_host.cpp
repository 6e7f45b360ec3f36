// sfw_planner_host.cpp — see sfw_planner_host.hpp.  Host control flow of the reference's
// SFWPlanner::findBestAction around ONE call into the CUDA scorer per tick.
#include "sfw_planner_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace social_force_window_planner {

SFWPlanner::SFWPlanner(const ControllerParams &params, const CostmapView *costmap,
                       std::vector<Point2D> footprint_spec, int device)
    : params_(params), costmap_(costmap), footprint_spec_(std::move(footprint_spec)), device_(device) {
  sfw_host::default_sample_sets(params_.max_vel_x_, params_.max_vel_th_, linvels_, angvels_); // :65-85
}

SFWPlanner::~SFWPlanner() {
  if (ctx_)
    sfw_destroy(ctx_);
}

void SFWPlanner::setSampleSets(std::vector<double> linvels, std::vector<double> angvels) {
  linvels_ = std::move(linvels);
  angvels_ = std::move(angvels);
}

bool SFWPlanner::ensureContext() {
  if (ctx_)
    return true;
  const int rc = sfw_create(&ctx_, device_, nullptr, nullptr);
  if (rc != SFW_OK) {
    error_ = sfw_last_error(nullptr);
    ctx_ = nullptr;
    return false;
  }
  return true;
}

uint64_t SFWPlanner::kernelLaunches() const { return ctx_ ? sfw_kernel_launches(ctx_) : 0; }

// The scene packer: agents (sensor-interface layout) + live costmap + footprint -> SfwScene, then one
// sfw_score.  agents[0] is the robot; all agents share agents[0]'s obstacle list (App. E-11).
bool SFWPlanner::score(const float rx, const float ry, const float rt, const float rvx, const float rvy,
                       const float rvt, double wpx, double wpy, const double *lin, uint32_t n_v,
                       const double *ang, uint32_t n_w, std::vector<float> &costs, SfwBest &best) {
  if (!ensureContext())
    return false;
  SfwParams p;
  std::memset(&p, 0, sizeof(p));
  p.max_vel_x = params_.max_vel_x_;
  p.max_trans_acc = params_.max_trans_acc_;
  p.max_rot_acc = params_.max_rot_acc_;
  p.sim_time = params_.sim_time_;
  p.sim_granularity = params_.sim_granularity_;
  p.robot_radius = params_.robot_radius_;
  p.social_weight = params_.social_weight_;
  p.costmap_weight = params_.costmap_weight_;
  p.angle_weight = params_.angle_weight_;
  p.distance_weight = params_.distance_weight_;
  p.vel_weight = params_.vel_weight_;

  SfwScene sc;
  std::memset(&sc, 0, sizeof(sc));
  sc.robot.x = rx;
  sc.robot.y = ry;
  sc.robot.theta = rt;
  sc.robot.vx = rvx;
  sc.robot.vy = rvy;
  sc.robot.vtheta = rvt;
  sc.robot.wpx = wpx;
  sc.robot.wpy = wpy;
  std::vector<SfwPed> peds;
  std::vector<double> obs, fp;
  if (!agents_.empty()) {
    const Agent &r = agents_[0];
    sc.robot.agent_x = r.position.x;
    sc.robot.agent_y = r.position.y;
    sc.robot.agent_vx = r.velocity.x;
    sc.robot.agent_vy = r.velocity.y;
    sc.robot.agent_radius = r.radius;
    for (const Point2D &o : r.obstacles1) {
      obs.push_back(o.x);
      obs.push_back(o.y);
    }
    peds.resize(agents_.size() - 1);
    for (size_t j = 1; j < agents_.size(); ++j) {
      const Agent &a = agents_[j];
      SfwPed &q = peds[j - 1];
      std::memset(&q, 0, sizeof(q));
      q.x = a.position.x;
      q.y = a.position.y;
      q.vx = a.velocity.x;
      q.vy = a.velocity.y;
      q.goal_x = a.goal_center.x;
      q.goal_y = a.goal_center.y;
      q.goal_radius = a.goal_radius;
      q.desired_velocity = a.desiredVelocity;
      q.radius = a.radius;
      q.has_goal = a.has_goal ? 1 : 0;
      q.group_id = a.groupId;
      q.id = a.id;
    }
  } else { // no sensor data yet: the robot alone at its pose
    sc.robot.agent_x = rx;
    sc.robot.agent_y = ry;
    sc.robot.agent_radius = params_.robot_radius_;
  }
  for (const Point2D &v : footprint_spec_) {
    fp.push_back(v.x);
    fp.push_back(v.y);
  }
  sc.costmap = costmap_->data;
  sc.size_x = costmap_->size_x;
  sc.size_y = costmap_->size_y;
  sc.resolution = costmap_->resolution;
  sc.origin_x = costmap_->origin_x;
  sc.origin_y = costmap_->origin_y;
  sc.peds = peds.empty() ? nullptr : peds.data();
  sc.n_peds = (uint32_t)peds.size();
  sc.obstacles_xy = obs.empty() ? nullptr : obs.data();
  sc.n_obstacles = (uint32_t)(obs.size() / 2);
  sc.footprint_xy = fp.empty() ? nullptr : fp.data();
  sc.n_footprint = (uint32_t)(fp.size() / 2);

  costs.assign((size_t)n_v * n_w, 0.0f);
  // the reference skips (0,0) in its grid loop only; its single scoreTrajectory calls always evaluate
  sfw_set_zero_sample(ctx_, (uint64_t)n_v * n_w == 1 ? 1 : 0);
  const int rc = sfw_score(ctx_, &p, nullptr, &sc, lin, n_v, ang, n_w, costs.data(), &best);
  if (rc != SFW_OK) {
    error_ = sfw_last_error(ctx_);
    staged_ = false;
    return false;
  }
  staged_ = true;
  return true;
}

std::vector<Point2D> SFWPlanner::trajectoryPoints(uint32_t sample_index) {
  std::vector<Point2D> out;
  if (!ctx_ || !staged_)
    return out;
  uint32_t n = 0;
  std::vector<double> xyz(3 * 65536);
  if (sfw_trajectory_points(ctx_, 0, sample_index, xyz.data(), 65536, &n) != SFW_OK)
    return out;
  out.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    out[i].x = xyz[3 * i];
    out[i].y = xyz[3 * i + 1];
  }
  return out;
}

bool SFWPlanner::mayIStop(double vl_x, double vl_y, double va, double x, double y, double th, double dt) {
  if (!ctx_ || !staged_) {
    error_ = "mayIStop: no scored tick yet";
    return false;
  }
  int32_t ok = 0;
  if (sfw_may_i_stop(ctx_, 0, vl_x, vl_y, va, x, y, th, dt, &ok, nullptr) != SFW_OK) {
    error_ = sfw_last_error(ctx_);
    return false;
  }
  return ok != 0;
}

std::vector<Marker> SFWPlanner::getMarkerArray() {
  std::vector<Marker> out;
  const size_t n = linvels_.size() * angvels_.size();
  if (!ctx_ || !staged_ || costs_.size() != n)
    return out; // no grid tick yet (the approach / rotate branches leave the markers alone)
  out.resize(n);
  std::vector<uint16_t> npts(n);
  uint32_t max_pts = 1;
  // first pass: counts only; second: the points
  if (sfw_marker_points(ctx_, 0, 0, 1, (uint32_t)n, nullptr, 0, npts.data()) != SFW_OK) {
    error_ = sfw_last_error(ctx_);
    out.clear();
    return out;
  }
  for (uint16_t v : npts)
    max_pts = std::max<uint32_t>(max_pts, v);
  std::vector<double> xyz(3 * (size_t)max_pts * n);
  if (sfw_marker_points(ctx_, 0, 0, 1, (uint32_t)n, xyz.data(), max_pts, npts.data()) != SFW_OK) {
    error_ = sfw_last_error(ctx_);
    out.clear();
    return out;
  }
  for (size_t i = 0; i < n; ++i) {
    Marker &m = out[i];
    m.id = (int)i;
    if (costs_[i] == SFW_COST_SKIPPED)
      continue; // :349-352
    m.points_xyz.resize(3 * (size_t)npts[i]);
    for (uint32_t k = 0; k < npts[i]; ++k) {
      m.points_xyz[3 * k] = xyz[((size_t)i * max_pts + k) * 3];
      m.points_xyz[3 * k + 1] = xyz[((size_t)i * max_pts + k) * 3 + 1];
      m.points_xyz[3 * k + 2] = 0.0;
    }
    if (costs_[i] < 0.0f) { // :376-386
      m.r = 1.0f, m.g = 0.0f, m.b = 0.0f, m.a = 0.6f;
    } else {
      m.r = 0.0f, m.g = 0.0f, m.b = 1.0f, m.a = 0.6f;
    }
  }
  if (best_i_ >= 0 && (size_t)best_i_ < n) { // :435-441
    Marker &m = out[best_i_];
    for (size_t k = 0; k < m.points_xyz.size() / 3; ++k)
      m.points_xyz[3 * k + 2] = 0.1;
    m.r = 0.0f, m.g = 1.0f, m.b = 0.0f, m.a = 1.0f;
  }
  return out;
}

// One control tick (reference src/sfw_planner.cpp:117-469).  What kind of tick this is — idle, goal reached, turn
// in place, approach, sample grid — and the command / waypoint that go with it is decided by sfw_host::PlanTracker
// (sfw_tick.hpp, shared with the ROS-typed plugin class in plugin/src/sfw_planner.cpp); this function only talks to
// the scorer: one sfw_score call for a proposal that has to be legal, one for the grid.
bool SFWPlanner::findBestAction(const Pose2D &global_pose, const Twist2D &global_vel, Twist2D &cmd_vel) {
  best_i_ = -1;
  // the reference narrows the robot state to float first (:145-152)
  const float rx = global_pose.x, ry = global_pose.y, rt = global_pose.yaw;
  const float rvx = global_vel.linear_x, rvy = global_vel.linear_y, rvt = global_vel.angular_z;

  sfw_host::TickLimits lim;
  lim.max_vel_x = params_.max_vel_x_, lim.min_vel_x = params_.min_vel_x_;
  lim.max_vel_th = params_.max_vel_th_, lim.min_vel_th = params_.min_vel_th_;
  lim.min_in_place_vel_th = params_.min_in_place_vel_th_;
  lim.yaw_goal_tolerance = params_.yaw_goal_tolerance_, lim.xy_goal_tolerance = params_.xy_goal_tolerance_;
  lim.wp_tolerance = params_.wp_tolerance_;
  lim.is_circular = params_.is_circular_;

  using sfw_host::TickKind;
  const sfw_host::TickPlan tick = tracker_.next(rx, ry, rt, lim);
  cmd_vel.linear_x = tick.vx;
  cmd_vel.linear_y = tick.vy;
  cmd_vel.angular_z = tick.vth;
  if (tick.kind == TickKind::Idle || tick.kind == TickKind::GoalReached)
    return true;

  SfwBest best;
  if (tick.kind == TickKind::TurnInPlace) {
    if (!tick.needs_scoring)
      return true;
    const double lin = tick.vx, ang = tick.vth; // a non-circular base must be able to sweep the turn (:199-218)
    return score(rx, ry, rt, rvx, rvy, rvt, 0.0, 0.0, &lin, 1, &ang, 1, costs_, best) && !(costs_[0] < 0.0f);
  }
  if (tick.kind == TickKind::Approach) {
    const double lin = tick.vx, ang = tick.vth;
    if (score(rx, ry, rt, rvx, rvy, rvt, tick.wpx, tick.wpy, &lin, 1, &ang, 1, costs_, best) && costs_[0] >= 0.0f) {
      markers_.assign(1, SampleMarker{tick.vx, tick.vth, costs_[0]});
      best_i_ = 0;
      return true;
    }
  }

  // the (v, w) grid (:338-417): ONE launch; the arg-min with the reference's tie-breaks comes back with it
  cmd_vel.linear_x = cmd_vel.linear_y = cmd_vel.angular_z = 0.0;
  if (!score(rx, ry, rt, rvx, rvy, rvt, tick.wpx, tick.wpy, linvels_.data(), (uint32_t)linvels_.size(), angvels_.data(),
             (uint32_t)angvels_.size(), costs_, best))
    return false;
  markers_.resize(costs_.size());
  for (size_t i = 0; i < costs_.size(); ++i)
    markers_[i] = SampleMarker{linvels_[i / angvels_.size()], angvels_[i % angvels_.size()], costs_[i]};
  if (!best.valid)
    return false; // nothing legal: zero twist (:456-468)
  best_i_ = (int)best.index;
  cmd_vel.linear_x = best.v;
  cmd_vel.angular_z = best.w;
  return true;
}

bool SFWPlanner::updatePlan(const std::vector<Pose2D> &new_plan) {
  std::vector<sfw_host::PlanPose> plan(new_plan.size());
  for (size_t i = 0; i < new_plan.size(); ++i)
    plan[i] = sfw_host::PlanPose{new_plan[i].x, new_plan[i].y, new_plan[i].yaw};
  tracker_.setPlan(plan);
  return true;
}

bool SFWPlanner::isGoalReached() { return tracker_.consumeGoalFlag(); }

} // namespace social_force_window_planner

// ---------------------------------------------------------------------------------------------------
// C wrapper so the tests (ctypes) can drive the class the way sfw_planner_node.cpp does.
// ---------------------------------------------------------------------------------------------------
using social_force_window_planner::Agent;
using social_force_window_planner::ControllerParams;
using social_force_window_planner::CostmapView;
using social_force_window_planner::Point2D;
using social_force_window_planner::Pose2D;
using social_force_window_planner::SFWPlanner;
using social_force_window_planner::Twist2D;

struct sfwh_planner {
  CostmapView view;
  SFWPlanner *pl;
};

extern "C" {

// ext: {min_vel_x, max_vel_th, min_vel_th, min_in_place_vel_th, yaw_goal_tolerance, xy_goal_tolerance,
//       wp_tolerance, is_circular} (same order as oracle/ref_harness.cpp's sfw_ref_find_best_action)
sfwh_planner *sfwh_create(const SfwParams *p, const double *ext, const SfwScene *scene, int device) {
  ControllerParams cp;
  cp.max_vel_x_ = p->max_vel_x;
  cp.max_trans_acc_ = p->max_trans_acc;
  cp.max_rot_acc_ = p->max_rot_acc;
  cp.sim_time_ = p->sim_time;
  cp.sim_granularity_ = p->sim_granularity;
  cp.robot_radius_ = p->robot_radius;
  cp.social_weight_ = p->social_weight;
  cp.costmap_weight_ = p->costmap_weight;
  cp.angle_weight_ = p->angle_weight;
  cp.distance_weight_ = p->distance_weight;
  cp.vel_weight_ = p->vel_weight;
  if (ext) {
    cp.min_vel_x_ = ext[0];
    cp.max_vel_th_ = ext[1];
    cp.min_vel_th_ = ext[2];
    cp.min_in_place_vel_th_ = ext[3];
    cp.yaw_goal_tolerance_ = ext[4];
    cp.xy_goal_tolerance_ = ext[5];
    cp.wp_tolerance_ = ext[6];
    cp.is_circular_ = ext[7] != 0.0;
  }
  sfwh_planner *h = new sfwh_planner();
  h->view.data = scene->costmap;
  h->view.size_x = scene->size_x;
  h->view.size_y = scene->size_y;
  h->view.resolution = scene->resolution;
  h->view.origin_x = scene->origin_x;
  h->view.origin_y = scene->origin_y;
  std::vector<Point2D> fp(scene->n_footprint);
  for (uint32_t i = 0; i < scene->n_footprint; ++i) {
    fp[i].x = scene->footprint_xy[2 * i];
    fp[i].y = scene->footprint_xy[2 * i + 1];
  }
  h->pl = new SFWPlanner(cp, &h->view, fp, device);
  // agents exactly as the sensor interface would hold them for this scene
  std::vector<Agent> ag(scene->n_peds + 1);
  std::vector<Point2D> obs(scene->n_obstacles);
  for (uint32_t i = 0; i < scene->n_obstacles; ++i) {
    obs[i].x = scene->obstacles_xy[2 * i];
    obs[i].y = scene->obstacles_xy[2 * i + 1];
  }
  ag[0].id = -1;
  ag[0].position = {scene->robot.agent_x, scene->robot.agent_y};
  ag[0].velocity = {scene->robot.agent_vx, scene->robot.agent_vy};
  ag[0].radius = scene->robot.agent_radius;
  ag[0].obstacles1 = obs;
  for (uint32_t j = 0; j < scene->n_peds; ++j) {
    const SfwPed &q = scene->peds[j];
    Agent &a = ag[j + 1];
    a.id = q.id;
    a.groupId = q.group_id;
    a.position = {q.x, q.y};
    a.velocity = {q.vx, q.vy};
    a.radius = q.radius;
    a.desiredVelocity = q.desired_velocity;
    a.has_goal = q.has_goal != 0;
    a.goal_center = {q.goal_x, q.goal_y};
    a.goal_radius = q.goal_radius;
    a.obstacles1 = obs;
  }
  h->pl->setAgents(std::move(ag));
  return h;
}

void sfwh_destroy(sfwh_planner *h) {
  if (!h)
    return;
  delete h->pl;
  delete h;
}

void sfwh_set_samples(sfwh_planner *h, const double *lin, uint32_t n_v, const double *ang, uint32_t n_w) {
  h->pl->setSampleSets(std::vector<double>(lin, lin + n_v), std::vector<double>(ang, ang + n_w));
}

void sfwh_get_samples(sfwh_planner *h, double *lin5, double *ang9) {
  for (size_t i = 0; i < h->pl->linvels().size() && i < 5; ++i)
    lin5[i] = h->pl->linvels()[i];
  for (size_t i = 0; i < h->pl->angvels().size() && i < 9; ++i)
    ang9[i] = h->pl->angvels()[i];
}

void sfwh_update_plan(sfwh_planner *h, const double *plan_xyt, uint32_t n) {
  std::vector<Pose2D> plan(n);
  for (uint32_t i = 0; i < n; ++i)
    plan[i] = Pose2D{plan_xyt[3 * i], plan_xyt[3 * i + 1], plan_xyt[3 * i + 2]};
  h->pl->updatePlan(plan);
}

// returns findBestAction's bool; cmd = {linear.x, linear.y, angular.z}
int sfwh_find_best_action(sfwh_planner *h, const double *pose_xyt, const double *vel_xyt, double *cmd,
                          int *wp_index, int *running, int *best_index, uint64_t *launches) {
  Twist2D out;
  const bool ok = h->pl->findBestAction(Pose2D{pose_xyt[0], pose_xyt[1], pose_xyt[2]},
                                        Twist2D{vel_xyt[0], vel_xyt[1], vel_xyt[2]}, out);
  cmd[0] = out.linear_x;
  cmd[1] = out.linear_y;
  cmd[2] = out.angular_z;
  if (wp_index)
    *wp_index = h->pl->wpIndex();
  if (running)
    *running = h->pl->running() ? 1 : 0;
  if (best_index)
    *best_index = h->pl->bestIndex();
  if (launches)
    *launches = h->pl->kernelLaunches();
  return ok ? 1 : 0;
}

// MarkerArray of the last grid tick: rgba_out[n*4], npts_out[n], xyz_out[n*max_pts*3]; returns n or -1
int sfwh_get_markers(sfwh_planner *h, float *rgba_out, uint32_t *npts_out, double *xyz_out, uint32_t max_pts) {
  std::vector<social_force_window_planner::Marker> mk = h->pl->getMarkerArray();
  if (mk.empty())
    return -1;
  for (size_t i = 0; i < mk.size(); ++i) {
    rgba_out[4 * i] = mk[i].r;
    rgba_out[4 * i + 1] = mk[i].g;
    rgba_out[4 * i + 2] = mk[i].b;
    rgba_out[4 * i + 3] = mk[i].a;
    const size_t np = mk[i].points_xyz.size() / 3;
    npts_out[i] = (uint32_t)np;
    for (size_t k = 0; k < np && k < max_pts; ++k)
      for (int c = 0; c < 3; ++c)
        xyz_out[(i * max_pts + k) * 3 + c] = mk[i].points_xyz[3 * k + c];
  }
  return (int)mk.size();
}

int sfwh_is_goal_reached(sfwh_planner *h) { return h->pl->isGoalReached() ? 1 : 0; }
const char *sfwh_last_error(sfwh_planner *h) { return h->pl->lastError().c_str(); }

} // extern "C"
