// plugin/src/sfw_planner.cpp — social_force_window_planner::SFWPlanner over the B200 scorer.
// See plugin/include/social_force_window_planner/sfw_planner.hpp for what this replaces and what it keeps.
#include <social_force_window_planner/sfw_planner.hpp>

#include <cstring>
#include <stdexcept>

namespace social_force_window_planner {

namespace {

template <typename T>
void read_param(rclcpp_lifecycle::LifecycleNode *node, const std::string &key, const rclcpp::ParameterValue &fallback,
                T &out) {
  nav2_util::declare_parameter_if_not_declared(node, key, fallback);
  node->get_parameter(key, out);
}

void set_twist(geometry_msgs::msg::Twist &t, double vx, double vy, double vth) {
  t.linear.x = vx;
  t.linear.y = vy;
  t.linear.z = 0.0;
  t.angular.x = 0.0;
  t.angular.y = 0.0;
  t.angular.z = vth;
}

void paint(visualization_msgs::msg::Marker &m, float r, float g, float b, float a) {
  m.color.r = r;
  m.color.g = g;
  m.color.b = b;
  m.color.a = a;
}

} // namespace

// Parameter names / defaults: reference sfw_planner.hpp:77-187 (config/local_planner.yaml uses these keys).
void ControllerParams::get(rclcpp_lifecycle::LifecycleNode *node, const std::string &name) {
  const std::string p = name + ".";
  read_param(node, p + "controller_frame", rclcpp::ParameterValue("odom"), controller_frame_);
  read_param(node, p + "robot_base_frame", rclcpp::ParameterValue("base_link"), robot_base_frame_);
  struct {
    const char *key;
    double fallback;
    double *value;
  } const doubles[] = {
      {"max_trans_vel", 0.7, &max_vel_x_},        {"min_trans_vel", 0.1, &min_vel_x_},
      {"max_rot_vel", 0.5, &max_vel_th_},         {"min_rot_vel", 0.1, &min_vel_th_},
      {"max_trans_acc", 1.0, &max_trans_acc_},    {"max_rot_acc", 1.0, &max_rot_acc_},
      {"min_in_place_rot_vel", 0.3, &min_in_place_vel_th_},
      {"yaw_goal_tolerance", 0.05, &yaw_goal_tolerance_},
      {"xy_goal_tolerance", 0.10, &xy_goal_tolerance_},
      {"wp_tolerance", 0.5, &wp_tolerance_},      {"sim_time", 1.0, &sim_time_},
      {"sim_granularity", 0.025, &sim_granularity_},
  };
  for (const auto &d : doubles)
    read_param(node, p + d.key, rclcpp::ParameterValue(d.fallback), *d.value);
  // defaults to whatever sim_granularity was just set to (:121-125)
  read_param(node, p + "angular_sim_granularity", rclcpp::ParameterValue(sim_granularity_), angular_sim_granularity_);
  read_param(node, p + "robot_radius", rclcpp::ParameterValue(0.35), robot_radius_);
  read_param(node, p + "people_radius", rclcpp::ParameterValue(0.35), people_radius_);
  read_param(node, p + "is_circular", rclcpp::ParameterValue(true), is_circular_);
  read_param(node, p + "sfm_goal_weight", rclcpp::ParameterValue(2.0), sfm_goal_weight_);
  read_param(node, p + "sfm_obstacle_weight", rclcpp::ParameterValue(20.0), sfm_obstacle_weight_);
  read_param(node, p + "sfm_people_weight", rclcpp::ParameterValue(12.0), sfm_people_weight_);
  read_param(node, p + "social_weight", rclcpp::ParameterValue(1.2), social_weight_);
  read_param(node, p + "costmap_weight", rclcpp::ParameterValue(2.0), costmap_weight_);
  read_param(node, p + "angle_weight", rclcpp::ParameterValue(0.7), angle_weight_);
  read_param(node, p + "distance_weight", rclcpp::ParameterValue(1.0), distance_weight_);
  read_param(node, p + "velocity_weight", rclcpp::ParameterValue(1.0), vel_weight_);
  read_param(node, p + "cuda_device", rclcpp::ParameterValue(0.0), cuda_device_);
}

SFWPlanner::SFWPlanner(const rclcpp_lifecycle::LifecycleNode::SharedPtr &parent, const std::string name,
                       std::shared_ptr<SFMSensorInterface> &sensor_iface, const nav2_costmap_2d::Costmap2D &costmap,
                       std::vector<geometry_msgs::msg::Point> footprint_spec)
    : node_(parent), name_(name), sensor_iface_(sensor_iface), costmap_(costmap), footprint_spec_(footprint_spec) {
  params_.get(node_.get(), name_);
  sfw_host::default_sample_sets(params_.max_vel_x_, params_.max_vel_th_, linvels_, angvels_);
  initializeMarkers();
  std::memset(&best_, 0, sizeof(best_));
  if (sfw_create(&ctx_, (int)params_.cuda_device_, nullptr, nullptr) != SFW_OK) {
    error_ = sfw_last_error(nullptr);
    RCLCPP_ERROR(node_->get_logger(), "SFWPlanner: no usable CUDA device (%s)", error_.c_str());
    throw std::runtime_error("social_force_window_planner (B200): " + error_);
  }
  RCLCPP_INFO(node_->get_logger(), "SFWPlanner: %zu x %zu velocity samples scored on CUDA device %d", linvels_.size(),
              angvels_.size(), (int)params_.cuda_device_);
}

SFWPlanner::~SFWPlanner() {
  if (ctx_)
    sfw_destroy(ctx_);
}

const char *SFWPlanner::lastKernel() const { return ctx_ ? sfw_last_kernel(ctx_) : "none"; }

void SFWPlanner::setSampleSets(const std::vector<double> &linvels, const std::vector<double> &angvels) {
  std::lock_guard<std::mutex> lock(configuration_mutex_);
  linvels_ = linvels;
  angvels_ = angvels;
  initializeMarkers();
}

// One LINE_STRIP marker per sample, ids in sample order (reference src/sfw_planner.cpp:91-110)
void SFWPlanner::initializeMarkers() {
  markers_.markers.assign(linvels_.size() * angvels_.size(), visualization_msgs::msg::Marker());
  int id = 0;
  for (auto &m : markers_.markers) {
    m.header.frame_id = params_.controller_frame_;
    m.ns = "trajectories";
    m.id = id++;
    m.type = 4;   // LINE_STRIP
    m.action = 0; // add / modify
    m.lifetime = rclcpp::Duration(0.3);
    m.scale.x = 0.01;
    m.color.a = 1.0;
    m.pose.orientation.w = 1.0;
  }
  pending_markers_ = Pending::None;
}

bool SFWPlanner::updatePlan(const std::vector<geometry_msgs::msg::PoseStamped> &new_plan) {
  std::vector<sfw_host::PlanPose> plan(new_plan.size());
  for (size_t i = 0; i < new_plan.size(); ++i) {
    plan[i].x = new_plan[i].pose.position.x;
    plan[i].y = new_plan[i].pose.position.y;
    plan[i].yaw = tf2::getYaw(new_plan[i].pose.orientation);
  }
  tracker_.setPlan(plan);
  if (plan.empty())
    RCLCPP_WARN(node_->get_logger(), "New local plan size = 0!");
  return true;
}

bool SFWPlanner::isGoalReached() { return tracker_.consumeGoalFlag(); }
void SFWPlanner::resetGoal() { tracker_.clearGoalFlag(); }

// Scene packer: the sensor interface's agent snapshot (agents[0] = robot, reference src/sensor_interface.cpp:
// 618-631) + the live costmap + the footprint -> one SfwScene, scored in one call.
bool SFWPlanner::score(float rx, float ry, float rt, float rvx, float rvy, float rvt, double wpx, double wpy,
                       const std::vector<sfm::Agent> &agents, const double *lin, uint32_t n_v, const double *ang,
                       uint32_t n_w) {
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
  SfwRobot &r = sc.robot;
  r.x = rx, r.y = ry, r.theta = rt;
  r.vx = rvx, r.vy = rvy, r.vtheta = rvt;
  r.wpx = wpx, r.wpy = wpy;
  peds_.clear();
  obstacles_.clear();
  if (!agents.empty()) {
    const sfm::Agent &me = agents.front();
    r.agent_x = me.position.getX(), r.agent_y = me.position.getY();
    r.agent_vx = me.velocity.getX(), r.agent_vy = me.velocity.getY();
    r.agent_radius = me.radius;
    for (const utils::Vector2d &o : me.obstacles1) { // every agent carries the same list (:513-524)
      obstacles_.push_back(o.getX());
      obstacles_.push_back(o.getY());
    }
    peds_.resize(agents.size() - 1);
    for (size_t j = 1; j < agents.size(); ++j) {
      const sfm::Agent &a = agents[j];
      SfwPed &q = peds_[j - 1];
      std::memset(&q, 0, sizeof(q));
      q.x = a.position.getX(), q.y = a.position.getY();
      q.vx = a.velocity.getX(), q.vy = a.velocity.getY();
      q.desired_velocity = a.desiredVelocity;
      q.radius = a.radius;
      q.group_id = a.groupId;
      q.id = a.id;
      q.has_goal = a.goals.empty() ? 0 : 1;
      if (q.has_goal) {
        q.goal_x = a.goals.front().center.getX(), q.goal_y = a.goals.front().center.getY();
        q.goal_radius = a.goals.front().radius;
      }
    }
  } else { // no sensor data yet: the robot alone at its pose
    r.agent_x = rx, r.agent_y = ry;
    r.agent_radius = params_.robot_radius_;
  }
  footprint_xy_.clear();
  for (const geometry_msgs::msg::Point &v : footprint_spec_) {
    footprint_xy_.push_back(v.x);
    footprint_xy_.push_back(v.y);
  }
  sc.costmap = costmap_.getCharMap();
  sc.size_x = costmap_.getSizeInCellsX();
  sc.size_y = costmap_.getSizeInCellsY();
  sc.resolution = costmap_.getResolution();
  sc.origin_x = costmap_.getOriginX();
  sc.origin_y = costmap_.getOriginY();
  sc.peds = peds_.empty() ? nullptr : peds_.data();
  sc.n_peds = (uint32_t)peds_.size();
  sc.obstacles_xy = obstacles_.empty() ? nullptr : obstacles_.data();
  sc.n_obstacles = (uint32_t)(obstacles_.size() / 2);
  sc.footprint_xy = footprint_xy_.empty() ? nullptr : footprint_xy_.data();
  sc.n_footprint = (uint32_t)(footprint_xy_.size() / 2);

  costs_.assign((size_t)n_v * n_w, 0.0f);
  // the reference skips (0,0) inside its sample loop only (:349); its single scoreTrajectory calls always evaluate
  sfw_set_zero_sample(ctx_, (uint64_t)n_v * n_w == 1 ? 1 : 0);
  if (sfw_score(ctx_, &p, nullptr, &sc, lin, n_v, ang, n_w, costs_.data(), &best_) != SFW_OK) {
    error_ = sfw_last_error(ctx_);
    RCLCPP_ERROR(node_->get_logger(), "SFWPlanner: scoring failed: %s", error_.c_str());
    return false;
  }
  return true;
}

bool SFWPlanner::findBestAction(const geometry_msgs::msg::PoseStamped &global_pose,
                                const geometry_msgs::msg::Twist &global_vel, geometry_msgs::msg::Twist &cmd_vel) {
  std::lock_guard<std::mutex> lock(configuration_mutex_);
  params_.get(node_.get(), name_); // parameters may change between ticks

  // the reference narrows the robot state to float before anything else (:145-152)
  const float rx = global_pose.pose.position.x, ry = global_pose.pose.position.y;
  const float rt = tf2::getYaw(global_pose.pose.orientation);
  const float rvx = global_vel.linear.x, rvy = global_vel.linear.y, rvt = global_vel.angular.z;

  sfw_host::TickLimits lim;
  lim.max_vel_x = params_.max_vel_x_, lim.min_vel_x = params_.min_vel_x_;
  lim.max_vel_th = params_.max_vel_th_, lim.min_vel_th = params_.min_vel_th_;
  lim.min_in_place_vel_th = params_.min_in_place_vel_th_;
  lim.yaw_goal_tolerance = params_.yaw_goal_tolerance_, lim.xy_goal_tolerance = params_.xy_goal_tolerance_;
  lim.wp_tolerance = params_.wp_tolerance_;
  lim.is_circular = params_.is_circular_;

  const sfw_host::TickPlan tick = tracker_.next(rx, ry, rt, lim);
  set_twist(cmd_vel, tick.vx, tick.vy, tick.vth);
  using sfw_host::TickKind;
  if (tick.kind == TickKind::Idle || tick.kind == TickKind::GoalReached) {
    if (tick.kind == TickKind::GoalReached)
      RCLCPP_INFO(node_->get_logger(), "GOAL REACHED!");
    return true;
  }
  // agents[0] is the robot as the sensor interface last saw it
  const std::vector<sfm::Agent> agents = sensor_iface_->getAgents();

  if (tick.kind == TickKind::TurnInPlace) {
    if (!tick.needs_scoring)
      return true;
    // a non-circular base sweeps area when it turns: the rotation has to be legal (:199-218; no waypoint)
    const double lin = tick.vx, ang = tick.vth;
    return score(rx, ry, rt, rvx, rvy, rvt, 0.0, 0.0, agents, &lin, 1, &ang, 1) && !(costs_[0] < 0.0f);
  }

  if (tick.kind == TickKind::Approach) {
    const double lin = tick.vx, ang = tick.vth;
    if (score(rx, ry, rt, rvx, rvy, rvt, tick.wpx, tick.wpy, agents, &lin, 1, &ang, 1) && costs_[0] >= 0.0f) {
      pending_markers_ = Pending::Approach;
      return true;
    }
    RCLCPP_INFO(node_->get_logger(), "approach command lv %.2f av %.2f is not legal: sampling", tick.vx, tick.vth);
  }

  // the (v, w) sample set: ONE launch, the arg-min with the reference's tie-breaks comes back with it
  const bool ok = score(rx, ry, rt, rvx, rvy, rvt, tick.wpx, tick.wpy, agents, linvels_.data(),
                        (uint32_t)linvels_.size(), angvels_.data(), (uint32_t)angvels_.size());
  pending_markers_ = ok ? Pending::Grid : Pending::None;
  if (ok && best_.valid) {
    set_twist(cmd_vel, best_.v, 0.0, best_.w);
    RCLCPP_INFO(node_->get_logger(), "BEST TRAJ FOUND -- lvel: %.2f, avel: %.2f, cost: %.3f\n", best_.v, best_.w,
                best_.cost);
    return true;
  }
  set_twist(cmd_vel, 0.0, 0.0, 0.0); // nothing legal: stop (:456-468)
  return false;
}

// The markers of the last tick are only built when somebody asks for them (the node publishes them every tick,
// src/sfw_planner_node.cpp:284-285): the rollouts' recorded points come from one sfw_marker_points launch.
visualization_msgs::msg::MarkerArray &SFWPlanner::getMarkers() {
  std::lock_guard<std::mutex> lock(configuration_mutex_);
  refreshMarkers();
  return markers_;
}

void SFWPlanner::refreshMarkers() {
  const Pending what = pending_markers_;
  pending_markers_ = Pending::None;
  if (what == Pending::None || markers_.markers.empty())
    return;
  const auto now = node_->get_clock()->now();
  if (what == Pending::Approach) {
    // The reference appends the approach rollout to marker 0 without clearing it first (its clearing loop runs
    // over copies, :306-309) and paints it green (:310-322).  Kept as it is: RViz shows what it always showed.
    uint32_t n = 0;
    std::vector<double> xyz(3 * 65536);
    if (sfw_trajectory_points(ctx_, 0, 0, xyz.data(), 65536, &n) != SFW_OK)
      return;
    visualization_msgs::msg::Marker &m = markers_.markers[0];
    for (uint32_t k = 0; k < n; ++k) {
      geometry_msgs::msg::Point q;
      q.x = xyz[3 * k], q.y = xyz[3 * k + 1], q.z = 0.0;
      m.points.push_back(q);
    }
    paint(m, 0.0f, 1.0f, 0.0f, 1.0f);
    return;
  }
  const size_t n = markers_.markers.size();
  if (costs_.size() != n)
    return;
  std::vector<uint16_t> npts(n);
  if (sfw_marker_points(ctx_, 0, 0, 1, (uint32_t)n, nullptr, 0, npts.data()) != SFW_OK)
    return;
  uint32_t longest = 1;
  for (uint16_t v : npts)
    longest = std::max<uint32_t>(longest, v);
  std::vector<double> xyz(3 * (size_t)longest * n);
  if (sfw_marker_points(ctx_, 0, 0, 1, (uint32_t)n, xyz.data(), longest, npts.data()) != SFW_OK)
    return;
  for (size_t i = 0; i < n; ++i) {
    visualization_msgs::msg::Marker &m = markers_.markers[i];
    m.header.stamp = now;
    m.points.clear();
    if (costs_[i] == SFW_COST_SKIPPED)
      continue; // the (0,0) sample keeps whatever colour it had (:349-352)
    m.points.resize(npts[i]);
    for (uint32_t k = 0; k < npts[i]; ++k) {
      const double *q = &xyz[((size_t)i * longest + k) * 3];
      m.points[k].x = q[0], m.points[k].y = q[1], m.points[k].z = 0.0;
    }
    if (costs_[i] < 0.0f)
      paint(m, 1.0f, 0.0f, 0.0f, 0.6f); // rejected
    else
      paint(m, 0.0f, 0.0f, 1.0f, 0.6f); // legal
  }
  if (best_.valid && best_.index < n) { // the chosen one: raised and green (:435-441)
    visualization_msgs::msg::Marker &m = markers_.markers[best_.index];
    for (auto &q : m.points)
      q.z = 0.1;
    paint(m, 0.0f, 1.0f, 0.0f, 1.0f);
  }
}

} // namespace social_force_window_planner
