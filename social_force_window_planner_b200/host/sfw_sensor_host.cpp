// sfw_sensor_host.cpp — see sfw_sensor_host.hpp.  Statement order follows the reference's callbacks
// (src/sensor_interface.cpp) so that the agent snapshot is the same, field for field.
#include "sfw_sensor_host.hpp"

#include <cmath>
#include <cstring>

namespace social_force_window_planner {

SFMSensorInterface::SFMSensorInterface(const InterfaceParams &params, int device)
    : iface_params_(params), device_(device) {
  // the agent list starts with the robot alone (reference :31-37)
  agents_.resize(1);
  agents_[0].desiredVelocity = iface_params_.max_robot_vel_x_;
  agents_[0].radius = iface_params_.robot_radius_;
  agents_[0].teleoperated = true;
  agents_[0].groupId = -1;
}

SFMSensorInterface::~SFMSensorInterface() {
  if (ctx_)
    sfw_destroy(ctx_);
}

bool SFMSensorInterface::ensureContext() {
  if (ctx_)
    return true;
  if (sfw_create(&ctx_, device_, nullptr, nullptr) != SFW_OK) {
    error_ = sfw_last_error(nullptr);
    ctx_ = nullptr;
    return false;
  }
  return true;
}

uint64_t SFMSensorInterface::kernelLaunches() const { return ctx_ ? sfw_kernel_launches(ctx_) : 0; }

bool SFMSensorInterface::lookup(const std::string &frame, PlanarTransform &out) const {
  auto it = tf_.find(frame);
  if (it == tf_.end())
    return false;
  out = it->second;
  return true;
}

void SFMSensorInterface::laserCb(const LaserScanMsg &laser) {
  if (!running_ || !odom_received_)
    return;
  laser_received_ = true;

  SfwLaserScan scan;
  std::memset(&scan, 0, sizeof(scan));
  scan.ranges = laser.ranges.data();
  scan.n_ranges = (uint32_t)laser.ranges.size();
  scan.angle_min = laser.angle_min;
  scan.angle_increment = laser.angle_increment;
  if (laser.frame_id != iface_params_.controller_frame_) { // :143
    PlanarTransform t;
    if (lookup(laser.frame_id, t)) {
      scan.has_tf = 1;
      scan.tf_x = t.x;
      scan.tf_y = t.y;
      scan.tf_yaw = t.yaw;
    }
    // a failed transform leaves the point untransformed in the reference (:160-169 `continue`)
  }

  // people positions in the controller frame (:176-209)
  people_mutex_.lock();
  PeopleMsg people = people_;
  people_mutex_.unlock();
  std::vector<double> people_xy;
  if (!people.people.empty() && people.frame_id != iface_params_.controller_frame_) {
    PlanarTransform t;
    if (!lookup(people.frame_id, t))
      return; // :198-204
    const double c = std::cos(t.yaw), s = std::sin(t.yaw);
    for (const PersonMsg &p : people.people) {
      people_xy.push_back(c * p.x - s * p.y + t.x);
      people_xy.push_back(s * p.x + c * p.y + t.y);
    }
  } else {
    for (const PersonMsg &p : people.people) {
      people_xy.push_back(p.x);
      people_xy.push_back(p.y);
    }
  }
  scan.people_xy = people_xy.data();
  scan.n_people = (uint32_t)(people_xy.size() / 2);

  std::vector<Point2D> points;
  if (scan.n_ranges) {
    if (!ensureContext())
      return;
    std::vector<double> xy(2 * (size_t)scan.n_ranges);
    uint32_t n = 0;
    if (sfw_laser_obstacles(ctx_, &scan, 1, iface_params_.max_obstacle_dist_, iface_params_.person_radius_,
                            xy.data(), scan.n_ranges, &n) != SFW_OK) {
      error_ = sfw_last_error(ctx_);
      return;
    }
    points.resize(n);
    for (uint32_t i = 0; i < n; ++i)
      points[i] = Point2D{xy[2 * i], xy[2 * i + 1]};
  }
  obs_mutex_.lock();
  obstacles_ = points;
  obs_mutex_.unlock();
}

void SFMSensorInterface::peopleCb(const PeopleMsg &people) {
  if (!running_ || !odom_received_)
    return;
  people_mutex_.lock();
  people_ = people;
  people_mutex_.unlock();

  std::vector<Agent> agents;
  const bool other_frame = people.frame_id != iface_params_.controller_frame_;
  PlanarTransform t;
  const bool have_tf = other_frame && lookup(people.frame_id, t);
  const double c = have_tf ? std::cos(t.yaw) : 1.0, s = have_tf ? std::sin(t.yaw) : 0.0;
  for (const PersonMsg &p : people.people) {
    Agent ag;
    ag.id = p.id;            // std::stoi(tags[0])
    ag.groupId = p.group_id; // std::stoi(tags[1])
    double px = p.x, py = p.y, yaw = p.yaw;
    if (other_frame) {
      if (!have_tf)
        return; // :463-467
      px = c * p.x - s * p.y + t.x;
      py = s * p.x + c * p.y + t.y;
      yaw = p.yaw + t.yaw;
    }
    ag.position = Point2D{px, py};
    // transformVector (:640-672): free vector, rotation only; identity inside one frame; a failed lookup
    // leaves a zero vector (nv stays default-constructed)
    double vx = p.vx, vy = p.vy;
    if (other_frame) {
      vx = c * p.vx - s * p.vy;
      vy = s * p.vx + c * p.vy;
    }
    ag.velocity = Point2D{vx, vy};
    ag.linearVelocity = std::sqrt(vx * vx + vy * vy);
    if (std::fabs(ag.linearVelocity) < 0.09)
      ag.yaw = yaw;
    else
      ag.yaw = std::atan2(vy, vx);
    ag.angularVelocity = p.wz;
    ag.radius = iface_params_.person_radius_;
    ag.teleoperated = false;
    // naive goal: position + naive_goal_time * velocity (:493-500)
    const double gt = iface_params_.naive_goal_time_;
    ag.has_goal = true;
    ag.goal_center = Point2D{px + gt * vx, py + gt * vy};
    ag.goal_radius = iface_params_.person_radius_;
    ag.desiredVelocity = iface_params_.people_velocity_;
    agents.push_back(ag);
  }
  // every agent carries the current obstacle list (:513-518)
  obs_mutex_.lock();
  std::vector<Point2D> obs_points = obstacles_;
  obs_mutex_.unlock();
  for (Agent &a : agents)
    a.obstacles1 = obs_points;

  agents_mutex_.lock();
  agents_.resize(people.people.size() + 1);
  agents_[0].obstacles1 = obs_points;
  for (size_t i = 1; i < agents_.size(); ++i)
    agents_[i] = agents[i - 1];
  agents_mutex_.unlock();
}

void SFMSensorInterface::odomCb(const OdometryMsg &odom) {
  if (!running_)
    return;
  odom_received_ = true;
  agents_mutex_.lock();
  Agent agent = agents_[0];
  agents_mutex_.unlock();
  agent.position = Point2D{odom.x, odom.y};
  agent.yaw = odom.yaw;
  agent.linearVelocity = std::sqrt(odom.vx * odom.vx + odom.vy * odom.vy);
  agent.angularVelocity = odom.wz;
  // odometry twist is expressed in the ROBOT frame and stored as is (:565-575)
  agent.velocity = Point2D{odom.vx, odom.vy};
  agents_mutex_.lock();
  agents_[0] = agent;
  agents_mutex_.unlock();
}

std::vector<Agent> SFMSensorInterface::getAgents() {
  agents_mutex_.lock();
  std::vector<Agent> agents = agents_;
  agents_mutex_.unlock();
  return agents;
}

} // namespace social_force_window_planner

// ---------------------------------------------------------------------------------------------------
// C wrapper so the tests (ctypes) can drive the class the way ROS would: same argument layout as
// oracle/ref_sensor_harness.cpp's sfw_ref_sensor_run.
// ---------------------------------------------------------------------------------------------------
using social_force_window_planner::Agent;
using social_force_window_planner::InterfaceParams;
using social_force_window_planner::SFMSensorInterface;

extern "C" int sfws_sensor_run(const float *ranges, uint32_t n_ranges, float angle_min, float angle_inc,
                               int laser_has_tf, const double *people, uint32_t n_people, int people_has_tf,
                               const double *odom, const double *params, const double *tf, double *agents_out,
                               double *obstacles_out, uint32_t max_obstacles, uint32_t *n_obstacles, int device,
                               uint64_t *launches) {
  InterfaceParams ip;
  ip.max_obstacle_dist_ = (float)params[0];
  ip.person_radius_ = (float)params[1];
  ip.naive_goal_time_ = (float)params[2];
  ip.people_velocity_ = (float)params[3];
  ip.robot_radius_ = (float)params[4];
  ip.max_robot_vel_x_ = (float)params[5];
  SFMSensorInterface iface(ip, device);
  iface.setTransform("laser", {tf[0], tf[1], tf[2]});
  iface.setTransform("map", {tf[0], tf[1], tf[2]});
  iface.start();
  social_force_window_planner::OdometryMsg od{odom[0], odom[1], odom[2], odom[3], odom[4], odom[5]};
  social_force_window_planner::PeopleMsg pp;
  pp.frame_id = people_has_tf ? "map" : "odom";
  for (uint32_t i = 0; i < n_people; ++i) {
    const double *r = people + 8 * i;
    pp.people.push_back({r[0], r[1], r[2], r[3], r[4], r[5], (int)r[6], (int)r[7]});
  }
  social_force_window_planner::LaserScanMsg ls;
  ls.frame_id = laser_has_tf ? "laser" : "odom";
  ls.angle_min = angle_min;
  ls.angle_increment = angle_inc;
  ls.ranges.assign(ranges, ranges + n_ranges);
  iface.odomCb(od);
  iface.peopleCb(pp);
  iface.laserCb(ls);
  iface.peopleCb(pp);
  iface.odomCb(od);
  if (launches)
    *launches = iface.kernelLaunches();
  if (!iface.lastError().empty())
    return -2;
  std::vector<Agent> ag = iface.getAgents();
  if (ag.size() != n_people + 1)
    return -1;
  for (size_t i = 0; i < ag.size(); ++i) {
    double *o = agents_out + 16 * i;
    const Agent &a = ag[i];
    o[0] = a.position.x;
    o[1] = a.position.y;
    o[2] = a.velocity.x;
    o[3] = a.velocity.y;
    o[4] = a.yaw;
    o[5] = a.linearVelocity;
    o[6] = a.angularVelocity;
    o[7] = a.radius;
    o[8] = a.desiredVelocity;
    o[9] = a.has_goal ? a.goal_center.x : 0.0;
    o[10] = a.has_goal ? a.goal_center.y : 0.0;
    o[11] = a.has_goal ? a.goal_radius : 0.0;
    o[12] = a.has_goal ? 1.0 : 0.0;
    o[13] = (double)a.groupId;
    o[14] = (double)a.id;
    o[15] = (double)a.obstacles1.size();
  }
  const std::vector<social_force_window_planner::Point2D> &obs = ag[0].obstacles1;
  *n_obstacles = (uint32_t)obs.size();
  for (size_t i = 0; i < obs.size() && i < max_obstacles; ++i) {
    obstacles_out[2 * i] = obs[i].x;
    obstacles_out[2 * i + 1] = obs[i].y;
  }
  return 0;
}
