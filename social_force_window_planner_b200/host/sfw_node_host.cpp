// sfw_node_host.cpp — see sfw_node_host.hpp.  Statement order follows src/sfw_planner_node.cpp.
#include "sfw_node_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace social_force_window_planner {

void SFWPlannerNode::configure(const ControllerParams &params, const InterfaceParams &iface_params,
                               const CostmapView *costmap, const std::string &costmap_global_frame,
                               const std::vector<Point2D> &footprint, int device) {
  costmap_ = costmap;
  global_frame_ = costmap_global_frame;
  sensor_iface_ = std::make_shared<SFMSensorInterface>(iface_params, device); // :62
  sfw_planner_ = std::make_shared<SFWPlanner>(params, costmap, footprint, device); // :74-76
}

void SFWPlannerNode::deactivate() { sensor_iface_->stop(); }

void SFWPlannerNode::setPlan(const PathMsg &path) { // :113-117
  sensor_iface_->start();
  global_plan_ = path;
}

void SFWPlannerNode::setTransform(const std::string &from_frame, const PlanarTransform &tf) {
  for (auto &e : tf_)
    if (e.first == from_frame) {
      e.second = tf;
      sensor_iface_->setTransform(from_frame, tf);
      return;
    }
  tf_.emplace_back(from_frame, tf);
  sensor_iface_->setTransform(from_frame, tf);
}

// reference :187-204.  Known frames: the costmap global frame and those registered with setTransform.
bool SFWPlannerNode::transformPose(const std::string &frame, const PoseStampedMsg &in_pose,
                                   PoseStampedMsg &out_pose) const {
  if (in_pose.frame_id == frame) {
    out_pose = in_pose;
    return true;
  }
  const bool to_global = frame == global_frame_;
  const std::string &other = to_global ? in_pose.frame_id : frame;
  if (!to_global && in_pose.frame_id != global_frame_)
    return false; // would need a chain through the global frame: not something the reference's callers do
  for (const auto &e : tf_)
    if (e.first == other) {
      const PlanarTransform &t = e.second;
      const double c = std::cos(t.yaw), s = std::sin(t.yaw);
      out_pose.frame_id = frame;
      if (to_global) {
        out_pose.pose.x = c * in_pose.pose.x - s * in_pose.pose.y + t.x;
        out_pose.pose.y = s * in_pose.pose.x + c * in_pose.pose.y + t.y;
        out_pose.pose.yaw = in_pose.pose.yaw + t.yaw;
      } else {
        const double dx = in_pose.pose.x - t.x, dy = in_pose.pose.y - t.y;
        out_pose.pose.x = c * dx + s * dy;
        out_pose.pose.y = -s * dx + c * dy;
        out_pose.pose.yaw = in_pose.pose.yaw - t.yaw;
      }
      return true;
    }
  return false; // tf2::TransformException
}

// reference :119-185
PathMsg SFWPlannerNode::transformGlobalPlan(const PoseStampedMsg &rpose) {
  if (global_plan_.poses.empty())
    throw PlannerException("Received plan with zero length");
  PoseStampedMsg robot_pose;
  if (!transformPose(global_plan_.frame_id, rpose, robot_pose))
    throw PlannerException("Unable to transform robot pose into global plan's frame");

  // plan poses farther than half the costmap's larger side cannot matter to the local planner
  const double max_costmap_dim = std::max(costmap_->size_x, costmap_->size_y);
  const double max_transform_dist = max_costmap_dim * costmap_->resolution / 2.0;
  auto dist = [&robot_pose](const Pose2D &p) { return std::hypot(robot_pose.pose.x - p.x, robot_pose.pose.y - p.y); };

  // nearest plan pose to the robot (the reference's min_by keeps the FIRST minimum)
  auto transformation_begin = global_plan_.poses.begin();
  {
    double lowest = dist(*transformation_begin);
    for (auto it = std::next(global_plan_.poses.begin()); it != global_plan_.poses.end(); ++it) {
      const double d = dist(*it);
      if (d < lowest) {
        lowest = d;
        transformation_begin = it;
      }
    }
  }
  // first pose after it that lies beyond that distance: the cut
  auto transformation_end = std::find_if(transformation_begin, global_plan_.poses.end(),
                                         [&](const Pose2D &p) { return dist(p) > max_transform_dist; });

  PathMsg transformed_plan;
  for (auto it = transformation_begin; it != transformation_end; ++it) {
    PoseStampedMsg in{global_plan_.frame_id, *it}, out;
    transformPose(global_frame_, in, out); // result unchecked in the reference too (:163)
    transformed_plan.poses.push_back(out.pose);
  }
  transformed_plan.frame_id = global_frame_;

  // drop what lies behind the nearest pose from the stored plan
  global_plan_.poses.erase(global_plan_.poses.begin(), transformation_begin);
  transformed_plan_ = transformed_plan;
  if (transformed_plan.poses.empty())
    throw PlannerException("Resulting plan has 0 poses in it.");
  return transformed_plan;
}

// reference :219-307
TwistStampedMsg SFWPlannerNode::computeVelocityCommands(const PoseStampedMsg &pose, const Twist2D &speed) {
  TwistStampedMsg velStamp;
  sensor_iface_->start();
  PoseStampedMsg robot_pose;
  if (!transformPose(global_frame_, pose, robot_pose))
    throw PlannerException("Unable to transform robot pose into costmap's frame");
  PathMsg transformed_plan = transformGlobalPlan(pose);
  if (transformed_plan.poses.empty())
    return velStamp;
  // the planner reads its agents from the sensor interface at the start of findBestAction (cpp:156)
  sfw_planner_->setAgents(sensor_iface_->getAgents());
  sfw_planner_->updatePlan(transformed_plan.poses);
  Twist2D drive_cmds;
  const bool ok = sfw_planner_->findBestAction(robot_pose.pose, speed, drive_cmds);
  if (!ok)
    return velStamp; // zero twist, unstamped
  velStamp.frame_id = global_frame_;
  velStamp.twist = drive_cmds;
  return velStamp;
}

} // namespace social_force_window_planner

// ---------------------------------------------------------------------------------------------------
// C wrapper with the argument layout of oracle/ref_node_harness.cpp's sfw_ref_node_run
// ---------------------------------------------------------------------------------------------------
using namespace social_force_window_planner;

extern "C" int sfwn_node_run(const SfwParams *p, const double *ext, const SfwScene *scene, const float *ranges,
                             uint32_t n_ranges, float angle_min, float angle_inc, const double *people,
                             uint32_t n_people, const double *odom, const double *plan_xyt, uint32_t n_plan,
                             int plan_has_tf, const double *tf, uint32_t ticks, double *cmd_out, int *status_out,
                             int *plan_left_out, int *goal_reached_out, int device, uint64_t *launches_out) {
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
  InterfaceParams ip; // the sensor interface reads robot_radius / max_trans_vel from the same parameter names
  ip.robot_radius_ = p->robot_radius;
  ip.max_robot_vel_x_ = (float)p->max_vel_x;
  CostmapView view;
  view.data = scene->costmap;
  view.size_x = scene->size_x;
  view.size_y = scene->size_y;
  view.resolution = scene->resolution;
  view.origin_x = scene->origin_x;
  view.origin_y = scene->origin_y;
  std::vector<Point2D> fp(scene->n_footprint);
  for (uint32_t i = 0; i < scene->n_footprint; ++i)
    fp[i] = Point2D{scene->footprint_xy[2 * i], scene->footprint_xy[2 * i + 1]};
  SFWPlannerNode plugin;
  plugin.configure(cp, ip, &view, "odom", fp, device);
  plugin.setTransform("map", {tf[0], tf[1], tf[2]});
  plugin.activate();
  PathMsg path;
  path.frame_id = plan_has_tf ? "map" : "odom";
  for (uint32_t i = 0; i < n_plan; ++i)
    path.poses.push_back(Pose2D{plan_xyt[3 * i], plan_xyt[3 * i + 1], plan_xyt[3 * i + 2]});
  plugin.setPlan(path);

  OdometryMsg od{odom[0], odom[1], odom[2], odom[3], odom[4], odom[5]};
  PeopleMsg pp;
  pp.frame_id = "odom";
  for (uint32_t i = 0; i < n_people; ++i) {
    const double *r = people + 8 * i;
    pp.people.push_back({r[0], r[1], r[2], r[3], r[4], r[5], (int)r[6], (int)r[7]});
  }
  LaserScanMsg ls;
  ls.frame_id = "odom";
  ls.angle_min = angle_min;
  ls.angle_increment = angle_inc;
  ls.ranges.assign(ranges, ranges + n_ranges);
  SFMSensorInterface &si = plugin.sensorInterface();
  si.odomCb(od);
  si.peopleCb(pp);
  si.laserCb(ls);
  si.peopleCb(pp);
  si.odomCb(od);

  const PoseStampedMsg pose{"odom", Pose2D{odom[0], odom[1], odom[2]}};
  const Twist2D speed{odom[3], odom[4], odom[5]};
  for (uint32_t t = 0; t < ticks; ++t) {
    try {
      TwistStampedMsg v = plugin.computeVelocityCommands(pose, speed);
      cmd_out[3 * t] = v.twist.linear_x;
      cmd_out[3 * t + 1] = v.twist.linear_y;
      cmd_out[3 * t + 2] = v.twist.angular_z;
      status_out[t] = v.frame_id.empty() ? 0 : 1;
    } catch (PlannerException &) {
      cmd_out[3 * t] = cmd_out[3 * t + 1] = cmd_out[3 * t + 2] = 0.0;
      status_out[t] = -1;
    }
    plan_left_out[t] = (int)plugin.globalPlan().poses.size();
    goal_reached_out[t] = plugin.isGoalReached() ? 1 : 0;
  }
  if (launches_out)
    *launches_out = plugin.planner().kernelLaunches() + si.kernelLaunches();
  plugin.deactivate();
  plugin.cleanup();
  return 0;
}
