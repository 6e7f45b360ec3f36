// ref_node_harness.cpp — the reference's WHOLE plugin on the CPU: SFWPlannerNode (src/sfw_planner_node.cpp)
// with its own SFWPlanner, SFMSensorInterface, CostmapModel and Trajectory, every source compiled unmodified
// from /root/reference against the stand-in headers of oracle/stubs_node + stubs_sensor + stubs.
// One call = what nav2's controller_server does for one control tick: sensor callbacks, setPlan,
// computeVelocityCommands.  TEST INFRASTRUCTURE ONLY (pins host/sfw_node_host.cpp end to end).
#define protected public
#define private public
#include <social_force_window_planner/sfw_planner_node.hpp>
#undef protected
#undef private

#include <cstdint>
#include <cstring>

#include "../include/sfw_b200.h"

using social_force_window_planner::SFWPlannerNode;

namespace {
geometry_msgs::msg::PoseStamped pose_of(const std::string &frame, double x, double y, double yaw) {
  geometry_msgs::msg::PoseStamped ps;
  ps.header.frame_id = frame;
  ps.pose.position.x = x;
  ps.pose.position.y = y;
  tf2::Quaternion q;
  q.setRPY(0, 0, yaw);
  ps.pose.orientation = tf2::toMsg(q);
  return ps;
}
} // namespace

// The same harness drives the reference's own planner core (entry sfw_ref_node_run) and — compiled against
// plugin/include, where social_force_window_planner/sfw_planner.hpp is the B200 core — the reference's UNMODIFIED
// node + sensor interface on top of plugin/src/sfw_planner.cpp (entry sfw_dropin_node_run, Makefile target dropin).
#ifndef SFW_NODE_ENTRY
#define SFW_NODE_ENTRY sfw_ref_node_run
#endif

extern "C" {

// scene: costmap + footprint (+ robot.wpx.. unused).  ranges/people/odom as in ref_sensor_harness.cpp (all in the
// controller frame "odom").  plan_xyt: n_plan poses in the plan frame (plan_has_tf: frame "map", brought to
// "odom" by tf = {x, y, yaw}).  ticks: computeVelocityCommands is called `ticks` times on the same inputs
// (waypoint bookkeeping and path pruning carry over).  Outputs per tick t: cmd_out[3t..] = twist (vx, vy, wz),
// status_out[t] = 1 ok / 0 zero twist / -1 PlannerException; plan_left_out[t] = poses left in the pruned plan.
int SFW_NODE_ENTRY(const SfwParams *params, const double *ext, const SfwScene *scene, const float *ranges,
                     uint32_t n_ranges, float angle_min, float angle_inc, const double *people, uint32_t n_people,
                     const double *odom, const double *plan_xyt, uint32_t n_plan, int plan_has_tf, const double *tf,
                     uint32_t ticks, double *cmd_out, int *status_out, int *plan_left_out, int *goal_reached_out) {
  auto node = std::make_shared<rclcpp_lifecycle::LifecycleNode>();
  const std::string name = "FollowPath";
  const std::string b = name + ".";
  node->set_parameter(b + "max_trans_vel", rclcpp::ParameterValue(params->max_vel_x));
  node->set_parameter(b + "max_trans_acc", rclcpp::ParameterValue(params->max_trans_acc));
  node->set_parameter(b + "max_rot_acc", rclcpp::ParameterValue(params->max_rot_acc));
  node->set_parameter(b + "sim_time", rclcpp::ParameterValue(params->sim_time));
  node->set_parameter(b + "sim_granularity", rclcpp::ParameterValue(params->sim_granularity));
  node->set_parameter(b + "robot_radius", rclcpp::ParameterValue((double)params->robot_radius));
  node->set_parameter(b + "social_weight", rclcpp::ParameterValue(params->social_weight));
  node->set_parameter(b + "costmap_weight", rclcpp::ParameterValue(params->costmap_weight));
  node->set_parameter(b + "angle_weight", rclcpp::ParameterValue(params->angle_weight));
  node->set_parameter(b + "distance_weight", rclcpp::ParameterValue(params->distance_weight));
  node->set_parameter(b + "velocity_weight", rclcpp::ParameterValue(params->vel_weight));
  if (ext) {
    node->set_parameter(b + "min_trans_vel", rclcpp::ParameterValue(ext[0]));
    node->set_parameter(b + "max_rot_vel", rclcpp::ParameterValue(ext[1]));
    node->set_parameter(b + "min_rot_vel", rclcpp::ParameterValue(ext[2]));
    node->set_parameter(b + "min_in_place_rot_vel", rclcpp::ParameterValue(ext[3]));
    node->set_parameter(b + "yaw_goal_tolerance", rclcpp::ParameterValue(ext[4]));
    node->set_parameter(b + "xy_goal_tolerance", rclcpp::ParameterValue(ext[5]));
    node->set_parameter(b + "wp_tolerance", rclcpp::ParameterValue(ext[6]));
    node->set_parameter(b + "is_circular", rclcpp::ParameterValue(ext[7] != 0.0));
  }
  auto buf = std::make_shared<tf2_ros::Buffer>();
  buf->tx = tf[0];
  buf->ty = tf[1];
  buf->yaw = tf[2];
  nav2_costmap_2d::Costmap2D cm(scene->size_x, scene->size_y, scene->resolution, scene->origin_x, scene->origin_y,
                                scene->costmap);
  std::vector<geometry_msgs::msg::Point> fp(scene->n_footprint);
  for (uint32_t i = 0; i < scene->n_footprint; ++i) {
    fp[i].x = scene->footprint_xy[2 * i];
    fp[i].y = scene->footprint_xy[2 * i + 1];
  }
  auto cmros = std::make_shared<nav2_costmap_2d::Costmap2DROS>(&cm, "odom", fp);

  SFWPlannerNode plugin;
  plugin.configure(node, name, buf, cmros);
  plugin.activate();

  nav_msgs::msg::Path path;
  path.header.frame_id = plan_has_tf ? "map" : "odom";
  for (uint32_t i = 0; i < n_plan; ++i)
    path.poses.push_back(pose_of(path.header.frame_id, plan_xyt[3 * i], plan_xyt[3 * i + 1], plan_xyt[3 * i + 2]));
  plugin.setPlan(path); // starts the sensor interface (:113-117)

  // the sensor callbacks ROS would deliver before the tick
  auto od = std::make_shared<nav_msgs::msg::Odometry>();
  od->header.frame_id = "odom";
  od->pose.pose = pose_of("odom", odom[0], odom[1], odom[2]).pose;
  od->twist.twist.linear.x = odom[3];
  od->twist.twist.linear.y = odom[4];
  od->twist.twist.angular.z = odom[5];
  auto pp = std::make_shared<people_msgs::msg::People>();
  pp->header.frame_id = "odom";
  for (uint32_t i = 0; i < n_people; ++i) {
    const double *r = people + 8 * i;
    people_msgs::msg::Person p;
    p.position.x = r[0];
    p.position.y = r[1];
    p.position.z = r[2];
    p.velocity.x = r[3];
    p.velocity.y = r[4];
    p.velocity.z = r[5];
    p.tags = {std::to_string((int)r[6]), std::to_string((int)r[7])};
    pp->people.push_back(p);
  }
  auto ls = std::make_shared<sensor_msgs::msg::LaserScan>();
  ls->header.frame_id = "odom";
  ls->angle_min = angle_min;
  ls->angle_increment = angle_inc;
  ls->ranges.assign(ranges, ranges + n_ranges);
  plugin.sensor_iface_->odomCb(od);
  plugin.sensor_iface_->peopleCb(pp);
  plugin.sensor_iface_->laserCb(ls);
  plugin.sensor_iface_->peopleCb(pp);
  plugin.sensor_iface_->odomCb(od);

  geometry_msgs::msg::Twist speed;
  speed.linear.x = odom[3];
  speed.linear.y = odom[4];
  speed.angular.z = odom[5];
  const geometry_msgs::msg::PoseStamped pose = pose_of("odom", odom[0], odom[1], odom[2]);
  for (uint32_t t = 0; t < ticks; ++t) {
    try {
      geometry_msgs::msg::TwistStamped v = plugin.computeVelocityCommands(pose, speed);
      cmd_out[3 * t] = v.twist.linear.x;
      cmd_out[3 * t + 1] = v.twist.linear.y;
      cmd_out[3 * t + 2] = v.twist.angular.z;
      // a successful tick stamps the frame (:303); a failed one returns the default-constructed message
      status_out[t] = v.header.frame_id.empty() ? 0 : 1;
    } catch (nav2_core::PlannerException &e) {
      cmd_out[3 * t] = cmd_out[3 * t + 1] = cmd_out[3 * t + 2] = 0.0;
      status_out[t] = -1;
    }
    plan_left_out[t] = (int)plugin.global_plan_.poses.size();
    goal_reached_out[t] = plugin.isGoalReached() ? 1 : 0;
  }
  plugin.deactivate();
  plugin.cleanup();
  return 0;
}

} // extern "C"
