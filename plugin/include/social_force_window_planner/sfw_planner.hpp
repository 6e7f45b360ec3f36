// social_force_window_planner/sfw_planner.hpp — the B200 planner core behind the reference's plugin surface.
//
// Drop-in replacement of ONE header + ONE source of robotics-upo/social_force_window_planner:
//   include/social_force_window_planner/sfw_planner.hpp   ->  this file
//   src/sfw_planner.cpp (+ costmap_model.cpp, trajectory.cpp, which only the old core used)
//                                                          ->  plugin/src/sfw_planner.cpp + libsfw_b200.so
// Everything else of the package stays as it is: sfw_plugin.xml, SFWPlannerNode (the nav2_core::Controller that
// pluginlib loads, src/sfw_planner_node.cpp), SFMSensorInterface, the launch / yaml files.  The class below has
// the reference's public signatures (reference include/.../sfw_planner.hpp:236-240 constructor, :261-263
// findBestAction, :271 updatePlan, :273-274 isGoalReached / resetGoal, :277-287 footprint accessors, :289
// getMarkers), so the reference's sfw_planner_node.cpp compiles against it UNMODIFIED — oracle/Makefile target
// `dropin` does exactly that and tests/test_gpu_dropin.py replays the whole-plugin fixtures through the result.
//
// What changed behind the surface: the (v, w) sample loop, scoreTrajectory, footprintCost and the lightsfm calls
// are one sfw_score() into hand-written sm_100a kernels (include/sfw_b200.h); the per-tick control flow that
// stays on the host is host/sfw_tick.hpp.  There is no CPU scoring path: without a usable CUDA device the
// constructor throws.
#ifndef _SFW_PLANNER_HPP_
#define _SFW_PLANNER_HPP_

#include <chrono>
#include <cmath>
#include <functional>
#include <memory>
#include <mutex>
#include <stdio.h>
#include <string>
#include <vector>

// the ROS / nav2 / tf2 headers the reference's header pulls in (its node source relies on some of them transitively)
#include "nav2_util/geometry_utils.hpp"
#include "nav2_util/node_utils.hpp"
#include "nav2_util/odometry_utils.hpp"
#include "rclcpp/rclcpp.hpp"
#include "rclcpp_lifecycle/lifecycle_node.hpp"
#include <tf2_ros/buffer.h>
#include <nav2_costmap_2d/cost_values.hpp>
#include <nav2_costmap_2d/costmap_2d.hpp>
#include <nav2_costmap_2d/footprint.hpp>
#include <tf2/LinearMath/Matrix3x3.h>
#include <tf2/utils.h>

#include <geometry_msgs/msg/point.hpp>
#include <geometry_msgs/msg/pose_stamped.hpp>
#include <geometry_msgs/msg/twist.hpp>
#include <visualization_msgs/msg/marker_array.hpp>

#include <social_force_window_planner/sensor_interface.hpp>

#include "sfw_b200.h"
#include "sfw_tick.hpp"

namespace social_force_window_planner {

// The controller's ROS parameters (names, types and defaults of reference sfw_planner.hpp:55-227, so existing
// yaml files keep working), plus `cuda_device`.
struct ControllerParams {
  void get(rclcpp_lifecycle::LifecycleNode *node, const std::string &name);

  std::string controller_frame_ = "odom", robot_base_frame_ = "base_link";
  double max_vel_x_ = 0.7, min_vel_x_ = 0.1, max_vel_th_ = 0.5, min_vel_th_ = 0.1;
  double max_trans_acc_ = 1.0, max_rot_acc_ = 1.0, min_in_place_vel_th_ = 0.3;
  double yaw_goal_tolerance_ = 0.05, xy_goal_tolerance_ = 0.1, wp_tolerance_ = 0.5;
  double sim_time_ = 1.0, sim_granularity_ = 0.025, angular_sim_granularity_ = 0.025;
  float robot_radius_ = 0.35f, people_radius_ = 0.35f;
  bool is_circular_ = true;
  float sfm_goal_weight_ = 2.0f, sfm_obstacle_weight_ = 20.0f, sfm_people_weight_ = 12.0f; // read, never applied upstream
  double social_weight_ = 1.2, costmap_weight_ = 2.0, angle_weight_ = 0.7, distance_weight_ = 1.0, vel_weight_ = 1.0;
  double cuda_device_ = 0.0; // which GPU scores (not a reference parameter)
};

class SFWPlanner {
public:
  SFWPlanner(const rclcpp_lifecycle::LifecycleNode::SharedPtr &parent, const std::string name,
             std::shared_ptr<SFMSensorInterface> &sensor_iface, const nav2_costmap_2d::Costmap2D &costmap,
             std::vector<geometry_msgs::msg::Point> footprint_spec);
  ~SFWPlanner();
  SFWPlanner(const SFWPlanner &) = delete;
  SFWPlanner &operator=(const SFWPlanner &) = delete;

  // One control tick: true + the command to send, or false (+ a zero / blocked command) when nothing is legal.
  bool findBestAction(const geometry_msgs::msg::PoseStamped &global_pose, const geometry_msgs::msg::Twist &global_vel,
                      geometry_msgs::msg::Twist &cmd_vel);
  bool updatePlan(const std::vector<geometry_msgs::msg::PoseStamped> &new_plan);
  bool isGoalReached();
  void resetGoal();

  void setFootprint(std::vector<geometry_msgs::msg::Point> footprint) { footprint_spec_ = footprint; }
  geometry_msgs::msg::Polygon getFootprintPolygon() const { return nav2_costmap_2d::toPolygon(footprint_spec_); }
  std::vector<geometry_msgs::msg::Point> getFootprint() const { return footprint_spec_; }

  // RViz markers of the last tick: one LINE_STRIP per sample (red rejected, blue legal, green + raised the winner)
  visualization_msgs::msg::MarkerArray &getMarkers();

  // B200 extras (not part of the reference surface)
  void setSampleSets(const std::vector<double> &linvels, const std::vector<double> &angvels); // denser (v, w) grids
  const std::string &lastError() const { return error_; }
  const char *lastKernel() const;

private:
  void initializeMarkers();
  void refreshMarkers();
  // one sfw_score call over lin x ang against waypoint (wpx, wpy); false on a library error (error_ set)
  bool score(float rx, float ry, float rt, float rvx, float rvy, float rvt, double wpx, double wpy,
             const std::vector<sfm::Agent> &agents, const double *lin, uint32_t n_v, const double *ang, uint32_t n_w);

  std::mutex configuration_mutex_;
  ControllerParams params_;
  rclcpp_lifecycle::LifecycleNode::SharedPtr node_;
  std::string name_;
  std::shared_ptr<SFMSensorInterface> sensor_iface_;
  const nav2_costmap_2d::Costmap2D &costmap_; // the LIVE costmap: re-read every tick (rolling window)
  std::vector<geometry_msgs::msg::Point> footprint_spec_;
  std::vector<double> linvels_, angvels_;
  visualization_msgs::msg::MarkerArray markers_;

  sfw_host::PlanTracker tracker_;
  sfw_ctx *ctx_ = nullptr;
  std::string error_;
  // results of the last scoring call
  std::vector<float> costs_;
  SfwBest best_;
  enum class Pending { None, Grid, Approach } pending_markers_ = Pending::None;
  // scratch of the scene packer
  std::vector<SfwPed> peds_;
  std::vector<double> obstacles_, footprint_xy_;
};

} // namespace social_force_window_planner

#endif
