// sfw_node_host.hpp — host-side mirror of the reference's nav2 plugin class, the surface nav2's
// controller_server talks to.
//
// Mirrors social_force_window_planner::SFWPlannerNode : nav2_core::Controller (reference
// include/social_force_window_planner/sfw_planner_node.hpp:53-176, src/sfw_planner_node.cpp:47-330): same
// method names (configure / cleanup / activate / deactivate / setPlan / computeVelocityCommands /
// isGoalReached) and the same per-tick sequence — transform the robot pose into the costmap frame, cut the
// global plan to the local costmap and prune what lies behind (transformGlobalPlan, :119-185), updatePlan,
// findBestAction, zero twist when no trajectory is valid — with ROS message types replaced by the fields
// that are read.  It owns one SFWPlanner and one SFMSensorInterface mirror; everything data-parallel under
// them runs in libsfw_b200.so.  With ROS 2 present the real plugin class is this control flow verbatim over
// the ROS types; `sfw_plugin.xml` and the exported class name do not change.
#ifndef SFW_NODE_HOST_HPP
#define SFW_NODE_HOST_HPP

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "sfw_planner_host.hpp"
#include "sfw_sensor_host.hpp"

namespace social_force_window_planner {

// nav2_core::PlannerException
class PlannerException : public std::runtime_error {
public:
  explicit PlannerException(const std::string &d) : std::runtime_error(d) {}
};

struct PoseStampedMsg {
  std::string frame_id;
  Pose2D pose;
};
struct PathMsg {
  std::string frame_id;
  std::vector<Pose2D> poses;
};
struct TwistStampedMsg {
  std::string frame_id; // empty on the zero-twist returns, like the reference's default-constructed message
  Twist2D twist;
};

class SFWPlannerNode {
public:
  SFWPlannerNode() = default;
  // reference configure(parent, name, tf, costmap_ros) (:47-77): costmap_ros contributes the live costmap,
  // its global frame and the robot footprint; the parameters come from the node's parameter server.
  void configure(const ControllerParams &params, const InterfaceParams &iface_params, const CostmapView *costmap,
                 const std::string &costmap_global_frame, const std::vector<Point2D> &footprint, int device = 0);
  void cleanup() {}
  void activate() {}
  void deactivate();
  void setPlan(const PathMsg &path);
  TwistStampedMsg computeVelocityCommands(const PoseStampedMsg &pose, const Twist2D &speed);
  bool isGoalReached() { return sfw_planner_->isGoalReached(); }

  // frame -> costmap global frame (what tf_->transform does); the inverse is derived when needed
  void setTransform(const std::string &from_frame, const PlanarTransform &tf);
  SFMSensorInterface &sensorInterface() { return *sensor_iface_; }
  SFWPlanner &planner() { return *sfw_planner_; }
  const PathMsg &globalPlan() const { return global_plan_; }
  const PathMsg &lastTransformedPlan() const { return transformed_plan_; } // what global_path_pub_ publishes

protected:
  PathMsg transformGlobalPlan(const PoseStampedMsg &rpose);
  bool transformPose(const std::string &frame, const PoseStampedMsg &in_pose, PoseStampedMsg &out_pose) const;

  std::shared_ptr<SFWPlanner> sfw_planner_;
  std::shared_ptr<SFMSensorInterface> sensor_iface_;
  const CostmapView *costmap_ = nullptr;
  std::string global_frame_;
  PathMsg global_plan_, transformed_plan_;
  std::vector<std::pair<std::string, PlanarTransform>> tf_;
};

} // namespace social_force_window_planner
#endif
