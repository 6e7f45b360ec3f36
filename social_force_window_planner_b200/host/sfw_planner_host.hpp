// sfw_planner_host.hpp — host-side mirror of the reference's planner core, calling the C ABI.
//
// Mirrors social_force_window_planner::SFWPlanner (reference include/social_force_window_planner/
// sfw_planner.hpp:234-398, src/sfw_planner.cpp:28-87,117-469,853-892): same method names, argument
// meaning and return conventions, with the ROS message types replaced by the plain fields the planner
// actually reads (this image has no ROS 2; INTEGRATION.md shows the two-line adapters).  Everything
// data-parallel — the (v, w) loop, scoreTrajectory, footprintCost, the lightsfm calls — is NOT here:
// it is one sfw_score() call into libsfw_b200.so (include/sfw_b200.h).  What stays on the host is the
// per-tick control flow of findBestAction (goal tolerance, rotate in place, waypoint selection, the
// approach branch: sfw_tick.hpp, shared with the ROS-typed plugin class of plugin/) and the scene packer that turns the sensor interface's agent list
// (reference src/sensor_interface.cpp:618-631) into an SfwScene.
#ifndef SFW_PLANNER_HOST_HPP
#define SFW_PLANNER_HOST_HPP

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sfw_b200.h"
#include "sfw_tick.hpp"

namespace social_force_window_planner {

// geometry_msgs::msg::PoseStamped as the planner reads it: position.x/y and tf2::getYaw(orientation)
struct Pose2D {
  double x = 0.0, y = 0.0, yaw = 0.0;
};
// geometry_msgs::msg::Twist as the planner reads / writes it
struct Twist2D {
  double linear_x = 0.0, linear_y = 0.0, angular_z = 0.0;
};
struct Point2D {
  double x = 0.0, y = 0.0;
};

// ControllerParams (reference sfw_planner.hpp:55-227): same field names and defaults.
struct ControllerParams {
  double max_vel_x_ = 0.7, min_vel_x_ = 0.1, max_vel_th_ = 0.5, min_vel_th_ = 0.1;
  double max_trans_acc_ = 1.0, max_rot_acc_ = 1.0, min_in_place_vel_th_ = 0.3;
  double yaw_goal_tolerance_ = 0.05, xy_goal_tolerance_ = 0.1, wp_tolerance_ = 0.5;
  double sim_time_ = 1.0, sim_granularity_ = 0.025, angular_sim_granularity_ = 0.025;
  float robot_radius_ = 0.35f, people_radius_ = 0.35f;
  bool is_circular_ = true;
  float sfm_goal_weight_ = 2.0f, sfm_obstacle_weight_ = 20.0f, sfm_people_weight_ = 12.0f; // read, never applied
  double social_weight_ = 1.2, costmap_weight_ = 2.0, angle_weight_ = 0.7, distance_weight_ = 1.0,
         vel_weight_ = 1.0;
};

// sfm::Agent as SFMSensorInterface fills it (reference src/sensor_interface.cpp:32-37,447-504,553-579)
struct Agent {
  int id = -1, groupId = -1;
  Point2D position, velocity;
  double yaw = 0.0, linearVelocity = 0.0, angularVelocity = 0.0; // set by the sensor interface, unused by the scorer
  bool teleoperated = false;
  double radius = 0.35, desiredVelocity = 0.6;
  bool has_goal = false;
  Point2D goal_center;
  double goal_radius = 0.0;
  std::vector<Point2D> obstacles1; // every agent carries the same list (sensor_interface.cpp:513-524)
};

// nav2_costmap_2d::Costmap2D as the planner reads it (getCharMap / getSize / getResolution / getOrigin).
// The planner holds a reference to the LIVE costmap (sfw_planner.hpp:362): the view is re-read every tick.
struct CostmapView {
  const uint8_t *data = nullptr;
  uint32_t size_x = 0, size_y = 0;
  double resolution = 0.05, origin_x = 0.0, origin_y = 0.0;
};

// What RViz gets per evaluated sample (reference markers_: colour by validity, points of the rollout)
struct SampleMarker {
  double v = 0.0, w = 0.0;
  float cost = 0.0f; // < 0: rejected (red in the reference), >= 0 valid, best one green
};

// visualization_msgs::msg::Marker as the planner fills it (reference src/sfw_planner.cpp:91-110,345-386,
// 435-441): id = sample index, LINE_STRIP of the rollout's recorded points, colour by outcome
struct Marker {
  int id = 0;
  float r = 0.f, g = 0.f, b = 0.f, a = 1.0f; // initializeMarkers: colour (0,0,0,1)
  std::vector<double> points_xyz;            // (x, y, z) triples
};

class SFWPlanner {
public:
  // reference ctor (sfw_planner.cpp:28-87): params, costmap reference, footprint; builds the 5 x 9
  // sample sets from max_vel_x_/max_vel_th_ in the reference's order.  device: CUDA ordinal.
  SFWPlanner(const ControllerParams &params, const CostmapView *costmap, std::vector<Point2D> footprint_spec,
             int device = 0);
  ~SFWPlanner();
  SFWPlanner(const SFWPlanner &) = delete;
  SFWPlanner &operator=(const SFWPlanner &) = delete;

  bool findBestAction(const Pose2D &global_pose, const Twist2D &global_vel, Twist2D &cmd_vel);
  bool updatePlan(const std::vector<Pose2D> &new_plan);
  bool isGoalReached();
  void resetGoal() { tracker_.clearGoalFlag(); }
  void setFootprint(std::vector<Point2D> footprint) { footprint_spec_ = std::move(footprint); }
  std::vector<Point2D> getFootprint() const { return footprint_spec_; }

  // the sensor-interface seam: agents[0] is the robot (sensor_iface_->getAgents(), cpp:156)
  void setAgents(std::vector<Agent> agents) { agents_ = std::move(agents); }
  // ControllerParams are re-read every tick in the reference (cpp:125): mutate between ticks freely
  ControllerParams &params() { return params_; }
  // override the sample sets (BASELINE configs use denser grids than the shipped 5 x 9)
  void setSampleSets(std::vector<double> linvels, std::vector<double> angvels);
  const std::vector<double> &linvels() const { return linvels_; }
  const std::vector<double> &angvels() const { return angvels_; }

  // getMarkers() equivalent: per-sample command + cost of the last grid tick, winner index (or -1),
  // and the recorded rollout points of any sample of the last tick
  const std::vector<SampleMarker> &getMarkers() const { return markers_; }
  int bestIndex() const { return best_i_; }
  // the whole MarkerArray of the last grid tick, as getMarkers() returns it in the reference: one marker per
  // sample with the rollout's recorded points (one sfw_marker_points launch), red = rejected, blue = valid,
  // green + raised to z = 0.1 = the chosen one; the skipped (0,0) sample keeps the initial colour, no points
  std::vector<Marker> getMarkerArray();
  // reference sfw_planner.hpp:348 / sfw_planner.cpp:718-765 (never called upstream): can the robot brake to a
  // stop from (vl_x, vl_y, va) at pose (x, y, th) without an illegal footprint?  Uses the scene of the last
  // scored tick; false (with lastError set) before any tick.
  bool mayIStop(double vl_x, double vl_y, double va, double x, double y, double th, double dt);
  std::vector<Point2D> trajectoryPoints(uint32_t sample_index);

  // introspection for tests
  int wpIndex() const { return tracker_.waypointIndex(); }
  bool running() const { return tracker_.running(); }
  const std::string &lastError() const { return error_; }
  uint64_t kernelLaunches() const;

private:
  // one sfw_score call; costs (n_v*n_w floats) and best are filled.  false on ABI error (error_ set).
  bool score(const float rx, const float ry, const float rt, const float rvx, const float rvy, const float rvt,
             double wpx, double wpy, const double *lin, uint32_t n_v, const double *ang, uint32_t n_w,
             std::vector<float> &costs, SfwBest &best);
  bool ensureContext();

  ControllerParams params_;
  const CostmapView *costmap_;
  std::vector<Point2D> footprint_spec_;
  std::vector<Agent> agents_;
  std::vector<double> linvels_, angvels_;
  sfw_host::PlanTracker tracker_; // plan bookkeeping + the decisions of a tick that need no scoring
  std::vector<SampleMarker> markers_;
  std::vector<float> costs_;
  int device_;
  sfw_ctx *ctx_ = nullptr;
  std::string error_;
  int best_i_ = -1;
  bool staged_ = false;
};

} // namespace social_force_window_planner
#endif
