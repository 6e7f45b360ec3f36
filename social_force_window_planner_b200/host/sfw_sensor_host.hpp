// sfw_sensor_host.hpp — host-side mirror of the reference's SFMSensorInterface, the producer of the
// agent snapshot the scoring path consumes.
//
// Mirrors social_force_window_planner::SFMSensorInterface (reference
// include/social_force_window_planner/sensor_interface.hpp:60-300, src/sensor_interface.cpp:20-631): same
// method names (laserCb / peopleCb / odomCb / getAgents / start / stop), same InterfaceParams fields and
// defaults, same callback semantics, with the ROS message types replaced by the plain fields the callbacks
// read (this image has no ROS 2).  tf2 is external to the reference: the planar transform a frame needs to
// reach the controller frame is registered with setTransform(); an unknown frame plays the role of a
// tf2::TransformException.
//
// The data-parallel part — the beams x people filter of laserCb (:103-229) — is NOT here: it is one
// sfw_laser_obstacles() call into libsfw_b200.so.  peopleCb / odomCb are message-rate host code.
#ifndef SFW_SENSOR_HOST_HPP
#define SFW_SENSOR_HOST_HPP

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "sfw_planner_host.hpp"

namespace social_force_window_planner {

// InterfaceParams (reference sensor_interface.hpp:60-155): same names, types and defaults
struct InterfaceParams {
  float max_robot_vel_x_ = 0.7f, robot_radius_ = 0.35f, person_radius_ = 0.35f;
  std::string robot_frame_ = "base_link", controller_frame_ = "odom";
  float max_obstacle_dist_ = 3.0f, naive_goal_time_ = 2.0f, people_velocity_ = 1.0f;
  std::string laser_topic_ = "scan", people_topic_ = "people", odom_topic_ = "odom";
};

// sensor_msgs/LaserScan as laserCb reads it
struct LaserScanMsg {
  std::string frame_id;
  float angle_min = 0.f, angle_increment = 0.f;
  std::vector<float> ranges;
};
// people_msgs/Person as peopleCb reads it: position.z carries the heading, velocity.z the turn rate,
// tags[0] / tags[1] the id and the group id (sensor_interface.cpp:448-449,458,488)
struct PersonMsg {
  double x = 0.0, y = 0.0, yaw = 0.0;
  double vx = 0.0, vy = 0.0, wz = 0.0;
  int id = 0, group_id = -1;
};
struct PeopleMsg {
  std::string frame_id;
  std::vector<PersonMsg> people;
};
// nav_msgs/Odometry as odomCb reads it (pose in the odom frame, twist in the ROBOT frame, :575)
struct OdometryMsg {
  double x = 0.0, y = 0.0, yaw = 0.0;
  double vx = 0.0, vy = 0.0, wz = 0.0;
};
// what tf_buffer_->transform(., controller_frame_) does to a point of the given frame, in the plane
struct PlanarTransform {
  double x = 0.0, y = 0.0, yaw = 0.0;
};

class SFMSensorInterface {
public:
  // device: CUDA ordinal of the context that runs the laser kernel
  explicit SFMSensorInterface(const InterfaceParams &params = InterfaceParams(), int device = 0);
  ~SFMSensorInterface();
  SFMSensorInterface(const SFMSensorInterface &) = delete;
  SFMSensorInterface &operator=(const SFMSensorInterface &) = delete;

  void laserCb(const LaserScanMsg &laser);   // reference :103-229
  void peopleCb(const PeopleMsg &people);    // :418-528
  void odomCb(const OdometryMsg &odom);      // :534-581
  std::vector<Agent> getAgents();            // :618-631
  void start() { running_ = true; }
  void stop() { running_ = false; }

  InterfaceParams &params() { return iface_params_; }
  void setTransform(const std::string &from_frame, const PlanarTransform &tf) { tf_[from_frame] = tf; }
  void clearTransform(const std::string &from_frame) { tf_.erase(from_frame); }
  const std::vector<Point2D> &obstacles() const { return obstacles_; } // what publish_obstacle_points shows
  const std::string &lastError() const { return error_; }
  uint64_t kernelLaunches() const;

private:
  bool lookup(const std::string &frame, PlanarTransform &out) const;
  bool ensureContext();

  InterfaceParams iface_params_;
  int device_;
  sfw_ctx *ctx_ = nullptr;
  std::map<std::string, PlanarTransform> tf_;
  std::vector<Agent> agents_; // 0: robot, 1..: others
  std::vector<Point2D> obstacles_;
  PeopleMsg people_;
  std::mutex agents_mutex_, obs_mutex_, people_mutex_, odom_mutex_;
  bool running_ = false, laser_received_ = false, odom_received_ = false;
  std::string error_;
};

} // namespace social_force_window_planner
#endif
