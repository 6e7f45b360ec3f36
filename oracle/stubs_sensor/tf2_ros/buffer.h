// tf2_ros::Buffer stand-in: `transform` applies ONE planar rigid transform (set by the harness) to
// points, poses and free vectors, or throws tf2::TransformException when told to fail.
#pragma once
#include <geometry_msgs/msg/point_stamped.hpp>
#include <geometry_msgs/msg/pose_stamped.hpp>
#include <tf2/utils.h>
namespace tf2_ros {
class Buffer {
public:
  double tx = 0, ty = 0, yaw = 0;
  bool fail = false;
  geometry_msgs::msg::PointStamped transform(const geometry_msgs::msg::PointStamped &in, const std::string &to) const {
    if (in.header.frame_id == to)
      return in; // tf: same frame -> identity
    check();
    geometry_msgs::msg::PointStamped o = in;
    o.header.frame_id = to;
    const double c = std::cos(yaw), s = std::sin(yaw);
    o.point.x = c * in.point.x - s * in.point.y + tx;
    o.point.y = s * in.point.x + c * in.point.y + ty;
    return o;
  }
  geometry_msgs::msg::PoseStamped transform(const geometry_msgs::msg::PoseStamped &in, const std::string &to) const {
    if (in.header.frame_id == to)
      return in;
    check();
    geometry_msgs::msg::PoseStamped o = in;
    o.header.frame_id = to;
    const double c = std::cos(yaw), s = std::sin(yaw);
    o.pose.position.x = c * in.pose.position.x - s * in.pose.position.y + tx;
    o.pose.position.y = s * in.pose.position.x + c * in.pose.position.y + ty;
    tf2::Quaternion q;
    q.setRPY(0, 0, tf2::getYaw(in.pose.orientation) + yaw);
    o.pose.orientation = tf2::toMsg(q);
    return o;
  }
  geometry_msgs::msg::Vector3Stamped transform(const geometry_msgs::msg::Vector3Stamped &in, const std::string &to) const {
    if (in.header.frame_id == to)
      return in;
    check();
    geometry_msgs::msg::Vector3Stamped o = in;
    o.header.frame_id = to;
    const double c = std::cos(yaw), s = std::sin(yaw);
    o.vector.x = c * in.vector.x - s * in.vector.y;
    o.vector.y = s * in.vector.x + c * in.vector.y;
    return o;
  }
private:
  void check() const { if (fail) throw tf2::TransformException("no transform (harness)"); }
};
}
