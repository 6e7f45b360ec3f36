// tf2_ros::Buffer stand-in: knows ONE planar rigid transform T (set by the harness) that takes any
// non-controller frame into the controller frame `fixed`; `transform` applies T towards `fixed`, T^-1 away
// from it, the identity inside one frame, or throws tf2::TransformException when told to fail.
#pragma once
#include <geometry_msgs/msg/point_stamped.hpp>
#include <geometry_msgs/msg/pose_stamped.hpp>
#include <tf2/utils.h>
namespace tf2_ros {
class Buffer {
public:
  double tx = 0, ty = 0, yaw = 0;
  bool fail = false;
  std::string fixed = "odom";
  // forward: p -> R p + t (and yaw + yaw0); inverse: p -> R^T (p - t)
  void apply(bool inverse, double &x, double &y, double *heading, bool is_vector) const {
    const double c = std::cos(yaw), s = std::sin(yaw);
    if (!inverse) {
      const double nx = c * x - s * y + (is_vector ? 0.0 : tx), ny = s * x + c * y + (is_vector ? 0.0 : ty);
      x = nx; y = ny;
      if (heading) *heading += yaw;
    } else {
      const double dx = x - (is_vector ? 0.0 : tx), dy = y - (is_vector ? 0.0 : ty);
      const double nx = c * dx + s * dy, ny = -s * dx + c * dy;
      x = nx; y = ny;
      if (heading) *heading -= yaw;
    }
  }
  geometry_msgs::msg::PointStamped transform(const geometry_msgs::msg::PointStamped &in, const std::string &to) const {
    if (in.header.frame_id == to)
      return in; // tf: same frame -> identity
    check();
    geometry_msgs::msg::PointStamped o = in;
    o.header.frame_id = to;
    apply(to != fixed, o.point.x, o.point.y, nullptr, false);
    return o;
  }
  geometry_msgs::msg::PoseStamped transform(const geometry_msgs::msg::PoseStamped &in, const std::string &to) const {
    if (in.header.frame_id == to)
      return in;
    check();
    geometry_msgs::msg::PoseStamped o = in;
    o.header.frame_id = to;
    double heading = tf2::getYaw(in.pose.orientation);
    apply(to != fixed, o.pose.position.x, o.pose.position.y, &heading, false);
    tf2::Quaternion q;
    q.setRPY(0, 0, heading);
    o.pose.orientation = tf2::toMsg(q);
    return o;
  }
  geometry_msgs::msg::Vector3Stamped transform(const geometry_msgs::msg::Vector3Stamped &in, const std::string &to) const {
    if (in.header.frame_id == to)
      return in;
    check();
    geometry_msgs::msg::Vector3Stamped o = in;
    o.header.frame_id = to;
    apply(to != fixed, o.vector.x, o.vector.y, nullptr, true);
    return o;
  }
  // (in, out, frame, timeout) form used by sfw_planner_node.cpp:196,212
  template <typename T, typename Tol> T &transform(const T &in, T &out, const std::string &to, Tol) const {
    out = transform(in, to);
    return out;
  }
private:
  void check() const { if (fail) throw tf2::TransformException("no transform (harness)"); }
};
}
