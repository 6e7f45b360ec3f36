// Minimal stand-in for tf2::getYaw on a quaternion message (oracle/_ref only).
#pragma once
#include <cmath>
#include <geometry_msgs/msg/point.hpp>
namespace tf2 {
inline double getYaw(const geometry_msgs::msg::Quaternion &q) {
  // yaw of a unit quaternion (ZYX convention), as tf2's getYaw computes it
  double sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
  double sarg = -2.0 * (q.x * q.z - q.w * q.y) / (sqx + sqy + sqz + sqw);
  if (sarg <= -0.99999)
    return -2.0 * std::atan2(q.y, q.x);
  if (sarg >= 0.99999)
    return 2.0 * std::atan2(q.y, q.x);
  return std::atan2(2.0 * (q.x * q.y + q.w * q.z), sqw + sqx - sqy - sqz);
}
}
