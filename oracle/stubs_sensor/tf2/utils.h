// tf2 stand-in: getYaw (from oracle/stubs), Quaternion::setRPY, toMsg, TransformException.
#pragma once
#include "../../stubs/tf2/utils.h"
#include <stdexcept>
namespace tf2 {
struct TransformException : public std::runtime_error { using std::runtime_error::runtime_error; };
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
  void setRPY(double roll, double pitch, double yaw) { // tf2::Quaternion::setRPY
    const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
    x = sr * cp * cy - cr * sp * sy;
    y = cr * sp * cy + sr * cp * sy;
    z = cr * cp * sy - sr * sp * cy;
    w = cr * cp * cy + sr * sp * sy;
  }
};
inline geometry_msgs::msg::Quaternion toMsg(const Quaternion &q) {
  geometry_msgs::msg::Quaternion m;
  m.x = q.x; m.y = q.y; m.z = q.z; m.w = q.w;
  return m;
}
}
