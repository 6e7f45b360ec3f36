// nav2_util::geometry_utils::euclidean_distance on stamped poses (planar hypot, as nav2 Foxy does)
#pragma once
#include <cmath>
#include <geometry_msgs/msg/pose_stamped.hpp>
namespace nav2_util { namespace geometry_utils {
inline double euclidean_distance(const geometry_msgs::msg::PoseStamped &a, const geometry_msgs::msg::PoseStamped &b) {
  const double dx = a.pose.position.x - b.pose.position.x, dy = a.pose.position.y - b.pose.position.y;
  return std::hypot(dx, dy);
}
}}
