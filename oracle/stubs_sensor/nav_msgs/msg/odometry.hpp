#pragma once
#include <geometry_msgs/msg/point_stamped.hpp>
namespace nav_msgs { namespace msg {
struct Odometry {
  using SharedPtr = std::shared_ptr<Odometry>;
  std_msgs::msg::Header header;
  std::string child_frame_id;
  geometry_msgs::msg::PoseWithCovariance pose;
  geometry_msgs::msg::TwistWithCovariance twist;
};
}}
