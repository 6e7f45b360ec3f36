#pragma once
#include <geometry_msgs/msg/point.hpp>
namespace geometry_msgs { namespace msg {
struct PointStamped { std_msgs::msg::Header header; Point point; };
struct Vector3Stamped { std_msgs::msg::Header header; Vector3 vector; };
struct PoseWithCovariance { Pose pose; };
struct TwistWithCovariance { Twist twist; };
}}
