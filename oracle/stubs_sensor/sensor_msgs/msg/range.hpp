#pragma once
#include <geometry_msgs/msg/point.hpp>
#include <memory>
namespace sensor_msgs { namespace msg { struct Range { std_msgs::msg::Header header; float range = 0; }; }}
