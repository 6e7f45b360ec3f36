// Minimal stand-in for visualization_msgs::msg::MarkerArray (oracle/_ref only).
#pragma once
#include <geometry_msgs/msg/point.hpp>
namespace visualization_msgs { namespace msg {
struct Marker {
  enum { POINTS = 8, ADD = 0 };
  std_msgs::msg::Header header;
  std::string ns;
  int id = 0;
  int type = 0;
  int action = 0;
  geometry_msgs::msg::Pose pose;
  geometry_msgs::msg::Vector3 scale;
  std_msgs::msg::ColorRGBA color;
  builtin_interfaces::msg::Duration lifetime;
  std::vector<geometry_msgs::msg::Point> points;
};
struct MarkerArray { std::vector<Marker> markers; };
}}
