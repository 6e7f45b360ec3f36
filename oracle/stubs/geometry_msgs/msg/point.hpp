// Minimal stand-in for the geometry_msgs / std_msgs / builtin_interfaces message structs the
// reference's planner core touches (oracle/_ref only; TEST INFRASTRUCTURE).
#pragma once
#include <string>
#include <vector>
namespace builtin_interfaces { namespace msg {
struct Time { int sec = 0; unsigned nanosec = 0; };
struct Duration { int sec = 0; unsigned nanosec = 0; };
}}
namespace std_msgs { namespace msg {
struct Header { builtin_interfaces::msg::Time stamp; std::string frame_id; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}}
namespace geometry_msgs { namespace msg {
struct Point { double x = 0, y = 0, z = 0; };
struct Point32 { float x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::msg::Header header; Pose pose; };
struct Twist { Vector3 linear; Vector3 angular; };
struct Polygon { std::vector<Point32> points; };
}}
