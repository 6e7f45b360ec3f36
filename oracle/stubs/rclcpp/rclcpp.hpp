// Minimal stand-in for the slice of rclcpp the reference's planner core uses: a logger, a clock,
// a string-keyed parameter store and no-op logging macros (oracle/_ref only; TEST INFRASTRUCTURE).
#pragma once
#include <geometry_msgs/msg/point.hpp>
#include <map>
#include <memory>
#include <string>

namespace rclcpp {
struct Logger {};
struct Time {
  operator builtin_interfaces::msg::Time() const { return builtin_interfaces::msg::Time(); }
};
struct Duration {
  explicit Duration(double seconds) : s_(seconds) {}
  operator builtin_interfaces::msg::Duration() const {
    builtin_interfaces::msg::Duration d;
    d.sec = (int)s_;
    d.nanosec = (unsigned)((s_ - (int)s_) * 1e9);
    return d;
  }
  double s_;
};
struct Clock {
  Time now() const { return Time(); }
};
struct ParameterValue {
  enum Kind { DOUBLE, BOOL, STRING } kind;
  double d = 0.0;
  bool b = false;
  std::string s;
  ParameterValue() : kind(DOUBLE) {}
  ParameterValue(double v) : kind(DOUBLE), d(v) {}
  ParameterValue(float v) : kind(DOUBLE), d(v) {}
  ParameterValue(int v) : kind(DOUBLE), d(v) {}
  ParameterValue(bool v) : kind(BOOL), b(v) {}
  ParameterValue(const char *v) : kind(STRING), s(v) {}
  ParameterValue(const std::string &v) : kind(STRING), s(v) {}
};
} // namespace rclcpp

#define RCLCPP_INFO(...) do { } while (0)
#define RCLCPP_WARN(...) do { } while (0)
#define RCLCPP_DEBUG(...) do { } while (0)
#define RCLCPP_ERROR(...) do { } while (0)
#define RCLCPP_INFO_ONCE(...) do { } while (0)
