// rclcpp stand-in for the sensor-interface build: the planner-core stand-in plus QoS / Subscription.
#pragma once
#include "../../stubs/rclcpp/rclcpp.hpp"
#include <functional>
namespace rclcpp {
struct SensorDataQoS {};
template <typename T> struct Subscription { using SharedPtr = std::shared_ptr<Subscription<T>>; };
}
