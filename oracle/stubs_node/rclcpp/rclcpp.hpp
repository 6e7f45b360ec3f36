#pragma once
#include "../../stubs_sensor/rclcpp/rclcpp.hpp"
namespace rclcpp {
inline Logger get_logger(const char *) { return Logger(); }
}
