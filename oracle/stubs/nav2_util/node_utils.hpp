// Minimal stand-in for nav2_util::declare_parameter_if_not_declared (oracle/_ref only).
#pragma once
#include <rclcpp_lifecycle/lifecycle_node.hpp>
namespace nav2_util {
template <typename NodeT>
void declare_parameter_if_not_declared(NodeT node, const std::string &name,
                                       const rclcpp::ParameterValue &value) {
  if (!node->has_parameter(name))
    node->declare_parameter(name, value);
}
}
