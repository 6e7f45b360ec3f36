// LifecycleNode stand-in with (inert) subscriptions and publishers; parameters as in oracle/stubs.
#pragma once
#include <rclcpp/rclcpp.hpp>
namespace rclcpp_lifecycle {
template <typename T> struct LifecyclePublisher {
  void on_activate() {}
  void on_deactivate() {}
  void publish(const T &m) { last = m; ++count; }
  T last;
  int count = 0;
};
class LifecycleNode {
public:
  using SharedPtr = std::shared_ptr<LifecycleNode>;
  rclcpp::Logger get_logger() const { return rclcpp::Logger(); }
  rclcpp::Clock *get_clock() { return &clock_; }
  bool has_parameter(const std::string &n) const { return params_.count(n) != 0; }
  void declare_parameter(const std::string &n, const rclcpp::ParameterValue &v) { params_.emplace(n, v); }
  void set_parameter(const std::string &n, const rclcpp::ParameterValue &v) { params_[n] = v; }
  bool get_parameter(const std::string &n, double &out) const {
    auto it = params_.find(n); if (it == params_.end()) return false; out = it->second.d; return true; }
  bool get_parameter(const std::string &n, float &out) const {
    auto it = params_.find(n); if (it == params_.end()) return false; out = static_cast<float>(it->second.d); return true; }
  bool get_parameter(const std::string &n, bool &out) const {
    auto it = params_.find(n); if (it == params_.end()) return false; out = it->second.b; return true; }
  bool get_parameter(const std::string &n, std::string &out) const {
    auto it = params_.find(n); if (it == params_.end()) return false; out = it->second.s; return true; }
  template <typename T, typename Q, typename F>
  typename rclcpp::Subscription<T>::SharedPtr create_subscription(const std::string &, const Q &, F &&) {
    return std::make_shared<rclcpp::Subscription<T>>();
  }
  template <typename T> std::shared_ptr<LifecyclePublisher<T>> create_publisher(const std::string &, int) {
    return std::make_shared<LifecyclePublisher<T>>();
  }
private:
  rclcpp::Clock clock_;
  std::map<std::string, rclcpp::ParameterValue> params_;
};
}
