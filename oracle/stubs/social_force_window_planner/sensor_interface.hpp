// Shadow of the reference's sensor_interface.hpp for oracle/_ref: exposes only the seam the
// planner core uses (SFMSensorInterface::getAgents, /root/reference/src/sfw_planner.cpp:156) so a
// test harness can inject a fixed agent snapshot instead of ROS subscriptions.  It must be found
// BEFORE /root/reference/include on the include path.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <lightsfm/angle.hpp>
#include <lightsfm/sfm.hpp>
#include <lightsfm/vector2d.hpp>
#include <vector>
namespace social_force_window_planner {
class SFMSensorInterface {
public:
  std::vector<sfm::Agent> getAgents() { return agents_; }
  void setAgents(const std::vector<sfm::Agent> &a) { agents_ = a; }
  void start() {}
  void stop() {}
private:
  std::vector<sfm::Agent> agents_;
};
}
