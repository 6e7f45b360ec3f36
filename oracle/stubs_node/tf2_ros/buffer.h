// adds tf2::durationFromSec to the sensor-build Buffer stand-in (sfw_planner_node.cpp:195,211)
#pragma once
#include "../../stubs_sensor/tf2_ros/buffer.h"
namespace tf2 {
typedef double Duration;
inline Duration durationFromSec(double s) { return s; }
}
