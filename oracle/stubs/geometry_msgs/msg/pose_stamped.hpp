#pragma once
#include <geometry_msgs/msg/point.hpp>
