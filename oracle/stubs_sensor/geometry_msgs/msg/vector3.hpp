#pragma once
#include <geometry_msgs/msg/point_stamped.hpp>
