// Marker lives with MarkerArray in the planner-core stand-ins
#pragma once
#include <visualization_msgs/msg/marker_array.hpp>
