// nav2 costmap cost constants (public nav2 API values; oracle/_ref only).
#pragma once
namespace nav2_costmap_2d {
static const unsigned char NO_INFORMATION = 255;
static const unsigned char LETHAL_OBSTACLE = 254;
static const unsigned char INSCRIBED_INFLATED_OBSTACLE = 253;
static const unsigned char FREE_SPACE = 0;
}
