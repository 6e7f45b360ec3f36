#pragma once
#include <geometry_msgs/msg/point.hpp>
#include <nav2_costmap_2d/costmap_2d.hpp>
#include <string>
#include <vector>
namespace nav2_costmap_2d {
class Costmap2DROS {
public:
  Costmap2DROS(Costmap2D *cm, std::string global_frame, std::vector<geometry_msgs::msg::Point> fp)
      : cm_(cm), global_frame_(std::move(global_frame)), fp_(std::move(fp)) {}
  Costmap2D *getCostmap() { return cm_; }
  std::string getGlobalFrameID() { return global_frame_; }
  std::string getBaseFrameID() { return "base_link"; }
  std::vector<geometry_msgs::msg::Point> getRobotFootprint() { return fp_; }
private:
  Costmap2D *cm_;
  std::string global_frame_;
  std::vector<geometry_msgs::msg::Point> fp_;
};
}
