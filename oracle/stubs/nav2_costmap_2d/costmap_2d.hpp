// Minimal stand-in for nav2_costmap_2d::Costmap2D: a borrowed row-major uint8 grid with the
// documented worldToMap / getCost semantics.  oracle/_ref only.
#pragma once
#include <nav2_costmap_2d/cost_values.hpp>
namespace nav2_costmap_2d {
class Costmap2D {
public:
  Costmap2D(unsigned sx, unsigned sy, double res, double ox, double oy, const unsigned char *data)
      : size_x_(sx), size_y_(sy), resolution_(res), origin_x_(ox), origin_y_(oy), costmap_(data) {}
  bool worldToMap(double wx, double wy, unsigned int &mx, unsigned int &my) const {
    if (wx < origin_x_ || wy < origin_y_)
      return false;
    mx = static_cast<unsigned int>((wx - origin_x_) / resolution_);
    my = static_cast<unsigned int>((wy - origin_y_) / resolution_);
    return mx < size_x_ && my < size_y_;
  }
  unsigned char getCost(unsigned int mx, unsigned int my) const { return costmap_[my * size_x_ + mx]; }
  unsigned char *getCharMap() const { return const_cast<unsigned char *>(costmap_); } // nav2: the raw grid
  unsigned int getSizeInCellsX() const { return size_x_; }
  unsigned int getSizeInCellsY() const { return size_y_; }
  double getResolution() const { return resolution_; }
  double getOriginX() const { return origin_x_; }
  double getOriginY() const { return origin_y_; }
private:
  unsigned size_x_, size_y_;
  double resolution_, origin_x_, origin_y_;
  const unsigned char *costmap_;
};
}
