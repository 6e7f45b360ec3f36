// Minimal stand-in for nav2_costmap_2d footprint helpers used by the reference (oracle/_ref only).
#pragma once
#include <algorithm>
#include <cmath>
#include <geometry_msgs/msg/point.hpp>
#include <limits>
#include <vector>
namespace nav2_costmap_2d {
inline double stubDistToSegment(double px, double py, double x0, double y0, double x1, double y1) {
  double A = px - x0, B = py - y0, C = x1 - x0, D = y1 - y0;
  double dot = A * C + B * D, len_sq = C * C + D * D;
  double param = len_sq > 0 ? dot / len_sq : -1.0;
  double xx, yy;
  if (param < 0) { xx = x0; yy = y0; } else if (param > 1) { xx = x1; yy = y1; } else { xx = x0 + param * C; yy = y0 + param * D; }
  return std::hypot(px - xx, py - yy);
}
inline void calculateMinAndMaxDistances(const std::vector<geometry_msgs::msg::Point> &fp,
                                        double &min_dist, double &max_dist) {
  min_dist = std::numeric_limits<double>::max();
  max_dist = 0.0;
  if (fp.size() <= 2) return;
  for (size_t i = 0; i < fp.size(); ++i) {
    const auto &a = fp[i];
    const auto &b = fp[(i + 1) % fp.size()];
    double vd = std::hypot(a.x, a.y);
    double ed = stubDistToSegment(0.0, 0.0, a.x, a.y, b.x, b.y);
    min_dist = std::min(min_dist, std::min(vd, ed));
    max_dist = std::max(max_dist, std::max(vd, ed));
  }
}
inline geometry_msgs::msg::Polygon toPolygon(const std::vector<geometry_msgs::msg::Point> &pts) {
  geometry_msgs::msg::Polygon p;
  for (const auto &q : pts) { geometry_msgs::msg::Point32 r; r.x = (float)q.x; r.y = (float)q.y; r.z = (float)q.z; p.points.push_back(r); }
  return p;
}
}
