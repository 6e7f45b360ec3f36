// Restatement of lightsfm's utils::Vector2d (un-vendored third party, see angle.hpp).
// TEST INFRASTRUCTURE ONLY.  normalized() of the zero vector is the zero vector.
#ifndef SFW_STUB_LIGHTSFM_VECTOR2D_HPP
#define SFW_STUB_LIGHTSFM_VECTOR2D_HPP
#include "angle.hpp"
#include <cmath>

namespace utils {

class Vector2d {
public:
  Vector2d() : x_(0.0), y_(0.0) {}
  Vector2d(double x, double y) : x_(x), y_(y) {}
  virtual ~Vector2d() {}

  double getX() const { return x_; }
  double getY() const { return y_; }
  void setX(double x) { x_ = x; }
  void setY(double y) { y_ = y; }
  void set(double x, double y) {
    x_ = x;
    y_ = y;
  }
  double operator()(int i) const { return i == 0 ? x_ : y_; }

  double squaredNorm() const { return x_ * x_ + y_ * y_; }
  double norm() const { return std::sqrt(x_ * x_ + y_ * y_); }
  Vector2d &normalize() {
    double n = norm();
    if (n > 0.0) {
      x_ /= n;
      y_ /= n;
    }
    return *this;
  }
  Vector2d normalized() const {
    Vector2d v(*this);
    v.normalize();
    return v;
  }
  Vector2d leftNormalVector() const { return Vector2d(-y_, x_); }
  Vector2d rightNormalVector() const { return Vector2d(y_, -x_); }
  Angle angle() const { return Angle::fromRadian(std::atan2(y_, x_)); }
  Angle angleTo(const Vector2d &o) const { return o.angle() - angle(); }
  double dot(const Vector2d &o) const { return x_ * o.x_ + y_ * o.y_; }

  Vector2d operator-() const { return Vector2d(-x_, -y_); }
  Vector2d operator+(const Vector2d &o) const {
    return Vector2d(x_ + o.x_, y_ + o.y_);
  }
  Vector2d operator-(const Vector2d &o) const {
    return Vector2d(x_ - o.x_, y_ - o.y_);
  }
  Vector2d operator*(double s) const { return Vector2d(x_ * s, y_ * s); }
  Vector2d operator/(double s) const { return Vector2d(x_ / s, y_ / s); }
  Vector2d &operator+=(const Vector2d &o) {
    x_ += o.x_;
    y_ += o.y_;
    return *this;
  }
  Vector2d &operator-=(const Vector2d &o) {
    x_ -= o.x_;
    y_ -= o.y_;
    return *this;
  }
  Vector2d &operator*=(double s) {
    x_ *= s;
    y_ *= s;
    return *this;
  }
  Vector2d &operator/=(double s) {
    x_ /= s;
    y_ /= s;
    return *this;
  }
  bool operator==(const Vector2d &o) const { return x_ == o.x_ && y_ == o.y_; }
  bool operator!=(const Vector2d &o) const { return !(*this == o); }

private:
  double x_, y_;
};

inline Vector2d operator*(double s, const Vector2d &v) { return v * s; }

} // namespace utils
#endif
