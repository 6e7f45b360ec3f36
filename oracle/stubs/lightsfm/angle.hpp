// Restatement of lightsfm's utils::Angle (robotics-upo/lightsfm, un-vendored, no pinned
// version: /root/reference/package.xml:31).  TEST INFRASTRUCTURE ONLY: lets the reference's
// own sources compile for oracle/_ref.  Written from the published upstream behaviour:
// radians stored wrapped to (-pi, pi]; sign() in {-1,0,1}.  Parity at this boundary is
// unpinned (the reference ships no tests and no lightsfm copy).
#ifndef SFW_STUB_LIGHTSFM_ANGLE_HPP
#define SFW_STUB_LIGHTSFM_ANGLE_HPP
#include <cmath>

namespace utils {

class Angle {
public:
  enum AngleRange { PI_RANGE, TWO_PI_RANGE };

  Angle() : value_(0.0) {}
  virtual ~Angle() {}

  static Angle fromRadian(double rad) { return Angle(rad); }
  static Angle fromDegree(double deg) { return Angle(deg * M_PI / 180.0); }

  double toRadian(AngleRange range = PI_RANGE) const {
    if (range == TWO_PI_RANGE && value_ < 0.0)
      return value_ + 2.0 * M_PI;
    return value_;
  }
  double toDegree(AngleRange range = PI_RANGE) const {
    return toRadian(range) * 180.0 / M_PI;
  }
  void setRadian(double rad) { value_ = wrap(rad); }
  void setDegree(double deg) { value_ = wrap(deg * M_PI / 180.0); }

  double cos() const { return std::cos(value_); }
  double sin() const { return std::sin(value_); }

  int sign() const {
    if (value_ == 0.0)
      return 0;
    return value_ > 0.0 ? 1 : -1;
  }

  Angle operator+(const Angle &o) const { return Angle(value_ + o.value_); }
  Angle operator-(const Angle &o) const { return Angle(value_ - o.value_); }
  Angle &operator+=(const Angle &o) {
    value_ = wrap(value_ + o.value_);
    return *this;
  }
  Angle &operator-=(const Angle &o) {
    value_ = wrap(value_ - o.value_);
    return *this;
  }
  bool operator==(const Angle &o) const { return value_ == o.value_; }
  bool operator!=(const Angle &o) const { return value_ != o.value_; }
  bool operator<(const Angle &o) const { return value_ < o.value_; }
  bool operator>(const Angle &o) const { return value_ > o.value_; }

private:
  explicit Angle(double rad) : value_(wrap(rad)) {}
  static double wrap(double v) {
    while (v <= -M_PI)
      v += 2.0 * M_PI;
    while (v > M_PI)
      v -= 2.0 * M_PI;
    return v;
  }
  double value_;
};

} // namespace utils
#endif
