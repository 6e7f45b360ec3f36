// sfw_tick.hpp — what ONE control tick has to decide before (and instead of) scoring trajectories.
//
// The reference's SFWPlanner::findBestAction (src/sfw_planner.cpp:117-469) interleaves four things: plan
// bookkeeping (updatePlan :853-892, closest / next waypoint :236-273), the goal logic (:167-233), the approach
// heuristic (:276-334) and the sample loop (:338-417).  Here the first three are a small state machine without any
// ROS or CUDA type — PlanTracker::next() says which KIND of tick this is and which command / waypoint goes with
// it — so that the two callers (plugin/src/sfw_planner.cpp with ROS messages, host/sfw_planner_host.cpp with plain
// structs) share one implementation and only differ in how they talk to the scorer (sfw_score, include/sfw_b200.h).
//
// Arithmetic follows the reference to the bit where it decides something: the robot pose arrives narrowed to
// float (:145-152), distances are squared doubles, the heading error is normalised in float (hpp:399-407), the
// waypoint bearing uses the float overloads of cos / sin (`using namespace std` + a float argument, :276-277).
#ifndef SFW_TICK_HPP
#define SFW_TICK_HPP

#include <cmath>
#include <cstddef>
#include <vector>

namespace sfw_host {

// The ControllerParams fields the tick logic reads (reference sfw_planner.hpp:55-227), re-read every tick (:125).
struct TickLimits {
  double max_vel_x = 0.7, min_vel_x = 0.1, max_vel_th = 0.5, min_vel_th = 0.1, min_in_place_vel_th = 0.3;
  double yaw_goal_tolerance = 0.05, xy_goal_tolerance = 0.1, wp_tolerance = 0.5;
  bool is_circular = true;
};

struct PlanPose {
  double x = 0.0, y = 0.0, yaw = 0.0;
};

enum class TickKind {
  Idle,        // no plan: zero command, "ok" (:131-142)
  GoalReached, // inside both goal tolerances: zero command, stop running, raise the goal flag (:182-186)
  TurnInPlace, // inside the position tolerance only: rotate towards the goal heading (:187-220)
  Approach,    // closer than kApproachRadius to the goal: one heuristic command, grid as fallback (:282-334)
  Grid         // the (v, w) sample set (:338-417)
};

struct TickPlan {
  TickKind kind = TickKind::Idle;
  double vx = 0.0, vy = 0.0, vth = 0.0; // the command this tick proposes (all kinds but Grid)
  double wpx = 0.0, wpy = 0.0;          // waypoint the trajectories are scored against (Approach, Grid)
  bool needs_scoring = false;           // the proposal only stands if its one trajectory is legal
};

class PlanTracker {
public:
  static constexpr double kApproachRadius = 1.5; // :281
  static constexpr double kFarAway2 = 9999.0;    // :239: a fresh plan farther than this keeps waypoint 0

  // updatePlan (:853-892): an empty plan stops the controller, anything else restarts the tracking
  void setPlan(const std::vector<PlanPose> &plan) {
    goal_flag_ = false;
    plan_ = plan;
    if (plan_.empty()) {
      running_ = false;
      wp_ = -1;
      return;
    }
    wp_ = 0;
    running_ = true;
    fresh_ = true;
  }

  // isGoalReached (:894-901): reports the goal exactly once
  bool consumeGoalFlag() {
    const bool was = goal_flag_;
    goal_flag_ = false;
    return was;
  }
  void clearGoalFlag() { goal_flag_ = false; }

  bool running() const { return running_; }
  int waypointIndex() const { return wp_; }
  const std::vector<PlanPose> &plan() const { return plan_; }

  // Decide this tick.  (rx, ry, rt): robot pose already narrowed to float by the caller.
  TickPlan next(float rx, float ry, float rt, const TickLimits &L) {
    TickPlan t;
    goal_flag_ = false;
    if (!running_)
      return t; // Idle
    const PlanPose &goal = plan_.back();
    const double gx = rx - goal.x, gy = ry - goal.y;
    const double goal_d2 = gx * gx + gy * gy;

    if (goal_d2 < L.xy_goal_tolerance * L.xy_goal_tolerance) {
      if (std::fabs(goal.yaw - rt) < L.yaw_goal_tolerance) {
        t.kind = TickKind::GoalReached;
        running_ = false;
        goal_flag_ = true;
        return t;
      }
      t.kind = TickKind::TurnInPlace;
      t.vth = heading_error(goal.yaw, rt) > 0.0f ? L.min_in_place_vel_th : -L.min_in_place_vel_th;
      t.needs_scoring = !L.is_circular; // a round base can always turn; any other must check the sweep (:199-218)
      return t;
    }

    if (fresh_) {
      fresh_ = false;
      wp_ = entry_waypoint(rx, ry, L.wp_tolerance * L.wp_tolerance);
    }
    skip_reached_waypoints(rx, ry, L.wp_tolerance * L.wp_tolerance);
    t.wpx = plan_[wp_].x;
    t.wpy = plan_[wp_].y;
    t.kind = TickKind::Grid;

    if (goal_d2 < kApproachRadius * kApproachRadius) {
      // slow down with the distance to the goal, turn with the bearing of the waypoint (:283-292)
      const double ox = t.wpx - rx, oy = t.wpy - ry;
      const double ahead = ox * std::cos(rt) + oy * std::sin(rt);   // float cos / sin, like the reference
      const double left = -ox * std::sin(rt) + oy * std::cos(rt);
      const double bearing = std::atan2(left, ahead);
      t.kind = TickKind::Approach;
      t.vx = L.min_vel_x + (L.max_vel_x - L.min_vel_x) * (std::sqrt(goal_d2) / kApproachRadius);
      t.vy = 0.0;
      t.vth = L.min_vel_th + (L.max_vel_th - L.min_vel_th) * std::fabs(bearing) / M_PI;
      if (bearing < 0.0)
        t.vth *= -1;
      t.needs_scoring = true;
    }
    return t;
  }

private:
  // goal heading minus robot heading, wrapped in float (hpp:399-407; the limits are (float)-pi and (float)pi)
  static float heading_error(double goal_yaw, float rt) {
    const float d = goal_yaw - rt;
    const float lo = -M_PI, hi = M_PI;
    return d >= lo ? lo + std::fmod(d - lo, hi - lo) : hi - std::fmod(lo - d, hi - lo);
  }

  // First tick on a new plan (:236-257): scanning from the goal backwards, the first waypoint within the
  // tolerance; if there is none, the closest one (of those nearer than kFarAway2), else waypoint 0.
  int entry_waypoint(float rx, float ry, double tol2) const {
    int chosen = 0;
    double nearest = kFarAway2;
    for (int i = (int)plan_.size() - 1; i >= 0; --i) {
      const double ex = rx - plan_[i].x, ey = ry - plan_[i].y;
      const double d2 = ex * ex + ey * ey;
      if (d2 < tol2)
        return i;
      if (d2 < nearest) {
        nearest = d2;
        chosen = i;
      }
    }
    return chosen;
  }

  // A waypoint the robot already stands on is behind it (:260-273); the last one is never skipped.
  void skip_reached_waypoints(float rx, float ry, double tol2) {
    const int last = (int)plan_.size() - 1;
    while (wp_ < last) {
      const double ex = rx - plan_[wp_].x, ey = ry - plan_[wp_].y;
      if (!(ex * ex + ey * ey < tol2))
        break;
      ++wp_;
    }
  }

  std::vector<PlanPose> plan_;
  int wp_ = -1;
  bool running_ = false, fresh_ = false, goal_flag_ = false;
};

// The reference's sample sets (src/sfw_planner.cpp:65-85): 5 linear velocities 0 .. max, 9 angular ones ordered
// 0, +s, -s, +2s, -2s, ... (the order matters: it is the arg-min's last tie-break).
inline void default_sample_sets(double max_vel_x, double max_vel_th, std::vector<double> &lin,
                                std::vector<double> &ang) {
  constexpr int kSteps = 4;
  lin.clear();
  ang.clear();
  const double dv = max_vel_x / kSteps, dw = max_vel_th / kSteps;
  for (int i = 0; i <= kSteps; ++i)
    lin.push_back(i * dv);
  ang.push_back(0.0);
  for (int i = 1; i <= kSteps; ++i) {
    ang.push_back(i * dw);
    ang.push_back(i * (-dw));
  }
}

} // namespace sfw_host
#endif
