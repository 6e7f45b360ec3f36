// ref_sensor_harness.cpp — drives the reference's OWN SFMSensorInterface (compiled unmodified from
// /root/reference/src/sensor_interface.cpp against oracle/stubs_sensor + oracle/stubs) on one set of
// laser / people / odometry messages and returns the agent snapshot getAgents() hands the planner.
// TEST INFRASTRUCTURE ONLY: it pins the restatement of laserCb / peopleCb / odomCb (oracle/sfw_oracle.c,
// social_force_window_planner_b200/host/sfw_sensor_host.cpp) to the reference's object code.  What stays
// external: tf2 (one planar rigid transform here) and lightsfm's Agent type (oracle/stubs/lightsfm).
#include <social_force_window_planner/sensor_interface.hpp>

#include <cstdint>
#include <cstring>

using social_force_window_planner::SFMSensorInterface;

extern "C" {

// people: n_people rows of 8 doubles {x, y, position.z (yaw), vx, vy, velocity.z (angular), id, groupId}
// odom:   {x, y, yaw, linear.x, linear.y, angular.z}
// params: {max_obstacle_dist, person_radius, naive_goal_time, people_velocity, robot_radius, max_trans_vel}
// tf:     {x, y, yaw} applied by tf_buffer_->transform when a message's frame is not the controller frame
// agents_out: (n_people + 1) rows of 16 doubles
//   {x, y, vx, vy, yaw, linearVelocity, angularVelocity, radius, desiredVelocity, goal_x, goal_y, goal_r,
//    n_goals, groupId, id, obstacles1.size()}
// obstacles_out: agents[0].obstacles1 (x, y) pairs, at most max_obstacles; *n_obstacles its size.
// Callback order: odom, people (so that laserCb has people to filter against), laser, people again (agents
// pick up the filtered obstacle list, reference :513-524), odom (robot agent refreshed, :553-579).
int sfw_ref_sensor_run(const float *ranges, uint32_t n_ranges, float angle_min, float angle_inc, int laser_has_tf,
                       const double *people, uint32_t n_people, int people_has_tf, const double *odom,
                       const double *params, const double *tf, double *agents_out, double *obstacles_out,
                       uint32_t max_obstacles, uint32_t *n_obstacles) {
  auto node = std::make_shared<rclcpp_lifecycle::LifecycleNode>();
  const std::string name = "FollowPath";
  node->set_parameter(name + ".sensor_interface.max_obstacle_dist", rclcpp::ParameterValue(params[0]));
  node->set_parameter(name + ".person_radius", rclcpp::ParameterValue(params[1]));
  node->set_parameter(name + ".sensor_interface.naive_goal_time", rclcpp::ParameterValue(params[2]));
  node->set_parameter(name + ".sensor_interface.people_velocity", rclcpp::ParameterValue(params[3]));
  node->set_parameter(name + ".robot_radius", rclcpp::ParameterValue(params[4]));
  node->set_parameter(name + ".max_trans_vel", rclcpp::ParameterValue(params[5]));
  auto buf = std::make_shared<tf2_ros::Buffer>();
  buf->tx = tf[0];
  buf->ty = tf[1];
  buf->yaw = tf[2];
  SFMSensorInterface iface(node, buf, name);
  iface.start();

  auto od = std::make_shared<nav_msgs::msg::Odometry>();
  od->header.frame_id = "odom";
  od->child_frame_id = "base_link";
  od->pose.pose.position.x = odom[0];
  od->pose.pose.position.y = odom[1];
  tf2::Quaternion q;
  q.setRPY(0, 0, odom[2]);
  od->pose.pose.orientation = tf2::toMsg(q);
  od->twist.twist.linear.x = odom[3];
  od->twist.twist.linear.y = odom[4];
  od->twist.twist.angular.z = odom[5];

  auto pp = std::make_shared<people_msgs::msg::People>();
  pp->header.frame_id = people_has_tf ? "map" : "odom";
  for (uint32_t i = 0; i < n_people; ++i) {
    const double *r = people + 8 * i;
    people_msgs::msg::Person p;
    p.position.x = r[0];
    p.position.y = r[1];
    p.position.z = r[2];
    p.velocity.x = r[3];
    p.velocity.y = r[4];
    p.velocity.z = r[5];
    p.tags = {std::to_string((int)r[6]), std::to_string((int)r[7])};
    pp->people.push_back(p);
  }

  auto ls = std::make_shared<sensor_msgs::msg::LaserScan>();
  ls->header.frame_id = laser_has_tf ? "laser" : "odom";
  ls->angle_min = angle_min;
  ls->angle_increment = angle_inc;
  ls->ranges.assign(ranges, ranges + n_ranges);

  iface.odomCb(od);
  iface.peopleCb(pp);
  iface.laserCb(ls);
  iface.peopleCb(pp);
  iface.odomCb(od);

  std::vector<sfm::Agent> ag = iface.getAgents();
  if (ag.size() != n_people + 1)
    return -1;
  for (size_t i = 0; i < ag.size(); ++i) {
    double *o = agents_out + 16 * i;
    const sfm::Agent &a = ag[i];
    o[0] = a.position.getX();
    o[1] = a.position.getY();
    o[2] = a.velocity.getX();
    o[3] = a.velocity.getY();
    o[4] = a.yaw.toRadian();
    o[5] = a.linearVelocity;
    o[6] = a.angularVelocity;
    o[7] = a.radius;
    o[8] = a.desiredVelocity;
    o[9] = a.goals.empty() ? 0.0 : a.goals.front().center.getX();
    o[10] = a.goals.empty() ? 0.0 : a.goals.front().center.getY();
    o[11] = a.goals.empty() ? 0.0 : a.goals.front().radius;
    o[12] = (double)a.goals.size();
    o[13] = (double)a.groupId;
    o[14] = i == 0 ? -1.0 : (double)a.id; // agents[0].id is never set by the reference (sensor_interface.cpp:32-37)
    o[15] = (double)a.obstacles1.size();
  }
  const std::vector<utils::Vector2d> &obs = ag[0].obstacles1;
  *n_obstacles = (uint32_t)obs.size();
  for (size_t i = 0; i < obs.size() && i < max_obstacles; ++i) {
    obstacles_out[2 * i] = obs[i].getX();
    obstacles_out[2 * i + 1] = obs[i].getY();
  }
  return 0;
}

} // extern "C"
