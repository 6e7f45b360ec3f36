"""Whole-plugin scenarios (sensor callbacks + setPlan + computeVelocityCommands) shared by the golden generator
(the reference's own SFWPlannerNode, oracle/_ref) and the GPU test of the host mirror."""
import dataclasses
import math

import numpy as np

import sensor_cases as SC
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.planner import ext_vector

WL = dataclasses.replace(S.WORKLOADS["C0"], steps=40)
ODOM = (0.0, 0.0, 0.0, 0.3, 0.0, 0.05)


def _line(x0, y0, x1, y1, n, yaw=0.0):
    return [(x0 + (x1 - x0) * k / (n - 1), y0 + (y1 - y0) * k / (n - 1), yaw) for k in range(n)]


def _to_frame(plan, tf):
    """Express controller-frame poses in a frame whose transform INTO the controller frame is tf."""
    tx, ty, tyaw = tf
    c, s = math.cos(tyaw), math.sin(tyaw)
    return [(c * (x - tx) + s * (y - ty), -s * (x - tx) + c * (y - ty), yaw - tyaw) for x, y, yaw in plan]


def make(name):
    """-> dict(params, ext, scene, scan, people, odom, plan, plan_has_tf, tf, ticks)"""
    kw = dict(hazards=False)
    ext = {}
    tf = (0.0, 0.0, 0.0)
    plan_has_tf = False
    ticks = 2
    odom = ODOM
    if name == "far_goal":
        plan = _line(-1.0, 0.0, 4.0, 0.5, 26)
    elif name == "far_goal_hazards":
        kw["hazards"] = True
        plan = _line(-0.5, 0.0, 4.0, 0.5, 19)
    elif name == "plan_in_map_frame":
        tf = (3.5, -1.25, 0.9)
        plan = _to_frame(_line(-1.0, 0.0, 4.0, 0.5, 26), tf)
        plan_has_tf = True
    elif name == "pruning_and_costmap_cut":
        # starts 3 m behind the robot, runs 9 m ahead: the head is pruned, the tail lies outside the 10 m costmap
        plan = _line(-3.0, 0.2, 9.0, 0.2, 49)
        ticks = 3
    elif name == "approach_goal":
        plan = _line(0.0, 0.0, 1.0, 0.3, 6, yaw=0.2)
    elif name == "goal_reached":
        plan = [(-1.0, 0.0, 0.0), (0.02, 0.01, 0.02)]
    elif name == "rotate_in_place":
        plan = [(-1.0, 0.0, 0.0), (0.02, 0.01, 1.0)]
    elif name == "all_blocked":
        plan = _line(0.0, 0.0, 3.0, 0.5, 4)
    elif name == "empty_plan":
        plan = []
        ticks = 1
    elif name == "wp_advance":
        plan = [(0.1 * k, 0.02 * k, 0.0) for k in range(40)]
        ext = dict(wp_tolerance=0.8)
        ticks = 3
    else:
        raise KeyError(name)
    scene = S.make_scene(WL, 3, **kw)
    if name == "all_blocked":
        scene.costmap[:] = 254
    scan = SC.make_scan(11, n_beams=360, n_people=5)
    people = SC.people_records(scan, 3)
    return dict(params=WL.params(), ext=ext_vector(**ext), scene=scene, scan=scan, people=people, odom=odom,
                plan=np.array(plan, dtype=np.float64).reshape(-1, 3), plan_has_tf=plan_has_tf, tf=tf, ticks=ticks)


NAMES = ["far_goal", "far_goal_hazards", "plan_in_map_frame", "pruning_and_costmap_cut", "approach_goal", "goal_reached",
         "rotate_in_place", "all_blocked", "empty_plan", "wp_advance"]
