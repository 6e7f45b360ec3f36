"""findBestAction scenarios shared by the CPU and GPU host-planner tests: each is run through the
reference's own findBestAction (oracle/_ref) and through the host mirror on identical inputs."""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

import oracle_lib as ol
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200._abi import SceneArray
from social_force_window_planner_b200.planner import SFWPlanner, ext_vector

_dp = C.POINTER(C.c_double)
WL = dataclasses.replace(S.WORKLOADS["C0"], steps=40)  # header defaults: sim_time 1.0 / 0.025


def make(seed=0, **kw):
    sc = S.make_scene(WL, seed, **kw)
    return WL.params(), sc


# name -> (scene kwargs, plan [(x, y, yaw)...], ext overrides, needs_gpu)
CASES = {
    "empty_plan_not_running": (dict(), [], {}, False),
    "goal_reached": (dict(), [(-1.0, 0.0, 0.0), (0.02, 0.01, 0.02)], {}, False),
    "rotate_in_place_circular_pos": (dict(), [(-1.0, 0.0, 0.0), (0.02, 0.01, 1.0)], {}, False),
    "rotate_in_place_circular_neg": (dict(), [(-1.0, 0.0, 0.0), (0.02, 0.01, -2.5)], {}, False),
    "rotate_in_place_polygon_ok": (dict(), [(-1.0, 0.0, 0.0), (0.02, 0.01, 1.0)], dict(is_circular=0.0), True),
    # a long rectangular base whose front corner already sits in the lethal box: rotation must be refused
    "rotate_in_place_polygon_blocked": (dict(hazards=True, footprint=np.array([[0.7, 0.2], [-0.3, 0.2], [-0.3, -0.2],
                                                                               [0.7, -0.2]])),
                                        [(0.3, 0.2, 0.0), (0.32, 0.22, 1.0)],
                                        dict(is_circular=0.0, xy_goal_tolerance=0.5), True),
    "approach_goal": (dict(), [(0.0, 0.0, 0.0), (0.5, 0.1, 0.0), (1.0, 0.3, 0.2)], {}, True),
    "approach_goal_blocked_falls_to_grid": (dict(hazards=True), [(0.0, 0.0, 0.0), (0.9, 0.2, 0.0)], {}, True),
    "grid_far_goal": (dict(), [(0.0, 0.0, 0.0), (1.0, 0.2, 0.0), (2.0, 0.4, 0.0), (3.0, 0.5, 0.0), (4.0, 0.5, 0.0)], {},
                      True),
    "grid_waypoint_advance": (dict(), [(0.1 * k, 0.02 * k, 0.0) for k in range(40)], dict(wp_tolerance=0.8), True),
    "grid_new_plan_closest_point": (dict(), [(-3.0 + 0.25 * k, 1.0, 0.0) for k in range(30)], dict(wp_tolerance=0.3), True),
    "grid_hazards": (dict(hazards=True), [(0.0, 0.0, 0.0), (3.0, 0.5, 0.0), (4.0, 0.5, 0.0)], {}, True),
    "grid_all_blocked": (dict(hazards=True, robot_xy=(0.0, 0.0)), [(0.0, 0.0, 0.0), (3.0, 0.5, 0.0)],
                         dict(), True),
}


def run_reference(name):
    kw, plan, ext, _ = CASES[name]
    p, sc = make(**kw)
    if name == "grid_all_blocked":
        sc.costmap[:] = 254
    sa = SceneArray([sc])
    plan_a = np.ascontiguousarray(plan, dtype=np.float64).reshape(-1, 3)
    cmd = np.zeros(3)
    wp, run = C.c_int(0), C.c_int(0)
    e = ext_vector(**ext)
    ok = ol.ref().sfw_ref_find_best_action(C.byref(p), e.ctypes.data_as(_dp), None, sa.ptr(0),
                                           plan_a.ctypes.data_as(_dp), len(plan_a), None, 0, None, 0,
                                           cmd.ctypes.data_as(_dp), C.byref(wp), C.byref(run))
    return bool(ok), tuple(cmd), wp.value, bool(run.value)


def run_host(name):
    kw, plan, ext, _ = CASES[name]
    p, sc = make(**kw)
    if name == "grid_all_blocked":
        sc.costmap[:] = 254
    pl = SFWPlanner(p, sc, **ext)
    try:
        pl.updatePlan(plan)
        r = sc.robot
        ok, cmd = pl.findBestAction((r[0], r[1], r[2]), (r[3], r[4], r[5]))
        return ok, cmd, pl.wp_index, pl.running, pl.kernel_launches, pl.last_error
    finally:
        pl.close()
