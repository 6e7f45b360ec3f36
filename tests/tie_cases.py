"""Known-answer scenes for the arg-min TIE-BREAKS of findBestAction (reference src/sfw_planner.cpp:344,394-414):
lower cost, then higher linvel, then lower |angvel|, then the later sample; a cost of exactly 10000 only wins
with linvel > 0 (the initial best is (10000, xv = 0, thetav = 0)); no valid sample => findBestAction returns false.

Exact ties need exact symmetry: a free map, the robot at the origin heading along +x, the waypoint on the x axis —
then the rollouts for +w and -w are mirror images (IEEE arithmetic is sign-symmetric), cost(v, +w) == cost(v, -w)
bit for bit, in FP64 and in the GPU's FP32 crowd state alike.  A pedestrian ON the axis moving along it keeps the
symmetry and makes the social work non-zero.  Sample order is the reference's own [0, +s, -s, +2s, -2s, ...]
(:65-85), generalised to any length.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200._abi import PED_DTYPE


def interleaved_angvels(n_half: int, max_vel_th: float = 0.5) -> np.ndarray:
    """[0, +s, -s, +2s, -2s, ...] with n_half magnitudes (the reference ships n_half = 4)."""
    step = max_vel_th / n_half
    out = [0.0]
    for i in range(1, n_half + 1):
        out += [i * step, i * (-step)]
    return np.array(out, dtype=np.float64)


def _free_scene(steps, ped_on_axis=False, blocked=False):
    wl = dataclasses.replace(S.WORKLOADS["C0"], steps=steps, n_peds=0)
    sc = S.make_scene(wl, 0, n_obstacles=0)
    sc.costmap = np.zeros((200, 200), dtype=np.uint8)
    if blocked:
        sc.costmap[:] = 254
    r = list(sc.robot)
    r[6], r[7] = 3.0, 0.0  # waypoint on the x axis
    sc.robot = tuple(r)
    if ped_on_axis:
        peds = np.zeros(1, dtype=PED_DTYPE)
        q = peds[0]
        q["x"], q["y"], q["vx"], q["vy"] = 2.0, 0.0, -0.5, 0.0
        q["goal_x"], q["goal_y"], q["goal_radius"] = 1.0, 0.0, 0.35
        q["desired_velocity"], q["radius"], q["has_goal"], q["group_id"], q["id"] = 1.0, 0.35, 1, -1, 1
        sc.peds = peds
    return wl.params(), sc


def _mirror_pairs(ang):
    """(column of +w, column of -w) pairs of an angvel array."""
    ang = list(ang)
    return [(i, ang.index(-w)) for i, w in enumerate(ang) if w > 0 and -w in ang]


def shipped_5x9():
    p, sc = _free_scene(40)
    lin, _ = S.reference_sample_arrays()
    return p, sc, lin, interleaved_angvels(4)


def shipped_5x9_ped_on_axis():
    p, sc = _free_scene(40, ped_on_axis=True)
    lin, _ = S.reference_sample_arrays()
    return p, sc, lin, interleaved_angvels(4)


def duplicated_rows():
    """Every linvel listed twice: whole rows tie, the later one must win."""
    p, sc = _free_scene(24)
    lin = np.array([0.0, 0.0, 0.35, 0.35, 0.7, 0.7, 0.7], dtype=np.float64)
    return p, sc, lin, interleaved_angvels(3)


def cost_10000_with_linvel():
    """vel_weight 20000, the other weights 0: cost = 20000 |0.7 - v_end| / 0.7 = exactly 10000 for v = 0.35 (reached
    from 0.3 inside the horizon) and 20000 for v = 0.  Every column of row 1 ties at 10000 with linvel > 0: the
    reference takes it (lowest |w|, i.e. w = 0)."""
    p, sc = _free_scene(40)
    p.vel_weight, p.distance_weight, p.angle_weight, p.costmap_weight, p.social_weight = 20000.0, 0.0, 0.0, 0.0, 0.0
    return p, sc, np.array([0.0, 0.35], dtype=np.float64), interleaved_angvels(4)


def cost_10000_linvel_zero():
    """vel_weight 10000 and only v = 0: every scored sample costs exactly 10000 with linvel == 0 — never better
    than the initial best, so findBestAction fails (valid = 0) although every trajectory is legal."""
    p, sc = _free_scene(40)
    p.vel_weight, p.distance_weight, p.angle_weight, p.costmap_weight, p.social_weight = 10000.0, 0.0, 0.0, 0.0, 0.0
    return p, sc, np.array([0.0], dtype=np.float64), interleaved_angvels(4)


def no_zero_w():
    """The shipped sets without w = 0: the best cost is shared by (v, +s) and (v, -s) — same cost, same linvel, same
    |w| — and the LATER sample (-s) wins (cost <= best_cost updates, :394)."""
    p, sc = _free_scene(40, ped_on_axis=True)
    lin, _ = S.reference_sample_arrays()
    return p, sc, lin, interleaved_angvels(4)[1:]


def all_costs_equal():
    """Every weight 0: every legal sample costs exactly 0.  Winner = highest linvel (:397-401), then lowest |w|
    (:403-407), then the later of (+s, -s)."""
    p, sc = _free_scene(24)
    p.vel_weight = p.distance_weight = p.angle_weight = p.costmap_weight = p.social_weight = 0.0
    lin = np.array([0.0, 0.7, 0.35, 0.7, 0.175], dtype=np.float64)  # unsorted on purpose, the maximum twice
    return p, sc, lin, interleaved_angvels(3)[1:]


def all_invalid():
    p, sc = _free_scene(20, blocked=True)
    lin, _ = S.reference_sample_arrays()
    return p, sc, lin, interleaved_angvels(4)


def big_grid():
    """40 x 64 = 2560 samples (several tiles of the thread-per-trajectory kernel, so the tile winners meet in the
    last block's reduction), every linvel duplicated, every |w| present twice and no w = 0: the winner is decided
    by "later row" and "later of (+s, -s)"."""
    p, sc = _free_scene(32, ped_on_axis=True)
    lin = np.repeat(np.array([0.7 * i / 19 for i in range(20)], dtype=np.float64), 2)
    return p, sc, lin, interleaved_angvels(32)[1:]


CASES = {
    "shipped_5x9": shipped_5x9,
    "shipped_5x9_ped_on_axis": shipped_5x9_ped_on_axis,
    "duplicated_rows": duplicated_rows,
    "cost_10000_with_linvel": cost_10000_with_linvel,
    "cost_10000_linvel_zero": cost_10000_linvel_zero,
    "no_zero_w": no_zero_w,
    "all_costs_equal": all_costs_equal,
    "all_invalid": all_invalid,
    "big_grid": big_grid,
}
mirror_pairs = _mirror_pairs
