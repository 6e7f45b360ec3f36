"""Named parity cases shared by tests/golden/make_golden.py (which runs the REFERENCE's own compiled
sources on them, in the build container) and the tests that replay them against the committed outputs."""
from __future__ import annotations

import dataclasses
import zlib

import numpy as np

from social_force_window_planner_b200 import scenes as S


def _case(wl, seed=0, lin_ang=None, params_mut=None, **scene_kw):
    sc = S.make_scene(wl, seed, **scene_kw)
    p = wl.params()
    if params_mut:
        params_mut(p)
    lin, ang = lin_ang if lin_ang is not None else wl.sample_arrays()
    return p, sc, lin, ang


def _yaml_like(p):
    # the shipped config/local_planner.yaml weights + accelerations (reference config/local_planner.yaml:7-32)
    p.max_vel_x = 0.8
    p.max_trans_acc = 0.15
    p.max_rot_acc = 0.52
    p.robot_radius = 0.4
    p.social_weight = 2.0
    p.costmap_weight = 2.0
    p.angle_weight = 0.6
    p.distance_weight = 1.0
    p.vel_weight = 0.8


def _grouped(wl, seed, groups, spread=0.45, **scene_kw):
    """Scene whose pedestrians carry people-message group tags (reference src/sensor_interface.cpp:449):
    ``groups`` = {group_id: [pedestrian indices]}.  Members are pulled next to the group's first member
    (``spread`` metres apart, alternating sides) and walk roughly with it, so that gaze, coherence and
    repulsion (contact at r_a + r_b = 0.7 m) are all exercised.  A lone tag (1 member) must stay inert."""
    p, sc, lin, ang = _case(wl, seed, **scene_kw)
    peds = sc.peds
    for gid, members in groups.items():
        lead = peds[members[0]]
        for n, j in enumerate(members):
            q = peds[j]
            q["group_id"] = gid
            if n == 0:
                continue
            q["x"] = lead["x"] + spread * n
            q["y"] = lead["y"] + 0.3 * (-1) ** n
            q["vx"] = lead["vx"] * (1.0 - 0.1 * n) + 0.05 * n
            q["vy"] = lead["vy"] * (1.0 + 0.07 * n) - 0.04 * n
            q["goal_x"] = q["x"] + 2.0 * q["vx"]
            q["goal_y"] = q["y"] + 2.0 * q["vy"]
    return p, sc, lin, ang


C0 = S.WORKLOADS["C0"]
CASES = {
    "c0_seed0": lambda: _case(C0, 0),
    "c0_seed1": lambda: _case(C0, 1),
    "c0_seed2": lambda: _case(C0, 2),
    "c0_seed3": lambda: _case(C0, 3),
    "c0_hazards_40steps": lambda: _case(dataclasses.replace(C0, steps=40), 0, hazards=True),
    "c0_hazards_seed5": lambda: _case(dataclasses.replace(C0, steps=40, n_peds=8), 5, hazards=True),
    "ref_5x9_samples_40steps": lambda: _case(dataclasses.replace(C0, steps=40), 2,
                                             lin_ang=S.reference_sample_arrays()),
    "ref_5x9_yaml_params": lambda: _case(dataclasses.replace(C0, steps=12), 3,
                                         lin_ang=S.reference_sample_arrays(0.8, 1.57), params_mut=_yaml_like),
    "point_footprint": lambda: _case(C0, 1, footprint=np.zeros((0, 2))),
    "point_footprint_hazards": lambda: _case(dataclasses.replace(C0, steps=40), 1, footprint=np.zeros((0, 2)),
                                             hazards=True),
    "square_footprint": lambda: _case(C0, 2, footprint=np.array([[0.3, 0.25], [-0.3, 0.25], [-0.3, -0.25],
                                                                 [0.3, -0.25]])),
    "no_peds_no_obstacles": lambda: _case(C0, 0, n_peds=0, n_obstacles=0),
    "no_obstacles": lambda: _case(C0, 0, n_obstacles=0),
    "one_ped": lambda: _case(C0, 4, n_peds=1),
    "odom_far_from_origin": lambda: _case(C0, 1, robot_xy=(1234.5, -987.25), robot_theta=0.7),
    "c1_shape_12x12": lambda: _case(dataclasses.replace(S.WORKLOADS["C1"], n_v=12, n_w=12), 0),
    "c3_shape_8x8": lambda: _case(dataclasses.replace(S.WORKLOADS["C3"], n_v=8, n_w=8), 7),
    "c4_shape_8x8": lambda: _case(dataclasses.replace(S.WORKLOADS["C4"], n_v=8, n_w=8), 0),
    # pedestrian groups (lightsfm computeGroupForce through the computeForces call, sfw_planner.cpp:592)
    "groups_pair_and_triple": lambda: _grouped(dataclasses.replace(C0, n_peds=7, steps=40), 2,
                                               {3: [0, 1], 7: [2, 3, 4], 9: [5]}),
    "groups_tight_contact": lambda: _grouped(dataclasses.replace(C0, n_peds=6, steps=32), 4,
                                             {0: [0, 2, 4, 5]}, spread=0.3),
    "groups_c1_shape_10x10": lambda: _grouped(dataclasses.replace(S.WORKLOADS["C1"], n_v=10, n_w=10), 1,
                                              {1: [0, 1, 2], 2: [5, 9], 4: [10, 11, 12, 13, 14]}),
}


def scene_crc(sc) -> int:
    """Checksum of everything the scorer reads from a scene: pins the scene generator itself."""
    c = zlib.crc32(np.ascontiguousarray(sc.costmap).tobytes())
    c = zlib.crc32(np.ascontiguousarray(sc.peds).tobytes(), c)
    c = zlib.crc32(np.ascontiguousarray(sc.obstacles, dtype=np.float64).tobytes(), c)
    c = zlib.crc32(np.ascontiguousarray(sc.footprint, dtype=np.float64).tobytes(), c)
    c = zlib.crc32(np.array(list(sc.robot) + [sc.resolution, sc.origin_x, sc.origin_y], dtype=np.float64).tobytes(), c)
    return c


# footprint known-answer poses (x, y, theta) evaluated on the "c0_hazards_40steps" scene
FOOTPRINT_POSES = [(0.0, 0.0, 0.0), (0.0, 0.0, 0.3), (0.2, 0.1, 0.3), (0.3, 0.2, -1.1), (0.4, -0.45, 2.0),
                   (4.9, 0.0, 0.3), (-4.7, 4.7, 1.0), (4.99, 4.99, 0.0), (-5.2, 0.0, 0.0), (0.1, -0.5, 0.0),
                   (1.0, 1.0, 0.77), (-2.0, 3.0, -2.5), (0.55, 0.2, 0.0), (0.7, 0.2, 0.0), (0.6, 0.45, 0.5),
                   (0.35, -0.55, 0.0)]
