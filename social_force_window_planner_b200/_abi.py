"""ctypes mirror of ``include/sfw_b200.h`` (the C ABI of the scorer).

Field order and types must match the header exactly; ``tests/test_abi_cpu.py`` checks the struct sizes
against the values the C library reports.  The structs follow the reference's own data:
``SfwParams`` = ControllerParams fields read on the path (reference
include/social_force_window_planner/sfw_planner.hpp:55-227), ``SfwPed`` = one sfm::Agent as
SFMSensorInterface::peopleCb builds it (reference src/sensor_interface.cpp:447-504).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

SFW_OK = 0
SFW_ERR_ARG = -1
SFW_ERR_CUDA = -2
SFW_ERR_UNSUPPORTED = -3
SFW_ERR_STATE = -4
SFW_COST_INVALID = -1.0
SFW_COST_SKIPPED = -2.0


class SfwParams(C.Structure):
    _fields_ = [
        ("max_vel_x", C.c_double),
        ("max_trans_acc", C.c_double),
        ("max_rot_acc", C.c_double),
        ("sim_time", C.c_double),
        ("sim_granularity", C.c_double),
        ("robot_radius", C.c_float),
        ("reserved0", C.c_float),
        ("social_weight", C.c_double),
        ("costmap_weight", C.c_double),
        ("angle_weight", C.c_double),
        ("distance_weight", C.c_double),
        ("vel_weight", C.c_double),
    ]


class SfwSfmParams(C.Structure):
    _fields_ = [
        ("force_factor_desired", C.c_double),
        ("force_factor_obstacle", C.c_double),
        ("force_sigma_obstacle", C.c_double),
        ("force_factor_social", C.c_double),
        ("force_factor_group_gaze", C.c_double),
        ("force_factor_group_coherence", C.c_double),
        ("force_factor_group_repulsion", C.c_double),
        ("lambda_", C.c_double),
        ("gamma", C.c_double),
        ("n", C.c_double),
        ("n_prime", C.c_double),
        ("relaxation_time", C.c_double),
    ]


class SfwRobot(C.Structure):
    _fields_ = [
        ("x", C.c_double), ("y", C.c_double), ("theta", C.c_double),
        ("vx", C.c_double), ("vy", C.c_double), ("vtheta", C.c_double),
        ("wpx", C.c_double), ("wpy", C.c_double),
        ("agent_x", C.c_double), ("agent_y", C.c_double),
        ("agent_vx", C.c_double), ("agent_vy", C.c_double),
        ("agent_radius", C.c_double),
    ]


class SfwPed(C.Structure):
    _fields_ = [
        ("x", C.c_double), ("y", C.c_double),
        ("vx", C.c_double), ("vy", C.c_double),
        ("goal_x", C.c_double), ("goal_y", C.c_double),
        ("goal_radius", C.c_double),
        ("desired_velocity", C.c_double),
        ("radius", C.c_double),
        ("has_goal", C.c_int32),
        ("group_id", C.c_int32),
        ("id", C.c_int32),
        ("reserved0", C.c_int32),
    ]


# numpy view of SfwPed so whole pedestrian arrays can be filled vectorised
PED_DTYPE = np.dtype(
    [("x", "f8"), ("y", "f8"), ("vx", "f8"), ("vy", "f8"), ("goal_x", "f8"), ("goal_y", "f8"),
     ("goal_radius", "f8"), ("desired_velocity", "f8"), ("radius", "f8"), ("has_goal", "i4"),
     ("group_id", "i4"), ("id", "i4"), ("reserved0", "i4")]
)
assert PED_DTYPE.itemsize == C.sizeof(SfwPed)


class SfwScene(C.Structure):
    _fields_ = [
        ("robot", SfwRobot),
        ("costmap", C.POINTER(C.c_uint8)),
        ("size_x", C.c_uint32), ("size_y", C.c_uint32),
        ("resolution", C.c_double), ("origin_x", C.c_double), ("origin_y", C.c_double),
        ("peds", C.POINTER(SfwPed)),
        ("n_peds", C.c_uint32),
        ("n_obstacles", C.c_uint32),
        ("obstacles_xy", C.POINTER(C.c_double)),
        ("footprint_xy", C.POINTER(C.c_double)),
        ("n_footprint", C.c_uint32),
        ("reserved0", C.c_uint32),
    ]


class SfwLaserScan(C.Structure):
    """One sensor_msgs/LaserScan as SFMSensorInterface::laserCb reads it (reference
    src/sensor_interface.cpp:103-229) plus the people it is filtered against."""
    _fields_ = [
        ("ranges", C.POINTER(C.c_float)),
        ("n_ranges", C.c_uint32),
        ("angle_min", C.c_float), ("angle_increment", C.c_float),
        ("has_tf", C.c_int32),
        ("tf_x", C.c_double), ("tf_y", C.c_double), ("tf_yaw", C.c_double),
        ("people_xy", C.POINTER(C.c_double)),
        ("n_people", C.c_uint32),
        ("reserved0", C.c_uint32),
    ]


class SfwBest(C.Structure):
    _fields_ = [
        ("valid", C.c_int32),
        ("index", C.c_uint32),
        ("cost", C.c_float),
        ("reserved0", C.c_float),
        ("v", C.c_double),
        ("w", C.c_double),
    ]


BEST_DTYPE = np.dtype(
    [("valid", "i4"), ("index", "u4"), ("cost", "f4"), ("reserved0", "f4"), ("v", "f8"), ("w", "f8")]
)
assert BEST_DTYPE.itemsize == C.sizeof(SfwBest)


class SfwLimits(C.Structure):
    _fields_ = [
        ("max_scenes", C.c_uint32),
        ("max_samples", C.c_uint32),
        ("max_peds", C.c_uint32),
        ("max_obstacles", C.c_uint32),
        ("max_cells", C.c_uint32),
    ]


def default_params() -> SfwParams:
    """ControllerParams header defaults (reference sfw_planner.hpp:56-66)."""
    p = SfwParams()
    p.max_vel_x = 0.7
    p.max_trans_acc = 1.0
    p.max_rot_acc = 1.0
    p.sim_time = 1.0
    p.sim_granularity = 0.025
    p.robot_radius = 0.35
    p.social_weight = 1.2
    p.costmap_weight = 2.0
    p.angle_weight = 0.7
    p.distance_weight = 1.0
    p.vel_weight = 1.0
    return p


def default_sfm_params() -> SfwSfmParams:
    """lightsfm sfm::Parameters defaults."""
    s = SfwSfmParams()
    (s.force_factor_desired, s.force_factor_obstacle, s.force_sigma_obstacle, s.force_factor_social,
     s.force_factor_group_gaze, s.force_factor_group_coherence, s.force_factor_group_repulsion,
     s.lambda_, s.gamma, s.n, s.n_prime, s.relaxation_time) = (
        2.0, 10.0, 0.2, 2.1, 3.0, 2.0, 1.0, 2.0, 0.35, 2.0, 3.0, 0.5)
    return s


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class SceneArray:
    """A contiguous ``SfwScene[n]`` built from python :class:`~.scenes.Scene` objects.

    Keeps the numpy buffers alive for as long as the array is referenced.
    """

    def __init__(self, scenes):
        self.scenes = list(scenes)
        self.n = len(self.scenes)
        self.array = (SfwScene * self.n)()
        self._keep = []
        for k, s in enumerate(self.scenes):
            sc = self.array[k]
            r = sc.robot
            (r.x, r.y, r.theta, r.vx, r.vy, r.vtheta, r.wpx, r.wpy, r.agent_x, r.agent_y, r.agent_vx,
             r.agent_vy, r.agent_radius) = [float(v) for v in s.robot]
            cm = np.ascontiguousarray(s.costmap, dtype=np.uint8)
            peds = np.ascontiguousarray(s.peds, dtype=PED_DTYPE)
            obs = np.ascontiguousarray(s.obstacles, dtype=np.float64).reshape(-1)
            fp = np.ascontiguousarray(s.footprint, dtype=np.float64).reshape(-1)
            self._keep.append((cm, peds, obs, fp))
            sc.costmap = cm.ctypes.data_as(C.POINTER(C.c_uint8))
            sc.size_x = cm.shape[1]
            sc.size_y = cm.shape[0]
            sc.resolution = float(s.resolution)
            sc.origin_x = float(s.origin_x)
            sc.origin_y = float(s.origin_y)
            sc.peds = peds.ctypes.data_as(C.POINTER(SfwPed))
            sc.n_peds = len(peds)
            sc.n_obstacles = len(obs) // 2
            sc.obstacles_xy = _dptr(obs)
            sc.footprint_xy = _dptr(fp)
            sc.n_footprint = len(fp) // 2

    def __len__(self):
        return self.n

    def ptr(self, k: int = 0):
        return C.cast(C.byref(self.array[k]), C.POINTER(SfwScene))
