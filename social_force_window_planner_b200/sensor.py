"""Python handle on the C++ host mirror of the reference's ``SFMSensorInterface``
(``host/sfw_sensor_host.{hpp,cpp}`` in ``libsfw_planner_host.so``): one call replays the callback sequence
ROS would deliver (odom, people, laser, people, odom) and returns the agent snapshot ``getAgents()`` hands
the planner.  The laser filter runs on the GPU (``sfw_laser_obstacles``); there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .planner import host_lib

_dp = C.POINTER(C.c_double)
# agent columns of the returned array (same as oracle/ref_sensor_harness.cpp)
AGENT_COLS = ("x", "y", "vx", "vy", "yaw", "linear_velocity", "angular_velocity", "radius", "desired_velocity",
              "goal_x", "goal_y", "goal_radius", "n_goals", "group_id", "id", "n_obstacles")
DEFAULT_PARAMS = (3.0, 0.35, 2.0, 1.0, 0.35, 0.7)  # InterfaceParams defaults (sensor_interface.hpp:62-65)


def sensor_run(scan: dict, people, odom, params=DEFAULT_PARAMS, people_has_tf=False, device=0):
    """Returns (agents float64[n + 1, 16], obstacle points float64[m, 2], kernel launches)."""
    h = host_lib()
    h.sfws_sensor_run.restype = C.c_int
    h.sfws_sensor_run.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_float, C.c_float, C.c_int, _dp, C.c_uint32,
                                  C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_uint32, C.POINTER(C.c_uint32), C.c_int,
                                  C.POINTER(C.c_uint64)]
    r = np.ascontiguousarray(scan["ranges"], dtype=np.float32)
    ppl = np.ascontiguousarray(people, dtype=np.float64).reshape(-1, 8)
    od = np.ascontiguousarray(odom, dtype=np.float64)
    pr = np.ascontiguousarray(params, dtype=np.float64)
    tf = np.ascontiguousarray(scan.get("tf") or (0.0, 0.0, 0.0), dtype=np.float64)
    agents = np.zeros((len(ppl) + 1, 16), dtype=np.float64)
    obs = np.zeros((max(len(r), 1), 2), dtype=np.float64)
    n = C.c_uint32(0)
    launches = C.c_uint64(0)
    rc = h.sfws_sensor_run(r.ctypes.data_as(C.POINTER(C.c_float)), len(r), scan["angle_min"], scan["angle_increment"],
                           1 if scan.get("tf") is not None else 0, ppl.ctypes.data_as(_dp), len(ppl),
                           1 if people_has_tf else 0, od.ctypes.data_as(_dp), pr.ctypes.data_as(_dp),
                           tf.ctypes.data_as(_dp), agents.ctypes.data_as(_dp), obs.ctypes.data_as(_dp), len(obs),
                           C.byref(n), device, C.byref(launches))
    if rc != 0:
        raise RuntimeError(f"sfws_sensor_run failed with {rc}")
    return agents, obs[:n.value].copy(), int(launches.value)
