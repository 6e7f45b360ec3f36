"""Python handle on the C++ host mirror of the reference's plugin class ``SFWPlannerNode``
(``host/sfw_node_host.{hpp,cpp}`` in ``libsfw_planner_host.so``): one call delivers the sensor callbacks,
``setPlan`` and ``ticks`` x ``computeVelocityCommands`` the way nav2's controller_server would."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import SceneArray, SfwParams, SfwScene
from .planner import host_lib

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
# (params, ext, scene, ranges, n, angle_min, angle_inc, people, n_people, odom, plan, n_plan, plan_has_tf, tf,
#  ticks, cmd_out, status_out, plan_left_out, goal_reached_out)
NODE_ARGTYPES = [C.POINTER(SfwParams), _dp, C.POINTER(SfwScene), C.POINTER(C.c_float), C.c_uint32, C.c_float, C.c_float,
                 _dp, C.c_uint32, _dp, _dp, C.c_uint32, C.c_int, _dp, C.c_uint32, _dp, _ip, _ip, _ip]


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def node_call(fn, params, ext, scene, scan, people, odom, plan, plan_has_tf, tf, ticks, extra=()):
    """Marshal one plugin run (this layout is shared with the test-only reference harness)."""
    sa = SceneArray([scene])
    r = np.ascontiguousarray(scan["ranges"], dtype=np.float32)
    ppl = _d(people).reshape(-1, 8)
    od, pl, tfa, e = _d(odom), _d(plan).reshape(-1, 3), _d(tf), _d(ext)
    cmd = np.zeros((ticks, 3))
    status = np.zeros(ticks, dtype=np.int32)
    left = np.zeros(ticks, dtype=np.int32)
    reached = np.zeros(ticks, dtype=np.int32)
    rc = fn(C.byref(params), e.ctypes.data_as(_dp), sa.ptr(0), r.ctypes.data_as(C.POINTER(C.c_float)), len(r),
            scan["angle_min"], scan["angle_increment"], ppl.ctypes.data_as(_dp), len(ppl), od.ctypes.data_as(_dp),
            pl.ctypes.data_as(_dp), len(pl), 1 if plan_has_tf else 0, tfa.ctypes.data_as(_dp), ticks,
            cmd.ctypes.data_as(_dp), status.ctypes.data_as(_ip), left.ctypes.data_as(_ip), reached.ctypes.data_as(_ip),
            *extra)
    if rc != 0:
        raise RuntimeError(f"plugin run failed with {rc}")
    return cmd, status, left, reached


def node_run(params, ext, scene, scan, people, odom, plan, plan_has_tf=False, tf=(0.0, 0.0, 0.0), ticks=1, device=0):
    """Returns (cmd[ticks, 3], status[ticks] (1 ok / 0 zero twist / -1 PlannerException), poses left in the
    pruned global plan, isGoalReached, kernel launches)."""
    h = host_lib()
    h.sfwn_node_run.restype = C.c_int
    h.sfwn_node_run.argtypes = NODE_ARGTYPES + [C.c_int, C.POINTER(C.c_uint64)]
    launches = C.c_uint64(0)
    out = node_call(h.sfwn_node_run, params, ext, scene, scan, people, odom, plan, plan_has_tf, tf, ticks,
                    extra=(device, C.byref(launches)))
    return out + (int(launches.value),)
