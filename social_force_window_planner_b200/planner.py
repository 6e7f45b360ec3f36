"""Python handle on the C++ host mirror of the reference's ``SFWPlanner``
(``host/sfw_planner_host.{hpp,cpp}`` -> ``libsfw_planner_host.so``).

Same call sequence as ``SFWPlannerNode`` drives in the reference (src/sfw_planner_node.cpp:277-284):
``updatePlan(plan)`` then ``findBestAction(pose, vel)`` every tick.  All scoring happens in the CUDA
library; constructing a planner does not touch the GPU (the context is created on the first scored tick).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._abi import SceneArray, SfwParams, SfwScene

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, "libsfw_planner_host.so")
_dp = C.POINTER(C.c_double)
_host = None


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        _lib.lib()  # libsfw_b200.so first (the host library links against it)
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError(f"{HOST_LIB_PATH} is missing: run `python -m social_force_window_planner_b200.build`")
        h = C.CDLL(HOST_LIB_PATH)
        h.sfwh_create.restype = C.c_void_p
        h.sfwh_create.argtypes = [C.POINTER(SfwParams), _dp, C.POINTER(SfwScene), C.c_int]
        h.sfwh_destroy.argtypes = [C.c_void_p]
        h.sfwh_set_samples.argtypes = [C.c_void_p, _dp, C.c_uint32, _dp, C.c_uint32]
        h.sfwh_get_samples.argtypes = [C.c_void_p, _dp, _dp]
        h.sfwh_update_plan.argtypes = [C.c_void_p, _dp, C.c_uint32]
        h.sfwh_find_best_action.restype = C.c_int
        h.sfwh_find_best_action.argtypes = [C.c_void_p, _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        h.sfwh_is_goal_reached.restype = C.c_int
        h.sfwh_is_goal_reached.argtypes = [C.c_void_p]
        h.sfwh_last_error.restype = C.c_char_p
        h.sfwh_last_error.argtypes = [C.c_void_p]
        _host = h
    return _host


# ext vector of the C wrapper: ControllerParams fields the scorer itself does not need
EXT_DEFAULTS = dict(min_vel_x=0.1, max_vel_th=0.5, min_vel_th=0.1, min_in_place_vel_th=0.3, yaw_goal_tolerance=0.05,
                    xy_goal_tolerance=0.1, wp_tolerance=0.5, is_circular=1.0)


def ext_vector(**kw) -> np.ndarray:
    d = dict(EXT_DEFAULTS)
    d.update(kw)
    return np.array([d[k] for k in EXT_DEFAULTS], dtype=np.float64)


class SFWPlanner:
    """``social_force_window_planner::SFWPlanner`` (host mirror) for one scene's costmap / agents."""

    def __init__(self, params: SfwParams, scene, device: int = 0, **ext):
        self._h = host_lib()
        self._sa = scene if isinstance(scene, SceneArray) else SceneArray([scene])
        self._ext = ext_vector(**ext)
        self._p = self._h.sfwh_create(C.byref(params), self._ext.ctypes.data_as(_dp), self._sa.ptr(0), device)
        self.wp_index = -1
        self.running = False
        self.best_index = -1
        self.kernel_launches = 0

    def close(self):
        if getattr(self, "_p", None):
            self._h.sfwh_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setSampleSets(self, linvels, angvels):
        lin = np.ascontiguousarray(linvels, dtype=np.float64)
        ang = np.ascontiguousarray(angvels, dtype=np.float64)
        self._h.sfwh_set_samples(self._p, lin.ctypes.data_as(_dp), len(lin), ang.ctypes.data_as(_dp), len(ang))

    def defaultSampleSets(self):
        lin, ang = np.zeros(5), np.zeros(9)
        self._h.sfwh_get_samples(self._p, lin.ctypes.data_as(_dp), ang.ctypes.data_as(_dp))
        return lin, ang

    def updatePlan(self, plan_xyt):
        plan = np.ascontiguousarray(plan_xyt, dtype=np.float64).reshape(-1, 3)
        self._h.sfwh_update_plan(self._p, plan.ctypes.data_as(_dp), len(plan))

    def findBestAction(self, pose_xyt, vel_xyt):
        """Returns (ok, (linear.x, linear.y, angular.z))."""
        pose = np.ascontiguousarray(pose_xyt, dtype=np.float64)
        vel = np.ascontiguousarray(vel_xyt, dtype=np.float64)
        cmd = np.zeros(3)
        wp, run, bi, nl = C.c_int(0), C.c_int(0), C.c_int(0), C.c_uint64(0)
        ok = self._h.sfwh_find_best_action(self._p, pose.ctypes.data_as(_dp), vel.ctypes.data_as(_dp),
                                           cmd.ctypes.data_as(_dp), C.byref(wp), C.byref(run), C.byref(bi), C.byref(nl))
        self.wp_index, self.running, self.best_index, self.kernel_launches = wp.value, bool(run.value), bi.value, nl.value
        return bool(ok), tuple(cmd)

    def getMarkers(self, n_samples: int, max_points: int = 128):
        """The MarkerArray of the last grid tick (reference getMarkers(), src/sfw_planner.cpp:112-114):
        (rgba float32[n, 4], n_points uint32[n], xyz float64[n, max_points, 3]) or None before a grid tick."""
        h = self._h
        h.sfwh_get_markers.restype = C.c_int
        h.sfwh_get_markers.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32), _dp, C.c_uint32]
        rgba = np.zeros((n_samples, 4), dtype=np.float32)
        npts = np.zeros(n_samples, dtype=np.uint32)
        xyz = np.zeros((n_samples, max_points, 3), dtype=np.float64)
        n = h.sfwh_get_markers(self._p, rgba.ctypes.data_as(C.POINTER(C.c_float)),
                               npts.ctypes.data_as(C.POINTER(C.c_uint32)), xyz.ctypes.data_as(_dp), max_points)
        if n < 0:
            return None
        assert n == n_samples
        return rgba, npts, xyz

    def isGoalReached(self) -> bool:
        return bool(self._h.sfwh_is_goal_reached(self._p))

    @property
    def last_error(self) -> str:
        return self._h.sfwh_last_error(self._p).decode()
