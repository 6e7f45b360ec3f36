"""ctypes access to the TEST-ONLY checkers: oracle/libsfw_oracle.so (C restatement) and
oracle/_ref/libsfw_ref.so (the reference's own sources compiled unmodified).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from social_force_window_planner_b200._abi import (SceneArray, SfwBest, SfwLaserScan, SfwParams, SfwScene,
                                                   SfwSfmParams)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsfw_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsfw_ref.so")

_dp = C.POINTER(C.c_double)


class SfwOracleMargins(C.Structure):
    _fields_ = [("goal", C.c_double), ("collision", C.c_double), ("theta", C.c_double),
                ("cell", C.c_double)]


MARGIN_DTYPE = np.dtype([("goal", "f8"), ("collision", "f8"), ("theta", "f8"), ("cell", "f8")])

# oracle/sfw_oracle.h: SfwOracleEvent / SfwOracleProbe (the branch probe)
EVENT_DTYPE = np.dtype([("kind", "i4"), ("step", "i4"), ("a", "i4"), ("b", "i4"), ("decision", "i4"),
                        ("reserved0", "i4"), ("margin", "f8"), ("weight", "f8")])
EV_GOAL, EV_COLLISION, EV_THETA, EV_GROUP = 1, 2, 3, 4


class SfwOracleProbe(C.Structure):
    _fields_ = [("goal_margin", C.c_double), ("collision_margin", C.c_double), ("theta_margin", C.c_double),
                ("theta_min_weight", C.c_double), ("flips", C.c_void_p), ("n_flips", C.c_uint32),
                ("max_events", C.c_uint32), ("events", C.c_void_p), ("n_events", C.c_uint32),
                ("reserved0", C.c_uint32)]


def build(ref: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile (the C oracle always; _ref only where
    /root/reference exists — elsewhere the prebuilt .so that travelled with the repo is kept)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "liboracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


def _ensure_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "sfw_oracle.c")):
        build(ref=False)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        _ensure_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.sfw_oracle_score.restype = C.c_int
        lib.sfw_oracle_score.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwSfmParams),
                                         C.POINTER(SfwScene), _dp, C.c_uint32, _dp, C.c_uint32, _dp,
                                         C.POINTER(SfwBest), C.POINTER(SfwOracleMargins)]
        lib.sfw_oracle_score_mt.restype = C.c_int
        lib.sfw_oracle_score_mt.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwSfmParams),
                                            C.POINTER(SfwScene), _dp, C.c_uint32, _dp, C.c_uint32,
                                            _dp, C.POINTER(SfwBest), C.c_int]
        lib.sfw_oracle_score_trajectory.restype = C.c_double
        lib.sfw_oracle_score_trajectory.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwSfmParams),
                                                    C.POINTER(SfwScene)] + [C.c_double] * 6 + [
            _dp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(SfwOracleMargins)]
        lib.sfw_oracle_score_trajectory_probe.restype = C.c_double
        lib.sfw_oracle_score_trajectory_probe.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwSfmParams),
                                                          C.POINTER(SfwScene)] + [C.c_double] * 6 + [
            C.POINTER(SfwOracleProbe)]
        lib.sfw_oracle_score_probe.restype = C.c_int
        lib.sfw_oracle_score_probe.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwSfmParams), C.POINTER(SfwScene),
                                               _dp, C.c_uint32, _dp, C.c_uint32, C.c_uint32, C.c_uint32,
                                               C.POINTER(SfwOracleProbe), _dp, C.c_void_p,
                                               C.POINTER(C.c_uint32), C.c_int]
        lib.sfw_oracle_argmin.restype = None
        lib.sfw_oracle_argmin.argtypes = [_dp, _dp, C.c_uint32, _dp, C.c_uint32, C.POINTER(SfwBest)]
        lib.sfw_oracle_footprint_cost.restype = C.c_double
        lib.sfw_oracle_footprint_cost.argtypes = [C.POINTER(SfwScene), C.c_double, C.c_double,
                                                  C.c_double, _dp]
        lib.sfw_oracle_line_cells.restype = C.c_int
        lib.sfw_oracle_line_cells.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int), C.c_int]
        lib.sfw_oracle_pair_force.restype = None
        lib.sfw_oracle_pair_force.argtypes = [C.POINTER(SfwSfmParams), _dp, _dp, _dp, _dp]
        lib.sfw_oracle_obstacle_force.restype = None
        lib.sfw_oracle_obstacle_force.argtypes = [C.POINTER(SfwSfmParams), C.c_double, C.c_double,
                                                  C.c_double, _dp, C.c_uint32, _dp]
        lib.sfw_oracle_laser_obstacles.restype = C.c_uint32
        lib.sfw_oracle_laser_obstacles.argtypes = [C.POINTER(SfwLaserScan), C.c_float, C.c_float, _dp]
        _oracle = lib
    return _oracle


def oracle_laser_obstacles(scan: dict, max_obstacle_dist=3.0, person_radius=0.35):
    """SFMSensorInterface::laserCb restated (oracle/sfw_oracle.c): obstacle points float64[m, 2] of one scan."""
    r = np.ascontiguousarray(scan["ranges"], dtype=np.float32)
    ppl = np.ascontiguousarray(scan.get("people", np.zeros((0, 2))), dtype=np.float64).reshape(-1, 2)
    a = SfwLaserScan()
    a.ranges = r.ctypes.data_as(C.POINTER(C.c_float))
    a.n_ranges = len(r)
    a.angle_min = scan["angle_min"]
    a.angle_increment = scan["angle_increment"]
    tf = scan.get("tf")
    a.has_tf = 1 if tf is not None else 0
    if tf is not None:
        a.tf_x, a.tf_y, a.tf_yaw = tf
    a.people_xy = ppl.ctypes.data_as(_dp)
    a.n_people = len(ppl)
    out = np.zeros((max(len(r), 1), 2), dtype=np.float64)
    n = oracle().sfw_oracle_laser_obstacles(C.byref(a), max_obstacle_dist, person_radius, out.ctypes.data_as(_dp))
    return out[:n].copy()


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.sfw_ref_score.restype = C.c_int
        lib.sfw_ref_score.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwSfmParams),
                                      C.POINTER(SfwScene), _dp, C.c_uint32, _dp, C.c_uint32, _dp,
                                      C.POINTER(SfwBest)]
        lib.sfw_ref_score_trajectory.restype = C.c_double
        lib.sfw_ref_score_trajectory.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwSfmParams),
                                                 C.POINTER(SfwScene)] + [C.c_double] * 6 + [
            _dp, C.c_uint32, C.POINTER(C.c_uint32)]
        lib.sfw_ref_footprint_cost.restype = C.c_double
        lib.sfw_ref_footprint_cost.argtypes = [C.POINTER(SfwScene), C.c_double, C.c_double, C.c_double]
        lib.sfw_ref_find_best_action.restype = C.c_int
        lib.sfw_ref_find_best_action.argtypes = [C.POINTER(SfwParams), _dp, C.POINTER(SfwSfmParams),
                                                 C.POINTER(SfwScene), _dp, C.c_uint32, _dp, C.c_uint32,
                                                 _dp, C.c_uint32, _dp, C.POINTER(C.c_int),
                                                 C.POINTER(C.c_int)]
        lib.sfw_ref_markers.restype = C.c_int
        lib.sfw_ref_markers.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwSfmParams), C.POINTER(SfwScene), _dp,
                                        C.c_uint32, _dp, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                        _dp, C.c_uint32]
        lib.sfw_ref_may_i_stop.restype = C.c_int
        lib.sfw_ref_may_i_stop.argtypes = [C.POINTER(SfwParams), C.POINTER(SfwScene)] + [C.c_double] * 7
        lib.sfw_ref_default_samples.restype = C.c_int
        lib.sfw_ref_default_samples.argtypes = [C.c_double, C.c_double, _dp, _dp]
        _ref = lib
    return _ref


REF_SENSOR_SO = os.path.join(ORACLE_DIR, "_ref", "libsfw_ref_sensor.so")
_ref_sensor = None


def have_ref_sensor() -> bool:
    return os.path.exists(REF_SENSOR_SO)


def ref_sensor_run(scan: dict, people: np.ndarray, odom, params=(3.0, 0.35, 2.0, 1.0, 0.35, 0.7), people_has_tf=False):
    """The reference's OWN SFMSensorInterface (src/sensor_interface.cpp compiled unmodified, oracle/_ref) on one
    laser / people / odometry message set.  ``people``: float64[n, 8] rows {x, y, yaw, vx, vy, wz, id, group}
    in the PEOPLE message frame; ``odom`` = (x, y, yaw, vx, vy, wz).  Returns (agents float64[n + 1, 16],
    obstacle points float64[m, 2]) — see oracle/ref_sensor_harness.cpp for the agent columns."""
    global _ref_sensor
    if _ref_sensor is None:
        lib = C.CDLL(REF_SENSOR_SO)
        lib.sfw_ref_sensor_run.restype = C.c_int
        lib.sfw_ref_sensor_run.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_float, C.c_float, C.c_int, _dp,
                                           C.c_uint32, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_uint32,
                                           C.POINTER(C.c_uint32)]
        _ref_sensor = lib
    r = np.ascontiguousarray(scan["ranges"], dtype=np.float32)
    ppl = np.ascontiguousarray(people, dtype=np.float64).reshape(-1, 8)
    od = np.ascontiguousarray(odom, dtype=np.float64)
    pr = np.ascontiguousarray(params, dtype=np.float64)
    tf = np.ascontiguousarray(scan.get("tf") or (0.0, 0.0, 0.0), dtype=np.float64)
    agents = np.zeros((len(ppl) + 1, 16), dtype=np.float64)
    obs = np.zeros((max(len(r), 1), 2), dtype=np.float64)
    n = C.c_uint32(0)
    rc = _ref_sensor.sfw_ref_sensor_run(r.ctypes.data_as(C.POINTER(C.c_float)), len(r), scan["angle_min"],
                                        scan["angle_increment"], 1 if scan.get("tf") is not None else 0,
                                        ppl.ctypes.data_as(_dp), len(ppl), 1 if people_has_tf else 0,
                                        od.ctypes.data_as(_dp), pr.ctypes.data_as(_dp), tf.ctypes.data_as(_dp),
                                        agents.ctypes.data_as(_dp), obs.ctypes.data_as(_dp), len(obs), C.byref(n))
    assert rc == 0
    return agents, obs[:n.value].copy()


REF_NODE_SO = os.path.join(ORACLE_DIR, "_ref", "libsfw_ref_node.so")
_ref_node = None


def have_ref_node() -> bool:
    return os.path.exists(REF_NODE_SO)


from social_force_window_planner_b200.node import NODE_ARGTYPES, node_call  # noqa: E402  (argument layout only)


def ref_node_run(params, ext, scene, scan, people, odom, plan, plan_has_tf=False, tf=(0.0, 0.0, 0.0), ticks=1):
    """The reference's WHOLE plugin (SFWPlannerNode + planner + sensor interface, compiled unmodified) for
    `ticks` control ticks: (cmd[ticks, 3], status[ticks], poses left in the pruned plan, isGoalReached)."""
    global _ref_node
    if _ref_node is None:
        lib = C.CDLL(REF_NODE_SO)
        lib.sfw_ref_node_run.restype = C.c_int
        lib.sfw_ref_node_run.argtypes = NODE_ARGTYPES
        _ref_node = lib
    return node_call(_ref_node.sfw_ref_node_run, params, ext, scene, scan, people, odom, plan, plan_has_tf, tf, ticks)


def ref_markers(params, scene, linvels, angvels, max_points=128, sfm=None):
    """MarkerArray the reference's findBestAction leaves after one grid tick: (ok, rgba[n,4], npts[n], xyz[n,max,3])."""
    sa = SceneArray([scene])
    lin, ang = _d(linvels), _d(angvels)
    n = len(lin) * len(ang)
    rgba = np.zeros((n, 4), dtype=np.float32)
    npts = np.zeros(n, dtype=np.uint32)
    xyz = np.zeros((n, max_points, 3), dtype=np.float64)
    ok = ref().sfw_ref_markers(C.byref(params), C.byref(sfm) if sfm else None, sa.ptr(0), lin.ctypes.data_as(_dp),
                               len(lin), ang.ctypes.data_as(_dp), len(ang), rgba.ctypes.data_as(C.POINTER(C.c_float)),
                               npts.ctypes.data_as(C.POINTER(C.c_uint32)), xyz.ctypes.data_as(_dp), max_points)
    return bool(ok), rgba, npts, xyz


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def oracle_score(params, scene, linvels, angvels, sfm=None, margins=False, threads=0):
    """Cost vector (float64), SfwBest and optional margins of one scene from the C oracle."""
    sa = scene if isinstance(scene, SceneArray) else SceneArray([scene])
    lin, ang = _d(linvels), _d(angvels)
    costs = np.empty(len(lin) * len(ang), dtype=np.float64)
    best = SfwBest()
    sp = C.byref(sfm) if sfm is not None else None
    if threads and threads > 1:
        rc = oracle().sfw_oracle_score_mt(C.byref(params), sp, sa.ptr(0), lin.ctypes.data_as(_dp),
                                          len(lin), ang.ctypes.data_as(_dp), len(ang),
                                          costs.ctypes.data_as(_dp), C.byref(best), int(threads))
        assert rc == 0
        return costs, best, None
    mg = np.empty(len(costs), dtype=MARGIN_DTYPE) if margins else None
    mp = mg.ctypes.data_as(C.POINTER(SfwOracleMargins)) if margins else None
    rc = oracle().sfw_oracle_score(C.byref(params), sp, sa.ptr(0), lin.ctypes.data_as(_dp), len(lin),
                                   ang.ctypes.data_as(_dp), len(ang), costs.ctypes.data_as(_dp),
                                   C.byref(best), mp)
    assert rc == 0
    return costs, best, mg


def _probe_cfg(margins, max_events, flips=None):
    cfg = SfwOracleProbe()
    cfg.goal_margin, cfg.collision_margin, cfg.theta_margin, cfg.theta_min_weight = margins
    cfg.max_events = max_events
    if flips is not None and len(flips):
        cfg.flips = flips.ctypes.data
        cfg.n_flips = len(flips)
    return cfg


def oracle_probe_grid(params, scene, linvels, angvels, margins, sfm=None, first=0, count=None, max_events=32,
                      threads=None):
    """Samples [first, first + count) of the grid with the branch probe on (oracle/sfw_oracle.h):
    (costs float64[count], events EVENT_DTYPE[count, max_events], n_events uint32[count])."""
    sa = scene if isinstance(scene, SceneArray) else SceneArray([scene])
    lin, ang = _d(linvels), _d(angvels)
    if count is None:
        count = len(lin) * len(ang) - first
    costs = np.empty(count, dtype=np.float64)
    events = np.zeros((count, max_events), dtype=EVENT_DTYPE)
    n_ev = np.zeros(count, dtype=np.uint32)
    cfg = _probe_cfg(margins, max_events)
    rc = oracle().sfw_oracle_score_probe(C.byref(params), C.byref(sfm) if sfm is not None else None, sa.ptr(0),
                                         lin.ctypes.data_as(_dp), len(lin), ang.ctypes.data_as(_dp), len(ang),
                                         first, count, C.byref(cfg), costs.ctypes.data_as(_dp), events.ctypes.data,
                                         n_ev.ctypes.data_as(C.POINTER(C.c_uint32)),
                                         int(threads or os.cpu_count() or 1))
    assert rc == 0
    return costs, events, n_ev


def oracle_probe_one(params, scene, v, w, margins, flips=None, sfm=None, max_events=32):
    """One trajectory with listed decisions forced (flips: EVENT_DTYPE array): (cost, events met)."""
    sa = scene if isinstance(scene, SceneArray) else SceneArray([scene])
    events = np.zeros(max_events, dtype=EVENT_DTYPE)
    fl = np.ascontiguousarray(flips, dtype=EVENT_DTYPE) if flips is not None else None
    cfg = _probe_cfg(margins, max_events, fl)
    cfg.events = events.ctypes.data
    c = oracle().sfw_oracle_score_trajectory_probe(C.byref(params), C.byref(sfm) if sfm is not None else None,
                                                   sa.ptr(0), float(v), 0.0, float(w), params.max_trans_acc, 0.0,
                                                   params.max_rot_acc, C.byref(cfg))
    return c, events[:min(cfg.n_events, max_events)].copy()


def ref_score(params, scene, linvels, angvels, sfm=None, want_best=True):
    """Same from the reference's own compiled sources (oracle/_ref)."""
    sa = scene if isinstance(scene, SceneArray) else SceneArray([scene])
    lin, ang = _d(linvels), _d(angvels)
    costs = np.empty(len(lin) * len(ang), dtype=np.float64)
    best = SfwBest()
    sp = C.byref(sfm) if sfm is not None else None
    rc = ref().sfw_ref_score(C.byref(params), sp, sa.ptr(0), lin.ctypes.data_as(_dp), len(lin),
                             ang.ctypes.data_as(_dp), len(ang), costs.ctypes.data_as(_dp),
                             C.byref(best) if want_best else None)
    assert rc == 0
    return costs, best


def oracle_argmin(costs, linvels, angvels) -> SfwBest:
    """Arg-min of a cost vector under the reference's sequential best-update (sfw_oracle_argmin)."""
    c64, lin, ang = _d(costs).reshape(-1), _d(linvels), _d(angvels)
    sb = SfwBest()
    oracle().sfw_oracle_argmin(c64.ctypes.data_as(_dp), lin.ctypes.data_as(_dp), len(lin), ang.ctypes.data_as(_dp),
                               len(ang), C.byref(sb))
    return sb


# ---- the drop-in proof: the reference's UNMODIFIED node + sensor interface on top of plugin/src/sfw_planner.cpp ----
DROPIN_NODE_SO = os.path.join(ORACLE_DIR, "_ref", "libsfw_dropin_node.so")
_dropin_node = None


def have_dropin_node() -> bool:
    return os.path.exists(DROPIN_NODE_SO)


def dropin_node_run(params, ext, scene, scan, people, odom, plan, plan_has_tf=False, tf=(0.0, 0.0, 0.0), ticks=1):
    """oracle/_ref/libsfw_dropin_node.so (oracle/Makefile target `dropin`): src/sfw_planner_node.cpp and
    src/sensor_interface.cpp of the reference, compiled unmodified against plugin/include's sfw_planner.hpp, with the
    B200 planner core (plugin/src/sfw_planner.cpp -> libsfw_b200.so) underneath.  Same call as ref_node_run."""
    global _dropin_node
    if _dropin_node is None:
        lib = C.CDLL(DROPIN_NODE_SO)
        lib.sfw_dropin_node_run.restype = C.c_int
        lib.sfw_dropin_node_run.argtypes = NODE_ARGTYPES
        _dropin_node = lib
    return node_call(_dropin_node.sfw_dropin_node_run, params, ext, scene, scan, people, odom, plan, plan_has_tf, tf,
                     ticks)
