"""ctypes access to the TEST-ONLY checkers: oracle/libsfw_oracle.so (C restatement) and
oracle/_ref/libsfw_ref.so (the reference's own sources compiled unmodified).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from social_force_window_planner_b200._abi import (SceneArray, SfwBest, SfwParams, SfwScene,
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
        _oracle = lib
    return _oracle


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
        lib.sfw_ref_default_samples.restype = C.c_int
        lib.sfw_ref_default_samples.argtypes = [C.c_double, C.c_double, _dp, _dp]
        _ref = lib
    return _ref


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
