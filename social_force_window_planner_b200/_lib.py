"""Loader of the C-ABI library ``libsfw_b200.so`` (see ``include/sfw_b200.h``).

The library is the product; there is no Python or CPU fallback.  If it has not been built
(``python -m social_force_window_planner_b200.build``) importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

from ._abi import SfwBest, SfwLaserScan, SfwLimits, SfwParams, SfwScene, SfwSfmParams

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFW_B200_LIB") or os.path.join(HERE, "libsfw_b200.so")  # override: kernel experiments

# every symbol include/sfw_b200.h declares
EXPORTS = [
    "sfw_abi_version", "sfw_create", "sfw_destroy", "sfw_last_error", "sfw_default_params",
    "sfw_default_sfm_params", "sfw_score", "sfw_score_batch", "sfw_upload", "sfw_run",
    "sfw_download", "sfw_sync", "sfw_set_row_slab", "sfw_trajectory_points", "sfw_stream",
    "sfw_device_costs", "sfw_device_best", "sfw_kernel_launches", "sfw_algorithmic_bytes",
    "sfw_last_kernel", "sfw_block_threads", "sfw_shared_prefix_steps", "sfw_h2d_bytes", "sfw_d2h_bytes", "sfw_laser_obstacles", "sfw_marker_points", "sfw_exchange_export", "sfw_exchange_connect",
    "sfw_exchange_sync", "sfw_exchange_fetch", "sfw_exchange_device_buffer", "sfw_exchange_expect", "sfw_exchange_connect_local", "sfw_set_zero_sample", "sfw_set_host_threads",
    "sfw_exchange_set_timeout", "sfw_exchange_merge", "sfw_exchange_merged_device", "sfw_set_policy", "sfw_set_prefix_sharing", "sfw_may_i_stop",
    "sfw_set_obstacle_cutoff", "sfw_obstacle_skip_fraction", "sfw_obstacle_layout",
]

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ctx = C.c_void_p


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m social_force_window_planner_b200.build` "
            "(this package has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.sfw_abi_version.restype = C.c_int
    lib.sfw_create.restype = C.c_int
    lib.sfw_create.argtypes = [C.POINTER(_ctx), C.c_int, C.c_void_p, C.POINTER(SfwLimits)]
    lib.sfw_destroy.restype = C.c_int
    lib.sfw_destroy.argtypes = [_ctx]
    lib.sfw_last_error.restype = C.c_char_p
    lib.sfw_last_error.argtypes = [_ctx]
    lib.sfw_default_params.restype = None
    lib.sfw_default_params.argtypes = [C.POINTER(SfwParams)]
    lib.sfw_default_sfm_params.restype = None
    lib.sfw_default_sfm_params.argtypes = [C.POINTER(SfwSfmParams)]
    score_args = [_ctx, C.POINTER(SfwParams), C.POINTER(SfwSfmParams), C.POINTER(SfwScene)]
    lib.sfw_score.restype = C.c_int
    lib.sfw_score.argtypes = score_args + [_dp, C.c_uint32, _dp, C.c_uint32, _fp, C.POINTER(SfwBest)]
    lib.sfw_score_batch.restype = C.c_int
    lib.sfw_score_batch.argtypes = score_args + [C.c_uint32, _dp, C.c_uint32, _dp, C.c_uint32, _fp,
                                                 C.POINTER(SfwBest)]
    lib.sfw_upload.restype = C.c_int
    lib.sfw_upload.argtypes = score_args + [C.c_uint32, _dp, C.c_uint32, _dp, C.c_uint32]
    lib.sfw_run.restype = C.c_int
    lib.sfw_run.argtypes = [_ctx]
    lib.sfw_download.restype = C.c_int
    lib.sfw_download.argtypes = [_ctx, _fp, C.POINTER(SfwBest)]
    lib.sfw_sync.restype = C.c_int
    lib.sfw_sync.argtypes = [_ctx]
    lib.sfw_set_row_slab.restype = C.c_int
    lib.sfw_set_row_slab.argtypes = [_ctx, C.c_uint32, C.c_uint32]
    lib.sfw_trajectory_points.restype = C.c_int
    lib.sfw_trajectory_points.argtypes = [_ctx, C.c_uint32, C.c_uint32, _dp, C.c_uint32,
                                          C.POINTER(C.c_uint32)]
    lib.sfw_stream.restype = C.c_void_p
    lib.sfw_stream.argtypes = [_ctx]
    lib.sfw_device_costs.restype = C.c_void_p
    lib.sfw_device_costs.argtypes = [_ctx]
    lib.sfw_device_best.restype = C.c_void_p
    lib.sfw_device_best.argtypes = [_ctx]
    lib.sfw_kernel_launches.restype = C.c_uint64
    lib.sfw_kernel_launches.argtypes = [_ctx]
    lib.sfw_algorithmic_bytes.restype = C.c_uint64
    lib.sfw_algorithmic_bytes.argtypes = [_ctx]
    lib.sfw_block_threads.restype = C.c_uint32
    lib.sfw_block_threads.argtypes = [_ctx]
    for f in (lib.sfw_h2d_bytes, lib.sfw_d2h_bytes):
        f.restype = C.c_uint64
        f.argtypes = [_ctx]
    lib.sfw_laser_obstacles.restype = C.c_int
    lib.sfw_laser_obstacles.argtypes = [_ctx, C.POINTER(SfwLaserScan), C.c_uint32, C.c_float, C.c_float, _dp,
                                        C.c_uint32, C.POINTER(C.c_uint32)]
    lib.sfw_marker_points.restype = C.c_int
    lib.sfw_marker_points.argtypes = [_ctx, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _dp, C.c_uint32,
                                      C.POINTER(C.c_uint16)]
    lib.sfw_may_i_stop.restype = C.c_int
    lib.sfw_may_i_stop.argtypes = [_ctx, C.c_uint32] + [C.c_double] * 7 + [C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
    lib.sfw_set_prefix_sharing.restype = C.c_int
    lib.sfw_set_prefix_sharing.argtypes = [_ctx, C.c_int]
    lib.sfw_set_host_threads.restype = C.c_int
    lib.sfw_set_host_threads.argtypes = [_ctx, C.c_int]
    lib.sfw_set_zero_sample.restype = C.c_int
    lib.sfw_set_zero_sample.argtypes = [_ctx, C.c_int]
    lib.sfw_set_policy.restype = C.c_int
    lib.sfw_set_policy.argtypes = [_ctx, C.c_int]
    lib.sfw_exchange_export.restype = C.c_int
    lib.sfw_exchange_export.argtypes = [_ctx, C.c_uint32, C.c_void_p]
    lib.sfw_exchange_connect.restype = C.c_int
    lib.sfw_exchange_connect.argtypes = [_ctx, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.sfw_exchange_sync.restype = C.c_int
    lib.sfw_exchange_sync.argtypes = [_ctx]
    lib.sfw_exchange_fetch.restype = C.c_int
    lib.sfw_exchange_fetch.argtypes = [_ctx, C.POINTER(SfwBest)]
    lib.sfw_exchange_device_buffer.restype = C.c_void_p
    lib.sfw_exchange_device_buffer.argtypes = [_ctx]
    lib.sfw_exchange_merged_device.restype = C.c_void_p
    lib.sfw_exchange_merged_device.argtypes = [_ctx]
    lib.sfw_exchange_connect_local.restype = C.c_int
    lib.sfw_exchange_connect_local.argtypes = [C.POINTER(_ctx), C.c_uint32]
    lib.sfw_exchange_expect.restype = C.c_int
    lib.sfw_exchange_expect.argtypes = [_ctx, C.POINTER(C.c_uint32)]
    lib.sfw_exchange_set_timeout.restype = C.c_int
    lib.sfw_exchange_set_timeout.argtypes = [_ctx, C.c_double]
    lib.sfw_exchange_merge.restype = C.c_int
    lib.sfw_exchange_merge.argtypes = [_ctx, C.POINTER(SfwBest)]
    lib.sfw_last_kernel.restype = C.c_char_p
    lib.sfw_last_kernel.argtypes = [_ctx]
    lib.sfw_shared_prefix_steps.restype = C.c_double
    lib.sfw_shared_prefix_steps.argtypes = [_ctx]
    lib.sfw_set_obstacle_cutoff.restype = C.c_int
    lib.sfw_set_obstacle_cutoff.argtypes = [_ctx, C.c_double]
    lib.sfw_obstacle_layout.restype = C.c_uint32
    lib.sfw_obstacle_layout.argtypes = [_dp, C.c_uint32] + [C.c_double] * 5 + [_fp, C.c_uint32]
    lib.sfw_obstacle_skip_fraction.restype = C.c_double
    lib.sfw_obstacle_skip_fraction.argtypes = [_ctx]
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load()
    return _lib
