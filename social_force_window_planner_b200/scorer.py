"""`Scorer`: thin object wrapper over one ``sfw_ctx`` of the C ABI.

Every method is a direct call into ``libsfw_b200.so``; results come from the CUDA kernels only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._abi import BEST_DTYPE, SceneArray, SfwBest, SfwLaserScan, SfwParams, SfwSfmParams

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


class SfwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sfw error {code}: {msg}")
        self.code = code


class Scorer:
    """One scoring context (device buffers + stream) on one GPU."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _lib.lib()
        self._ctx = C.c_void_p()
        rc = self._lib.sfw_create(C.byref(self._ctx), device, C.c_void_p(stream) if stream else None,
                                  None)
        if rc != 0:
            msg = self._lib.sfw_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise SfwError(rc, msg)
        self._keep = None
        self.n_scenes = 0
        self.n_samples = 0

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.sfw_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise SfwError(rc, self._lib.sfw_last_error(self._ctx).decode())

    # -- split form -------------------------------------------------------------------------
    def upload(self, params: SfwParams, scenes, linvels, angvels, sfm: SfwSfmParams | None = None):
        sa = scenes if isinstance(scenes, SceneArray) else SceneArray(scenes)
        lin = np.ascontiguousarray(linvels, dtype=np.float64)
        ang = np.ascontiguousarray(angvels, dtype=np.float64)
        self._keep = (sa, lin, ang)
        self._check(self._lib.sfw_upload(self._ctx, C.byref(params), C.byref(sfm) if sfm else None,
                                         sa.ptr(0), len(sa), lin.ctypes.data_as(_dp), len(lin),
                                         ang.ctypes.data_as(_dp), len(ang)))
        self.n_scenes = len(sa)
        self.n_samples = len(lin) * len(ang)

    POLICY_AUTO, POLICY_THROUGHPUT, POLICY_LATENCY = 0, 1, 2

    def set_policy(self, policy: int):
        """Kernel selection policy (``sfw_set_policy``): AUTO switches small grids to the low-latency kernel."""
        self._check(self._lib.sfw_set_policy(self._ctx, policy))

    def set_host_threads(self, n_threads: int):
        """``sfw_set_host_threads``: host workers that pack a batch of scenes (1 = the calling thread only)."""
        self._check(self._lib.sfw_set_host_threads(self._ctx, int(n_threads)))

    def set_zero_sample(self, score_it: bool):
        """``sfw_set_zero_sample``: score the (0,0) sample (single scoreTrajectory calls of the reference) instead
        of skipping it (its grid loop)."""
        self._check(self._lib.sfw_set_zero_sample(self._ctx, 1 if score_it else 0))

    def set_prefix_sharing(self, on):
        """Rollout prefix sharing (``sfw_set_prefix_sharing``); bit-identical results.  False / 0 = never,
        True / 1 = when the library's cost model says it pays, 2 = whenever the batch allows it."""
        self._check(self._lib.sfw_set_prefix_sharing(self._ctx, int(on)))

    def set_obstacle_cutoff(self, cutoff_log2: float):
        """Far-field cutoff of the pedestrians' obstacle force (``sfw_set_obstacle_cutoff``): clusters of obstacle
        points whose every term is below ``2**-cutoff_log2`` of the force factor are skipped.  Default 24; <= 0
        sums every term.  Applies from the next upload."""
        self._check(self._lib.sfw_set_obstacle_cutoff(self._ctx, float(cutoff_log2)))

    def set_row_slab(self, row_begin: int, row_end: int):
        self._check(self._lib.sfw_set_row_slab(self._ctx, row_begin, row_end))

    def run(self):
        self._check(self._lib.sfw_run(self._ctx))

    def sync(self):
        self._check(self._lib.sfw_sync(self._ctx))

    def download(self, want_costs: bool = True):
        best = np.zeros(self.n_scenes, dtype=BEST_DTYPE)
        costs = np.empty((self.n_scenes, self.n_samples), dtype=np.float32) if want_costs else None
        self._check(self._lib.sfw_download(
            self._ctx, costs.ctypes.data_as(_fp) if want_costs else None,
            best.ctypes.data_as(C.POINTER(SfwBest))))
        return costs, best

    # -- one-call form ------------------------------------------------------------------------
    def score(self, params, scenes, linvels, angvels, sfm=None, want_costs=True, out=None):
        """``sfw_score_batch``: returns (costs[n_scenes, n_v*n_w] float32 or None, best[n_scenes]).
        ``out`` = (costs, best) of a previous call with the same shapes: the caller's output buffers are reused,
        as a C/C++ caller of the ABI would (a fresh multi-megabyte array per tick is page-faulted in every time)."""
        sa = scenes if isinstance(scenes, SceneArray) else SceneArray(scenes)
        lin = np.ascontiguousarray(linvels, dtype=np.float64)
        ang = np.ascontiguousarray(angvels, dtype=np.float64)
        self._keep = (sa, lin, ang)
        self.n_scenes = len(sa)
        self.n_samples = len(lin) * len(ang)
        if out is not None and out[1].shape == (self.n_scenes,) and (
                not want_costs or (out[0] is not None and out[0].shape == (self.n_scenes, self.n_samples))):
            costs, best = (out[0] if want_costs else None), out[1]
        else:
            best = np.zeros(self.n_scenes, dtype=BEST_DTYPE)
            costs = np.empty((self.n_scenes, self.n_samples), dtype=np.float32) if want_costs else None
        self._check(self._lib.sfw_score_batch(
            self._ctx, C.byref(params), C.byref(sfm) if sfm else None, sa.ptr(0), len(sa),
            lin.ctypes.data_as(_dp), len(lin), ang.ctypes.data_as(_dp), len(ang),
            costs.ctypes.data_as(_fp) if want_costs else None, best.ctypes.data_as(C.POINTER(SfwBest))))
        return costs, best

    def trajectory_points(self, scene: int, sample_index: int, max_points: int = 65535):
        n = C.c_uint32(0)
        buf = np.zeros((max_points, 3), dtype=np.float64)
        self._check(self._lib.sfw_trajectory_points(self._ctx, scene, sample_index,
                                                    buf.ctypes.data_as(_dp), max_points, C.byref(n)))
        return buf[: min(n.value, max_points)].copy(), n.value

    def marker_points(self, scene: int, first: int = 0, stride: int = 1, count: int | None = None,
                      max_points: int | None = None):
        """``sfw_marker_points``: recorded rollout points of samples first, first+stride, ... in one launch.
        Returns (xyz float64[count, max_points, 3], n_points uint16[count])."""
        if count is None:
            count = (self.n_samples - first + stride - 1) // stride
        if max_points is None:
            max_points = 256
        xyz = np.zeros((count, max_points, 3), dtype=np.float64)
        n = np.zeros(count, dtype=np.uint16)
        self._check(self._lib.sfw_marker_points(self._ctx, scene, first, stride, count, xyz.ctypes.data_as(_dp),
                                                max_points, n.ctypes.data_as(C.POINTER(C.c_uint16))))
        return xyz, n

    # -- multi-GPU: winner exchange fused into the scorer's epilogue -----------------------------------
    def exchange_export(self, max_scenes: int) -> bytes:
        """Allocate this rank's gather buffer; returns its 64-byte cudaIpc handle (ship it to every rank)."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.sfw_exchange_export(self._ctx, max_scenes, buf))
        return buf.raw

    def exchange_connect(self, rank: int, world: int, handles):
        """``handles``: the ``world`` exported handles in rank order."""
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._xchg_world = world
        self._xchg_counts = None
        self._check(self._lib.sfw_exchange_connect(self._ctx, rank, world, C.c_char_p(blob)))

    @staticmethod
    def exchange_connect_local(scorers):
        """Connect the given scorers of THIS process as ranks 0 .. len - 1 (``sfw_exchange_connect_local``); each
        must have called ``exchange_export`` with the same ``max_scenes``."""
        arr = (C.c_void_p * len(scorers))(*[s._ctx.value for s in scorers])
        rc = scorers[0]._lib.sfw_exchange_connect_local(arr, len(scorers))
        for s in scorers:
            s._xchg_world = len(scorers)
            s._xchg_counts = None
        if rc != 0:
            msgs = [s._lib.sfw_last_error(s._ctx).decode() for s in scorers]
            raise SfwError(rc, "; ".join(m for m in msgs if m))

    def exchange_sync(self):
        """Device-side wait (enqueued on the context stream) for every rank's records of the latest run."""
        self._check(self._lib.sfw_exchange_sync(self._ctx))

    def exchange_expect(self, scenes_per_rank):
        """How many scenes each rank stages per tick from now on (``sfw_exchange_expect``); None = every rank
        stages what this rank stages."""
        if scenes_per_rank is None:
            self._xchg_counts = None
            self._check(self._lib.sfw_exchange_expect(self._ctx, None))
            return
        cnt = np.ascontiguousarray(scenes_per_rank, dtype=np.uint32)
        assert len(cnt) == self._xchg_world
        self._xchg_counts = cnt.copy()
        self._check(self._lib.sfw_exchange_expect(self._ctx, cnt.ctypes.data_as(C.POINTER(C.c_uint32))))

    def exchange_set_timeout(self, seconds: float):
        """Bound of the device-side arrival wait; a peer that does not deliver makes fetch / merge raise."""
        self._check(self._lib.sfw_exchange_set_timeout(self._ctx, float(seconds)))

    def exchange_fetch(self):
        """Gathered winners of the latest run.  Equal scene counts on every rank: BEST_DTYPE[world, n_scenes]
        (rank major); after ``exchange_expect`` with differing counts: BEST_DTYPE[sum(counts)], rank after rank."""
        counts = getattr(self, "_xchg_counts", None)
        total = int(counts.sum()) if counts is not None else self._xchg_world * self.n_scenes
        out = np.zeros(total, dtype=BEST_DTYPE)
        self._check(self._lib.sfw_exchange_fetch(self._ctx, out.ctypes.data_as(C.POINTER(SfwBest))))
        if counts is None or (counts == counts[0]).all():
            return out.reshape(self._xchg_world, -1)
        return out

    def exchange_merge(self, sync: bool = True):
        """Row-slab mode: wait for every rank's slab winners and merge them per scene ON THE DEVICE with the
        reference's tie-break order (``sfw_exchange_merge``).  Returns BEST_DTYPE[n_scenes], or None when
        ``sync`` is False (the merged records stay on the device)."""
        if not sync:
            self._check(self._lib.sfw_exchange_merge(self._ctx, None))
            return None
        out = np.zeros(self.n_scenes, dtype=BEST_DTYPE)
        self._check(self._lib.sfw_exchange_merge(self._ctx, out.ctypes.data_as(C.POINTER(SfwBest))))
        return out

    def may_i_stop(self, scene: int, vl_x, vl_y, va, x, y, th, dt):
        """``sfw_may_i_stop`` (SFWPlanner::mayIStop, reference src/sfw_planner.cpp:718-765): (can_stop, steps)."""
        ok, steps = C.c_int32(0), C.c_uint32(0)
        self._check(self._lib.sfw_may_i_stop(self._ctx, scene, vl_x, vl_y, va, x, y, th, dt, C.byref(ok), C.byref(steps)))
        return bool(ok.value), int(steps.value)

    # -- the step before the path: laser scans -> obstacle points ----------------------------------
    def laser_obstacles(self, scans, max_obstacle_dist: float = 3.0, person_radius: float = 0.35):
        """``sfw_laser_obstacles`` (SFMSensorInterface::laserCb, reference src/sensor_interface.cpp:103-229).

        ``scans``: list of dicts ``ranges`` (float32[n]), ``angle_min``, ``angle_increment``, optional
        ``tf`` = (x, y, yaw) laser->controller frame, optional ``people`` (float64[k, 2], controller
        frame).  Returns one float64[m, 2] array of obstacle points per scan, beam order kept."""
        arr = (SfwLaserScan * len(scans))()
        keep = []
        cap = 1
        for a, sc in zip(arr, scans):
            r = np.ascontiguousarray(sc["ranges"], dtype=np.float32)
            ppl = np.ascontiguousarray(sc.get("people", np.zeros((0, 2))), dtype=np.float64).reshape(-1, 2)
            keep += [r, ppl]
            a.ranges = r.ctypes.data_as(_fp)
            a.n_ranges = len(r)
            a.angle_min = sc["angle_min"]
            a.angle_increment = sc["angle_increment"]
            tf = sc.get("tf")
            a.has_tf = 1 if tf is not None else 0
            if tf is not None:
                a.tf_x, a.tf_y, a.tf_yaw = tf
            a.people_xy = ppl.ctypes.data_as(_dp)
            a.n_people = len(ppl)
            cap = max(cap, len(r))
        out = np.zeros((len(scans), cap, 2), dtype=np.float64)
        cnt = np.zeros(len(scans), dtype=np.uint32)
        self._check(self._lib.sfw_laser_obstacles(self._ctx, arr, len(scans), max_obstacle_dist, person_radius,
                                                  out.ctypes.data_as(_dp), cap,
                                                  cnt.ctypes.data_as(C.POINTER(C.c_uint32))))
        return [out[k, :cnt[k]].copy() for k in range(len(scans))]

    # -- introspection --------------------------------------------------------------------------
    @property
    def stream(self) -> int:
        return self._lib.sfw_stream(self._ctx) or 0

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.sfw_kernel_launches(self._ctx))

    @property
    def algorithmic_bytes(self) -> int:
        return int(self._lib.sfw_algorithmic_bytes(self._ctx))

    @property
    def h2d_bytes(self) -> int:
        return int(self._lib.sfw_h2d_bytes(self._ctx))

    @property
    def d2h_bytes(self) -> int:
        return int(self._lib.sfw_d2h_bytes(self._ctx))

    @property
    def last_kernel(self) -> str:
        return self._lib.sfw_last_kernel(self._ctx).decode()

    @property
    def block_threads(self) -> int:
        """Threads per block of the staged batch's launch plan (``sfw_block_threads``)."""
        return int(self._lib.sfw_block_threads(self._ctx))

    @property
    def obstacle_skip_fraction(self) -> float:
        """Share of (pedestrian, obstacle cluster) combinations out of reach at the start poses of the staged batch."""
        return float(self._lib.sfw_obstacle_skip_fraction(self._ctx))

    @property
    def shared_prefix_steps(self) -> float:
        """Mean number of leading steps a sample of the staged batch takes from a shared path (0: sharing off)."""
        return float(self._lib.sfw_shared_prefix_steps(self._ctx))
