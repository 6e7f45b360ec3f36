"""CPU: the obstacle-cluster layout the host packer builds (sfw_obstacle_layout = the routine sfw_upload uses;
DESIGN.md 3 and 4.1 item 8).  lightsfm sums the obstacle force over every point (reached through
the computeForces call at reference src/sfw_planner.cpp:592), so the layout must hold every point exactly once, and a
cluster may only be skipped by a pedestrian for whom EVERY term of it is below 2^-cutoff of the force factor."""
import ctypes as C
import math

import numpy as np
import pytest

from social_force_window_planner_b200 import _lib

SIGMA, R_MAX, CUT = 0.2, 0.35, 24.0
SCALE = float(np.float32(1.4426950408889634 / SIGMA))
FAR = 1.0e15


def _layout(pts, ref=(0.0, 0.0), cutoff=CUT, r_max=R_MAX):
    lib = _lib.load()
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
    dp = pts.ctypes.data_as(C.POINTER(C.c_double))
    n = lib.sfw_obstacle_layout(dp, len(pts), ref[0], ref[1], SIGMA, r_max, cutoff, None, 0)
    assert n == 10 * ((len(pts) + 7) // 8)
    out = np.zeros((max(n, 1), 2), dtype=np.float32)
    assert lib.sfw_obstacle_layout(dp, len(pts), ref[0], ref[1], SIGMA, r_max, cutoff,
                                   out.ctypes.data_as(C.POINTER(C.c_float)), n) == n
    return out[:n].reshape(-1, 10, 2)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 32, 100, 523])
def test_every_point_once_and_bounds_cover(n):
    rng = np.random.default_rng(n)
    pts = rng.uniform(-6.0, 6.0, (n, 2))
    if n >= 9:
        pts[3] = pts[5]  # duplicate points are legal (two laser hits on one spot)
    ref = (1.25, -0.5)
    cl = _layout(pts, ref)
    assert cl.shape[0] == (n + 7) // 8
    want = np.stack([np.float32((pts[:, 0] - ref[0]) * SCALE), np.float32((pts[:, 1] - ref[1]) * SCALE)], 1)
    got = cl[:, 2:, :].reshape(-1, 2)
    real = got[:, 0] < FAR / 2
    assert real.sum() == n and np.all(got[~real] == np.float32([FAR, 0.0]))
    # padding only at the tail of the last cluster
    assert np.all(real[:n]) and not real[n:].any()
    key = lambda a: sorted(map(tuple, a.tolist()))  # noqa: E731
    assert key(got[real]) == key(want)
    for g in range(cl.shape[0]):
        c, reach2 = cl[g, 0], float(cl[g, 1, 0])
        p = cl[g, 2:][cl[g, 2:, 0] < FAR / 2].astype(np.float64)
        rad = np.hypot(p[:, 0] - c[0], p[:, 1] - c[1]).max()
        assert math.sqrt(reach2) >= rad + R_MAX * SCALE + CUT  # never tighter than the exact bound
        assert math.sqrt(reach2) <= (rad + R_MAX * SCALE + CUT) * (1 + 1e-5)


def test_skipped_terms_are_below_the_cutoff():
    """For pedestrians the device test would skip, every term exp2(-(|p - o| - r)) of the cluster is < 2^-24."""
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.normal((2.0, 1.0), 0.3, (40, 2)), rng.normal((-3.0, -2.0), 0.5, (30, 2))])
    cl = _layout(pts)
    peds = rng.uniform(-8.0, 8.0, (400, 2)) * SCALE
    n_skip = 0
    for g in range(cl.shape[0]):
        c, reach2 = cl[g, 0].astype(np.float64), float(cl[g, 1, 0])
        p = cl[g, 2:][cl[g, 2:, 0] < FAR / 2].astype(np.float64)
        d2c = (peds[:, 0] - c[0]) ** 2 + (peds[:, 1] - c[1]) ** 2
        for q in peds[d2c > reach2]:
            d = np.hypot(p[:, 0] - q[0], p[:, 1] - q[1])
            assert np.all(d - R_MAX * SCALE > CUT)
            n_skip += 1
    assert n_skip > 100  # the scene does exercise the skip


def test_clusters_are_compact_and_deterministic():
    # two walls far apart: no cluster may straddle them, whatever the input order
    a = np.stack([np.linspace(-1, 1, 16), np.full(16, 3.0)], 1)
    b = np.stack([np.full(16, -4.0), np.linspace(-1, 1, 16)], 1)
    pts = np.concatenate([a, b])
    rng = np.random.default_rng(0)
    first = _layout(pts[rng.permutation(32)])
    for _ in range(3):
        again = _layout(pts[rng.permutation(32)])
        assert np.array_equal(first, again)  # the layout depends on the point set only
    for g in range(4):
        p = first[g, 2:].astype(np.float64) / SCALE
        assert np.ptp(p[:, 0]) < 2.1 and np.ptp(p[:, 1]) < 2.1
        assert (np.abs(p[:, 1] - 3.0) < 1e-5).all() or (np.abs(p[:, 0] + 4.0) < 1e-5).all()


def test_cutoff_off_and_bad_input():
    pts = np.random.default_rng(1).uniform(-3, 3, (20, 2))
    cl = _layout(pts, cutoff=0.0)
    assert np.all(np.isinf(cl[:, 1, 0]))  # reach = inf: nothing is ever skipped
    nan = pts.copy()
    nan[4, 0] = np.nan
    cl = _layout(nan)  # must not crash or loop; the NaN point is still in the layout
    assert np.isnan(cl[:, 2:, 0]).sum() == 1
    lib = _lib.load()
    assert lib.sfw_obstacle_layout(None, 5, 0.0, 0.0, SIGMA, R_MAX, CUT, None, 0) == 0
