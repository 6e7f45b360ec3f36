"""Shared parity checker: CUDA cost vector / winner vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): per-trajectory cost within 1e-4 relative (RTOL), validity identical,
arg-min index identical whenever the runner-up differs by more than 1e-3.  ONE tolerance, everywhere.

The model has genuine discontinuities: a pedestrian's goal pops inside its goal radius, the rollout dies when
the robot touches a pedestrian (reference src/sfw_planner.cpp:613-627), lightsfm's angular term carries
sign(theta), group repulsion switches on at contact.  An FP32 evaluator cannot take such a decision the same
way as an FP64 one when it is taken within rounding distance of its switching surface — but then it must agree
with the oracle ON THE BRANCH IT TOOK.  The oracle's branch probe (oracle/sfw_oracle.h, SfwOracleProbe) records
every decision a trajectory takes within the margins below and can re-run the trajectory with listed decisions
forced the other way.  The check per trajectory:

  1. GPU cost within RTOL of the oracle's cost, same validity            -> "base" (the usual case), else
  2. some subset (<= MAX_FLIPS) of the recorded near decisions, forced the other way, gives an oracle cost
     within RTOL of the GPU's with the same validity                     -> "branch", else
  3. the test fails.

There is no looser tolerance for near-discontinuity trajectories.  Every call appends its statistics to
``STATS`` (tests/conftest.py writes them to gpurun_out/parity_stats.json at session end).
"""
from __future__ import annotations

import itertools

import numpy as np

import oracle_lib as ol

RTOL = 1e-4
GOAL_MARGIN = 5e-6       # metres
COLLISION_MARGIN = 5e-6  # metres
THETA_MARGIN = 5e-6      # radians
THETA_MIN_WEIGHT = 1e-6  # angular terms smaller than this cannot move a cost by 1e-4 relative
MARGINS = (GOAL_MARGIN, COLLISION_MARGIN, THETA_MARGIN, THETA_MIN_WEIGHT)
MAX_EVENTS = 32          # recorded decisions per trajectory
MAX_FLIPS = 3            # decisions forced the other way at once
MAX_CANDIDATES = 8       # ... chosen among the most influential recorded decisions
MAX_BRANCH_RUNS = 4000   # oracle re-runs one compare() may spend on branch resolution

STATS = []               # one dict per compare() / check_samples() call, in call order
_LABEL = [None]          # set by the conftest fixture to the running test's id


def _rel(g, o):
    return abs(g - o) / max(abs(o), 1e-12)


def _match(g, o):
    """Same validity and, when both valid, within RTOL."""
    if (g >= 0) != (o >= 0):
        return False
    return o < 0 or _rel(g, o) <= RTOL


def _influence(ev):
    """Rough size of the jump a decision causes, to order the candidates."""
    k = ev["kind"]
    if k == ol.EV_COLLISION:
        return 1e9
    if k == ol.EV_GOAL:
        return 4.0           # desired force k v_des / tau
    if k == ol.EV_GROUP:
        return float(ev["weight"])
    return 2.0 * float(ev["weight"])


def _forced(ev):
    f = ev.copy()
    if f["kind"] == ol.EV_THETA:
        f["decision"] = -1 if f["decision"] > 0 else 1
    else:
        f["decision"] = 0 if f["decision"] else 1
    return f


def resolve_branch(params, scene, v, w, g, events, sfm=None, budget=None):
    """Find a subset of the recorded near decisions which, forced the other way, makes the oracle agree with the
    GPU value g.  Returns (oracle cost on that branch, number of forced decisions, runs) or (None, 0, runs)."""
    cand = sorted(range(len(events)), key=lambda i: -_influence(events[i]))[:MAX_CANDIDATES]
    runs = 0
    for k in range(1, MAX_FLIPS + 1):
        for sub in itertools.combinations(cand, k):
            if budget is not None and budget[0] <= 0:
                return None, 0, runs
            flips = np.array([_forced(events[i]) for i in sub], dtype=ol.EVENT_DTYPE)
            c, _ = ol.oracle_probe_one(params, scene, v, w, MARGINS, flips, sfm=sfm, max_events=1)
            runs += 1
            if budget is not None:
                budget[0] -= 1
            if _match(g, c):
                return c, k, runs
    return None, 0, runs


def check_samples(params, scene, lin, ang, idx, gpu_costs, oracle_costs=None, events=None, n_events=None, sfm=None,
                  label=None):
    """Samples ``idx`` (flat indices into the lin x ang grid) of a GPU cost vector against the oracle with the
    two-branch rule.  ``oracle_costs/events/n_events`` may come from a committed fixture (tests/golden); otherwise
    they are computed here.  Returns (stats, resolved oracle costs[len(idx)])."""
    lin = np.asarray(lin, dtype=np.float64)
    ang = np.asarray(ang, dtype=np.float64)
    idx = np.asarray(idx, dtype=np.int64)
    n_w = len(ang)
    gc = np.asarray(gpu_costs, dtype=np.float64).reshape(-1)
    assert len(gc) == len(idx)
    if oracle_costs is None:
        if len(idx) and np.array_equal(idx, np.arange(idx[0], idx[0] + len(idx))):
            oracle_costs, events, n_events = ol.oracle_probe_grid(params, scene, lin, ang, MARGINS, sfm=sfm,
                                                                  first=int(idx[0]), count=len(idx),
                                                                  max_events=MAX_EVENTS)
        else:
            oracle_costs = np.empty(len(idx))
            events = np.zeros((len(idx), MAX_EVENTS), dtype=ol.EVENT_DTYPE)
            n_events = np.zeros(len(idx), dtype=np.uint32)
            for k, i in enumerate(idx):
                v, w = lin[i // n_w], ang[i % n_w]
                if v == 0.0 and w == 0.0:
                    oracle_costs[k] = -2.0
                    continue
                c, ev = ol.oracle_probe_one(params, scene, v, w, MARGINS, sfm=sfm, max_events=MAX_EVENTS)
                oracle_costs[k] = c
                events[k, :len(ev)] = ev
                n_events[k] = len(ev)
    oc = np.asarray(oracle_costs, dtype=np.float64)
    skipped = oc == -2.0
    assert np.array_equal(skipped, gc == -2.0), "skipped (0,0) sample mismatch"
    resolved = oc.copy()
    near = (np.asarray(n_events) > 0) & ~skipped
    base_ok = np.array([_match(g, o) for g, o in zip(gc, oc)]) | skipped
    budget = [MAX_BRANCH_RUNS]
    n_branch = flips_max = runs_total = 0
    failures = []
    for k in np.nonzero(~base_ok)[0]:
        ev = events[k][:min(int(n_events[k]), events.shape[1])]
        i = int(idx[k])
        c = None
        if len(ev):
            c, nf, runs = resolve_branch(params, scene, lin[i // n_w], ang[i % n_w], gc[k], ev, sfm=sfm, budget=budget)
            runs_total += runs
        if c is None:
            failures.append((i, float(gc[k]), float(oc[k]), int(n_events[k])))
        else:
            resolved[k] = c
            n_branch += 1
            flips_max = max(flips_max, nf)
    both = (gc >= 0) & (resolved >= 0)
    rel = np.zeros(len(gc))
    rel[both] = np.abs(gc[both] - resolved[both]) / np.maximum(np.abs(resolved[both]), 1e-12)
    ok = np.ones(len(gc), dtype=bool)
    for f in failures:
        ok[np.nonzero(idx == f[0])[0]] = False
    stats = dict(label=label or _LABEL[0], n=int(len(gc)), valid=int((oc >= 0).sum()), near=int(near.sum()),
                 base_ok=int((base_ok & ~skipped).sum()), branch_resolved=int(n_branch), max_flips=int(flips_max),
                 branch_runs=int(runs_total), unresolved=len(failures),
                 max_rel_clear=float(rel[both & ~near & ok].max()) if (both & ~near & ok).any() else 0.0,
                 max_rel_near=float(rel[both & near & ok].max()) if (both & near & ok).any() else 0.0,
                 validity_flips_near=int(((gc >= 0) != (oc >= 0))[near].sum()),
                 max_events_per_traj=int(np.max(n_events)) if len(gc) else 0)
    STATS.append(stats)
    assert not failures, (f"{len(failures)} trajectories match neither the oracle nor any branch of its near "
                          f"decisions within {RTOL}: (index, gpu, oracle, near decisions) {failures[:8]} ({stats})")
    return stats, resolved


def compare(params, scene, lin, ang, gpu_costs, gpu_best, sfm=None, max_near_frac=None, label=None):
    """Whole cost vector + winner of one scene against the oracle.  ``max_near_frac`` is accepted for old call
    sites and ignored: near trajectories are held to the same RTOL through the branch rule."""
    lin = np.ascontiguousarray(lin, dtype=np.float64)
    ang = np.ascontiguousarray(ang, dtype=np.float64)
    n = len(lin) * len(ang)
    gc = np.asarray(gpu_costs, dtype=np.float64).reshape(-1)
    assert gc.shape == (n,)
    stats, resolved = check_samples(params, scene, lin, ang, np.arange(n), gc, sfm=sfm, label=label)
    # winner: self-consistent with the GPU cost vector under the reference's sequential rule ...
    from social_force_window_planner_b200._abi import SfwBest
    import ctypes as C
    dp = C.POINTER(C.c_double)

    def argmin(costs):
        sb = SfwBest()
        c64 = np.ascontiguousarray(costs, dtype=np.float64)
        ol.oracle().sfw_oracle_argmin(c64.ctypes.data_as(dp), lin.ctypes.data_as(dp), len(lin),
                                      ang.ctypes.data_as(dp), len(ang), C.byref(sb))
        return sb

    sb = argmin(gc)
    assert int(gpu_best["valid"]) == sb.valid
    if sb.valid:
        assert int(gpu_best["index"]) == sb.index, "GPU arg-min disagrees with its own cost vector"
        assert float(gpu_best["v"]) == sb.v and float(gpu_best["w"]) == sb.w
    # ... and identical to the oracle's (on the branches the GPU took) whenever its runner-up is > 1e-3 away
    ob = argmin(resolved)
    assert int(gpu_best["valid"]) == ob.valid
    if ob.valid:
        o_sorted = np.sort(resolved[resolved >= 0])
        gap = (o_sorted[1] - o_sorted[0]) if len(o_sorted) > 1 else np.inf
        if gap > 1e-3:
            assert int(gpu_best["index"]) == ob.index, (
                f"arg-min mismatch: gpu {int(gpu_best['index'])} oracle {ob.index} gap {gap}")
        stats["argmin_gap"] = float(gap)
    stats["best_oracle"] = ob.index if ob.valid else -1
    stats["best_gpu"] = int(gpu_best["index"]) if gpu_best["valid"] else -1
    return stats
