"""The oracle's branch probe (oracle/sfw_oracle.h: SfwOracleProbe) and the two-branch parity rule built on it
(tests/parity.py), exercised on the CPU: the checker itself has to be right before it may judge the GPU."""
import dataclasses

import numpy as np
import pytest

import golden_cases as G
import oracle_lib as ol
import parity
from social_force_window_planner_b200 import scenes as S

WIDE = (0.05, 0.05, 0.05, 1e-6)  # margins wide enough that ordinary scenes hold near decisions


def _case(n_peds=5, steps=24, seed=0, n_v=7, n_w=9):
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=n_v, n_w=n_w, steps=steps, n_peds=n_peds)
    return wl.params(), S.make_scene(wl, seed), *wl.sample_arrays()


@pytest.mark.parametrize("name", ["c0_seed0", "c0_hazards_seed5", "groups_tight_contact"])
def test_probe_without_flips_is_the_plain_oracle(name):
    p, sc, lin, ang = G.CASES[name]()
    base, _, _ = ol.oracle_score(p, sc, lin, ang)
    for margins in (parity.MARGINS, WIDE):
        costs, ev, n_ev = ol.oracle_probe_grid(p, sc, lin, ang, margins, threads=3)
        assert np.array_equal(costs, base), "recording decisions must not change a single bit"
    assert n_ev.max() > 0, "wide margins should meet some decisions"


def test_goal_pop_event_and_its_other_branch():
    p, sc, lin, ang = _case()
    # pedestrian 0 starts 1e-7 m outside its goal radius: lightsfm heads for the goal, one ulp closer it brakes
    q = sc.peds[0]
    q["x"], q["y"], q["vx"], q["vy"] = 1.2, 0.25, 0.6, 0.0
    q["goal_x"], q["goal_y"], q["goal_radius"] = 1.2 + 0.35 + 1e-7, 0.25, 0.35
    v, w = lin[4], ang[4]
    c0, ev = ol.oracle_probe_one(p, sc, v, w, parity.MARGINS)
    goal0 = [e for e in ev if e["kind"] == ol.EV_GOAL and e["step"] == 0 and e["a"] == 1]
    assert len(goal0) == 1 and goal0[0]["decision"] == 0 and goal0[0]["margin"] < 2e-7
    c1, _ = ol.oracle_probe_one(p, sc, v, w, parity.MARGINS, flips=np.array([parity._forced(goal0[0])]))
    assert c0 >= 0 and c1 >= 0 and abs(c1 - c0) / c0 > 10 * parity.RTOL, (c0, c1)
    # the other branch is what the model gives when the pedestrian really starts inside the radius
    q["goal_x"] = 1.2 + 0.35 - 1e-7
    c2, _ = ol.oracle_probe_one(p, sc, v, w, parity.MARGINS)
    assert abs(c2 - c1) / c1 < 1e-6, (c1, c2)


def test_collision_flip_changes_validity_both_ways():
    p, sc, lin, ang = G.CASES["c0_hazards_40steps"]()  # a pedestrian walks across the robot's path
    costs, ev, n_ev = ol.oracle_probe_grid(p, sc, lin, ang, (0.0, 0.08, 0.0, 1.0))
    seen = set()
    for k in range(len(costs)):
        for e in ev[k][:n_ev[k]]:
            if e["kind"] != ol.EV_COLLISION:
                continue
            v, w = lin[k // len(ang)], ang[k % len(ang)]
            c, _ = ol.oracle_probe_one(p, sc, v, w, parity.MARGINS, flips=np.array([parity._forced(e)]))
            if e["decision"] == 1:  # the rollout died here: forced on, it survives at least this step
                assert costs[k] == -1.0
                seen.add("revive")
            else:
                assert c == -1.0
                seen.add("kill")
    assert seen == {"kill", "revive"}, seen


def test_theta_flip_moves_the_cost_by_about_its_weight():
    p, sc, lin, ang = _case(n_peds=8, steps=16, seed=1)
    v, w = lin[5], ang[2]
    c0, ev = ol.oracle_probe_one(p, sc, v, w, (0.0, 0.0, 0.3, 1e-3), max_events=64)
    th = [e for e in ev if e["kind"] == ol.EV_THETA]
    assert th, "a 0.3 rad margin should meet angular decisions"
    e = max(th, key=lambda e: e["weight"])
    c1, _ = ol.oracle_probe_one(p, sc, v, w, parity.MARGINS, flips=np.array([parity._forced(e)]))
    assert c1 != c0 and abs(c1 - c0) < 50 * e["weight"] * p.social_weight + 1e-9


def test_two_branch_rule_accepts_branches_and_nothing_else():
    p, sc, lin, ang = _case()
    q = sc.peds[0]
    q["x"], q["y"], q["vx"], q["vy"] = 1.2, 0.25, 0.6, 0.0
    q["goal_x"], q["goal_y"], q["goal_radius"] = 1.2 + 0.35 + 1e-7, 0.25, 0.35
    n = len(lin) * len(ang)
    base, ev, n_ev = ol.oracle_probe_grid(p, sc, lin, ang, parity.MARGINS)
    assert (n_ev > 0).sum() > n // 2  # every rollout sees the pedestrian's step-0 decision
    n0 = len(parity.STATS)
    # (i) the oracle itself, rounded to float like the GPU's cost vector
    st, _ = parity.check_samples(p, sc, lin, ang, np.arange(n), base.astype(np.float32))
    assert st["branch_resolved"] == 0 and st["unresolved"] == 0 and st["near"] == int((n_ev > 0).sum())
    # (ii) an evaluator that took the goal decision the other way on some trajectories
    other = base.copy()
    for k in (10, 23, 40):
        e = [e for e in ev[k][:n_ev[k]] if e["kind"] == ol.EV_GOAL][0]
        other[k], _ = ol.oracle_probe_one(p, sc, lin[k // len(ang)], ang[k % len(ang)], parity.MARGINS,
                                          flips=np.array([parity._forced(e)]))
        assert not parity._match(other[k], base[k])
    st, res = parity.check_samples(p, sc, lin, ang, np.arange(n), other.astype(np.float32))
    assert st["branch_resolved"] == 3 and st["max_flips"] == 1 and np.array_equal(res[[10, 23, 40]], other[[10, 23, 40]])
    # (iii) an evaluator that is simply wrong by 5e-4 on one trajectory: no branch explains it
    bad = base.copy()
    bad[23] *= 1.0 + 5e-4
    with pytest.raises(AssertionError, match="neither the oracle nor any branch"):
        parity.check_samples(p, sc, lin, ang, np.arange(n), bad.astype(np.float32))
    # (iv) validity must match the branch too
    bad = base.copy()
    bad[23] = -1.0
    with pytest.raises(AssertionError):
        parity.check_samples(p, sc, lin, ang, np.arange(n), bad.astype(np.float32))
    del parity.STATS[n0:]  # self-tests of the checker are not GPU statistics


def test_group_contact_event():
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=3, n_w=3, steps=12, n_peds=4)
    p, sc = wl.params(), S.make_scene(wl, 2)
    lin, ang = wl.sample_arrays()
    a, b = sc.peds[0], sc.peds[1]
    a["group_id"] = b["group_id"] = 7
    a["x"], a["y"], a["vx"], a["vy"] = 1.5, 0.5, 0.0, 0.0
    b["x"], b["y"], b["vx"], b["vy"] = 1.5 + 0.7 + 1e-7, 0.5, 0.0, 0.0  # radii 0.35 + 0.35: just not touching
    for q in (a, b):
        q["goal_x"], q["goal_y"] = q["x"] + 2.0, q["y"]
    c0, ev = ol.oracle_probe_one(p, sc, lin[2], ang[1], parity.MARGINS)
    g = [e for e in ev if e["kind"] == ol.EV_GROUP and e["step"] == 0]
    assert len(g) == 1 and (g[0]["a"], g[0]["b"], g[0]["decision"]) == (1, 2, 0)
    c1, _ = ol.oracle_probe_one(p, sc, lin[2], ang[1], parity.MARGINS, flips=np.array([parity._forced(g[0])]))
    assert c1 != c0
