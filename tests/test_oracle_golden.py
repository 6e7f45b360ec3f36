"""CPU: the oracle (oracle/sfw_oracle.c) against the committed golden vectors that the REFERENCE's own
compiled sources produced (tests/golden/make_golden.py), and — where oracle/_ref exists — against the
reference harness live.  Bar: bit-exact doubles (same arithmetic, same order, no fast-math)."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_cases as G
import oracle_lib as ol
from social_force_window_planner_b200._abi import SceneArray

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_matches_reference_golden(name):
    p, sc, lin, ang = G.CASES[name]()
    assert G.scene_crc(sc) == int(GOLD[name + "/crc"][0]), "scene generator drifted from the golden inputs"
    costs, best, _ = ol.oracle_score(p, sc, lin, ang)
    gold = GOLD[name + "/costs"]
    assert np.array_equal(costs, gold), f"max abs diff {np.max(np.abs(costs - gold))}"
    gb = GOLD[name + "/best"]
    assert best.valid == int(gb[0])
    if best.valid:
        assert best.index == int(gb[1]) and best.v == gb[2] and best.w == gb[3]


def test_oracle_mt_equals_single_thread():
    p, sc, lin, ang = G.CASES["c0_hazards_40steps"]()
    a, ba, _ = ol.oracle_score(p, sc, lin, ang)
    b, bb, _ = ol.oracle_score(p, sc, lin, ang, threads=4)
    assert np.array_equal(a, b) and ba.index == bb.index


@pytest.mark.parametrize("kind", ["poly", "point"])
def test_footprint_known_answers(kind):
    case = "c0_hazards_40steps" if kind == "poly" else "point_footprint_hazards"
    p, sc, lin, ang = G.CASES[case]()
    sa = SceneArray([sc])
    got = np.array([ol.oracle().sfw_oracle_footprint_cost(sa.ptr(0), *pose, None) for pose in G.FOOTPRINT_POSES])
    assert np.array_equal(got, GOLD["footprint/" + kind])
    # the codes the reference distinguishes all occur (costmap_model.cpp:26-30)
    if kind == "poly":
        assert {-1.0, -2.0, -3.0} <= set(got.tolist())


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["c0_seed0", "c0_hazards_seed5", "odom_far_from_origin"])
def test_oracle_matches_reference_live(name):
    p, sc, lin, ang = G.CASES[name]()
    oc, ob, _ = ol.oracle_score(p, sc, lin, ang)
    rc, rb = ol.ref_score(p, sc, lin, ang)
    assert np.array_equal(oc, rc)
    assert (ob.valid, ob.index, ob.v, ob.w) == (rb.valid, rb.index, rb.v, rb.w)


def test_bresenham_known_answers():
    """LineIterator (reference include/social_force_window_planner/line_iterator.hpp:37-124)."""
    buf = (C.c_int * 64)()
    n = ol.oracle().sfw_oracle_line_cells(0, 0, 5, 2, buf, 32)
    cells = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
    assert cells[0] == (0, 0) and cells[-1] == (5, 2) and n == 6
    n = ol.oracle().sfw_oracle_line_cells(3, 3, 3, 3, buf, 32)
    assert n == 1 and (buf[0], buf[1]) == (3, 3)
    n = ol.oracle().sfw_oracle_line_cells(2, 7, -1, 0, buf, 32)
    cells = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
    assert n == 8 and cells[0] == (2, 7) and cells[-1] == (-1, 0)
    ys = [c[1] for c in cells]
    assert ys == list(range(7, -1, -1))


def test_closed_form_cost_empty_scene():
    """P = 0, M = 0, free map: cost = w_v |vmax - v_S|/vmax + w_d d^2 + w_a |dtheta|/pi."""
    import math
    from social_force_window_planner_b200 import scenes as S
    wl = S.WORKLOADS["C0"]
    sc = S.make_scene(wl, 0, n_peds=0, n_obstacles=0)
    sc.costmap[:] = 0
    p = wl.params()
    lin, ang = np.array([0.5]), np.array([0.0])
    costs, _, _ = ol.oracle_score(p, sc, lin, ang)
    # straight rollout from v0 = float(0.3) accelerating at 1 m/s^2 towards 0.5
    v, x, dt = float(np.float32(0.3)), 0.0, p.sim_time / wl.steps
    for _ in range(wl.steps):
        v = min(0.5, v + p.max_trans_acc * dt)
        x += v * dt
    d2 = (3.0 - x) ** 2 + 0.5 ** 2
    # normalizeAngle runs in float (sfw_planner.hpp:399-407): mn + fmodf(val - mn, mx - mn)
    f = np.float32
    val, mn, mx = f(math.atan2(0.5, 3.0 - x)), f(-math.pi), f(math.pi)
    dth = abs(float(mn + np.fmod(f(val - mn), f(mx - mn)))) / math.pi
    want = p.vel_weight * abs(p.max_vel_x - v) / p.max_vel_x + p.distance_weight * d2 + p.angle_weight * dth
    assert abs(costs[0] - want) < 1e-9 * want


def test_may_i_stop_golden_is_what_the_reference_does():
    import ctypes as C
    import json
    import os
    import pytest
    import oracle_lib as ol
    from social_force_window_planner_b200._abi import SceneArray
    if not ol.have_ref():
        pytest.skip("oracle/_ref not built")
    import golden_cases as G
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "may_i_stop_golden.json")))
    p, sc, lin, ang = G.CASES["c0_hazards_40steps"]()
    sa = SceneArray([sc])
    for g in gold:
        assert ol.ref().sfw_ref_may_i_stop(C.byref(p), sa.ptr(0), *g["args"], g["dt"]) == g["can_stop"]
