"""CPU: the committed whole-plugin fixture is what the reference's own plugin does (when oracle/_ref is built),
and the host library exports the plugin mirror."""
import json
import os

import numpy as np
import pytest

import node_cases as N
import oracle_lib as ol

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "node_golden.json")))


@pytest.mark.skipif(not ol.have_ref_node(), reason="oracle/_ref/libsfw_ref_node.so not built")
@pytest.mark.parametrize("name", N.NAMES)
def test_node_golden_is_what_the_reference_plugin_does(name):
    cmd, status, left, reached = ol.ref_node_run(**N.make(name))
    g = GOLD[name]
    assert cmd.tolist() == g["cmd"] and status.tolist() == g["status"]
    assert left.tolist() == g["plan_left"] and reached.tolist() == g["goal_reached"]


def test_host_library_exports_the_plugin_mirror():
    from social_force_window_planner_b200.planner import host_lib
    h = host_lib()
    for sym in ("sfwn_node_run", "sfws_sensor_run", "sfwh_find_best_action", "sfwh_get_markers"):
        assert hasattr(h, sym), sym


def test_fixture_covers_every_outcome():
    st = [s for g in GOLD.values() for s in g["status"]]
    assert {1, 0, -1} <= set(st)
    assert any(any(g["goal_reached"]) for g in GOLD.values())
    assert GOLD["pruning_and_costmap_cut"]["plan_left"][0] < 49  # the passed head of the plan was pruned
