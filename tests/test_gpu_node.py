"""-m gpu: the host mirror of the plugin class SFWPlannerNode — sensor callbacks, setPlan,
computeVelocityCommands, with the laser filter and every scored trajectory on the GPU — against what the
reference's WHOLE plugin returned on the same messages (tests/golden/node_golden.json, produced by the
reference's sources compiled unmodified: tests/golden/make_node_golden.py)."""
import json
import os

import pytest

import node_cases as N

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "node_golden.json")))


@pytest.mark.parametrize("name", N.NAMES)
def test_compute_velocity_commands_matches_reference_plugin(name):
    from social_force_window_planner_b200.node import node_run
    cmd, status, left, reached, launches = node_run(**N.make(name))
    g = GOLD[name]
    assert status.tolist() == g["status"], name
    assert cmd.tolist() == g["cmd"], name            # commands are sample-set values: exact
    assert left.tolist() == g["plan_left"], name     # same pruning of the global plan
    assert reached.tolist() == g["goal_reached"], name
    if name != "empty_plan":
        assert launches >= 1                          # at least the laser kernel ran
