"""-m gpu: THE DROP-IN.  The reference's plugin class SFWPlannerNode (the nav2_core::Controller that sfw_plugin.xml
registers) and its SFMSensorInterface, compiled from the reference's sources WITHOUT ANY EDIT against
plugin/include/social_force_window_planner/sfw_planner.hpp, run on top of plugin/src/sfw_planner.cpp + the CUDA
scorer (oracle/Makefile target `dropin` -> oracle/_ref/libsfw_dropin_node.so, built where /root/reference exists and
shipped to the GPU box like the other _ref objects).  Replayed: the whole-plugin fixtures the reference's own,
unmodified plugin produced (tests/golden/node_golden.json) — sensor callbacks, setPlan, computeVelocityCommands."""
import json
import os

import pytest

import node_cases as N
import oracle_lib as ol

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "node_golden.json")))


@pytest.mark.skipif(not ol.have_dropin_node(), reason="oracle/_ref/libsfw_dropin_node.so not built (needs /root/reference)")
@pytest.mark.parametrize("name", N.NAMES)
def test_reference_node_on_the_b200_planner_matches_the_reference_plugin(name):
    cmd, status, left, reached = ol.dropin_node_run(**N.make(name))
    g = GOLD[name]
    assert status.tolist() == g["status"], name
    assert cmd.tolist() == g["cmd"], name            # commands are sample-set values / closed-form proposals: exact
    assert left.tolist() == g["plan_left"], name     # the node's own pruning of the global plan
    assert reached.tolist() == g["goal_reached"], name


def test_the_dropin_really_runs_the_cuda_library():
    """The binary must have libsfw_b200.so mapped (no hidden CPU path) and must not contain the reference's planner
    core (scoreTrajectory is the reference's private scorer; our class has no such member)."""
    import subprocess
    if not ol.have_dropin_node():
        pytest.skip("drop-in binary not built")
    ol.dropin_node_run(**N.make("far_goal"))
    maps = open("/proc/self/maps").read()
    assert "libsfw_b200.so" in maps and "libsfw_dropin_node.so" in maps
    syms = subprocess.run(["nm", "-DC", ol.DROPIN_NODE_SO], capture_output=True, text=True).stdout
    assert "SFWPlannerNode::computeVelocityCommands" in syms      # the reference's node is in there ...
    assert "SFWPlanner::findBestAction" in syms                   # ... on our planner core ...
    assert "scoreTrajectory" not in syms and "CostmapModel::footprintCost" not in syms  # ... not on the reference's
