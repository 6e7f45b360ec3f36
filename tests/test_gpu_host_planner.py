"""-m gpu: the host mirror of SFWPlanner::findBestAction (C++ control flow + one CUDA scoring call per
tick) against what the reference's own findBestAction returned for the same scene / plan / odometry
(tests/golden/host_golden.json, produced by oracle/_ref)."""
import json
import os

import pytest

import host_cases as H

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "host_golden.json")))


@pytest.mark.parametrize("name", list(H.CASES))
def test_find_best_action_matches_reference(name):
    ok, cmd, wp, running, launches, err = H.run_host(name)
    g = GOLD[name]
    assert (ok, list(cmd), wp, running) == (g["ok"], g["cmd"], g["wp_index"], g["running"]), (name, err)
    needs_gpu = H.CASES[name][3]
    assert (launches > 0) == needs_gpu
