"""CPU: host mirror of SFWPlanner vs the reference's own findBestAction on the branches that need no
scoring (no GPU here), the default 5 x 9 sample sets, and the golden outputs for the scored branches."""
import json
import os

import numpy as np
import pytest

import host_cases as H
import oracle_lib as ol

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "host_golden.json")))


@pytest.mark.parametrize("name", [n for n, c in H.CASES.items() if not c[3]])
def test_unscored_branches_match_reference(name):
    ok, cmd, wp, running, launches, err = H.run_host(name)
    g = GOLD[name]
    assert (ok, list(cmd), wp, running) == (g["ok"], g["cmd"], g["wp_index"], g["running"]), (name, err)
    assert launches == 0
    if ol.have_ref():
        assert H.run_reference(name) == (ok, cmd, wp, running)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", list(H.CASES))
def test_golden_is_what_the_reference_does(name):
    ok, cmd, wp, running = H.run_reference(name)
    g = GOLD[name]
    assert (ok, list(cmd), wp, running) == (g["ok"], g["cmd"], g["wp_index"], g["running"])


def test_default_sample_sets_are_the_references():
    from social_force_window_planner_b200 import scenes as S
    from social_force_window_planner_b200.planner import SFWPlanner
    p, sc = H.make()
    pl = SFWPlanner(p, sc)
    lin, ang = pl.defaultSampleSets()
    pl.close()
    rl, ra = S.reference_sample_arrays(0.7, 0.5)
    assert np.array_equal(lin, rl) and np.array_equal(ang, ra)
    if ol.have_ref():
        import ctypes as C
        l5, a9 = np.zeros(5), np.zeros(9)
        dp = C.POINTER(C.c_double)
        ol.ref().sfw_ref_default_samples(0.7, 0.5, l5.ctypes.data_as(dp), a9.ctypes.data_as(dp))
        assert np.array_equal(lin, l5) and np.array_equal(ang, a9)
