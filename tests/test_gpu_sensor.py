"""-m gpu: sfw_laser_obstacles (CUDA, through the C ABI) vs the oracle's restatement of
SFMSensorInterface::laserCb (reference src/sensor_interface.cpp:103-229).

Bar: same beams kept (count and order identical); coordinates within 1e-6 m — the float cosine/sine
of the scan angle may differ by one float ulp between glibc's cosf and the device's correctly rounded
double routine, everything else is the same IEEE arithmetic."""
import numpy as np
import pytest

import oracle_lib as ol
import sensor_cases as SC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in SC.CASES if n != "empty"])
def test_laser_obstacles_vs_oracle(scorer, name):
    sc = SC.CASES[name]()
    got = scorer.laser_obstacles([sc])[0]
    want = ol.oracle_laser_obstacles(sc)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if len(want):
        assert np.max(np.abs(got - want)) <= 1e-6
    assert scorer.last_kernel == "sfw_laser_kernel"


def test_laser_batch_of_scans(scorer):
    """One launch, one block per scan; ragged beam / people counts; equals the single-scan calls."""
    names = ["room_720", "ragged_257", "no_people", "dense_1440", "all_rejected", "single_beam", "room_tf"]
    scans = [SC.CASES[n]() for n in names]
    outs = scorer.laser_obstacles(scans)
    for n, sc, o in zip(names, scans, outs):
        single = scorer.laser_obstacles([sc])[0]
        assert np.array_equal(o, single), n
        want = ol.oracle_laser_obstacles(sc)
        assert o.shape == want.shape, n


def test_laser_points_feed_the_scorer(scorer):
    """The kept points are SfwScene::obstacles_xy: a scene scored with the device-filtered points equals the
    same scene scored with the oracle-filtered points within the scorer's own tolerance."""
    import dataclasses
    import parity
    from social_force_window_planner_b200 import scenes as S
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=9, n_w=9)
    scene = S.make_scene(wl, 0)
    scan = SC.make_scan(9, n_beams=360, n_people=0)
    scan["people"] = np.stack([scene.peds["x"], scene.peds["y"]], 1)
    pts = scorer.laser_obstacles([scan])[0]
    scene.obstacles = pts[::8].copy()  # a laser leaves hundreds of points; keep the scene small
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [scene], lin, ang)
    print(parity.compare(p, scene, lin, ang, costs[0], best[0]))


import os

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sensor_golden.npz"))
YAW = 4  # agent column


@pytest.mark.parametrize("key", [k[:-len("/agents")] for k in GOLD.files if k.endswith("/agents")])
def test_host_sensor_interface_vs_reference_snapshot(key):
    """Host mirror of SFMSensorInterface (laser filter on the GPU, people/odom callbacks on the host) against
    the agent snapshot the reference's own compiled callbacks produced (committed fixture)."""
    from social_force_window_planner_b200.sensor import sensor_run
    name, ptf = (key[:-len("/people_tf")], True) if key.endswith("/people_tf") else (key, False)
    k = list(SC.CASES).index(name)
    sc = SC.CASES[name]()
    ppl = SC.people_records(sc, k, sc.get("tf") if ptf else None)
    agents, obs, launches = sensor_run(sc, ppl, SC.ODOM, people_has_tf=ptf)
    g_agents, g_obs = GOLD[key + "/agents"], GOLD[key + "/obstacles"]
    assert launches == (1 if len(sc["ranges"]) else 0)
    assert obs.shape == g_obs.shape
    if len(g_obs):
        assert np.max(np.abs(obs - g_obs)) <= 1e-6
    assert agents.shape == g_agents.shape
    cols = [c for c in range(16) if c != YAW]
    # positions / velocities / goals go through the same IEEE expressions as the reference: bit-equal
    assert np.array_equal(agents[:, cols], g_agents[:, cols]), np.argwhere(agents[:, cols] != g_agents[:, cols])[:5]
    # the heading takes a quaternion round trip in the reference (setRPY -> getYaw): compare as angles
    d = np.angle(np.exp(1j * (agents[:, YAW] - g_agents[:, YAW])))
    assert np.max(np.abs(d)) <= 1e-12
