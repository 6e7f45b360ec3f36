"""CPU: the oracle's restatement of SFMSensorInterface::laserCb (reference src/sensor_interface.cpp:103-229)
against a straightforward numpy evaluation of the same rules, and its invariants."""
import numpy as np

import oracle_lib as ol
import sensor_cases as SC


@np.errstate(invalid="ignore", over="ignore")
def _numpy_laser(scan, max_dist=3.0, person_radius=0.35):
    r = np.asarray(scan["ranges"], dtype=np.float32)
    n = len(r)
    ang = np.empty(n, dtype=np.float32)
    a = np.float32(scan["angle_min"])
    for i in range(n):  # float accumulation, reference :118,127
        ang[i] = a
        a = np.float32(a + np.float32(scan["angle_increment"]))
    ok = np.isfinite(r) & (r < np.float32(max_dist))
    x = (r * np.cos(ang, dtype=np.float32)).astype(np.float64)
    y = (r * np.sin(ang, dtype=np.float32)).astype(np.float64)
    if scan.get("tf") is not None:
        tx, ty, yaw = scan["tf"]
        c, s = np.cos(yaw), np.sin(yaw)
        x, y = c * x - s * y + tx, s * x + c * y + ty
    ppl = np.asarray(scan.get("people", np.zeros((0, 2)))).reshape(-1, 2)
    for px, py in ppl:
        d = np.hypot((x - px).astype(np.float32), (y - py).astype(np.float32))
        ok &= ~(d <= np.float32(person_radius))
    return np.stack([x[ok], y[ok]], 1)


def test_oracle_laser_matches_numpy_rules():
    for name, mk in SC.CASES.items():
        sc = mk()
        got = ol.oracle_laser_obstacles(sc)
        want = _numpy_laser(sc)
        assert got.shape == want.shape, name
        if len(want):
            # numpy's float32 cos/sin may differ from glibc's cosf/sinf by an ulp
            assert np.max(np.abs(got - want)) <= 2e-6, name


def test_oracle_laser_invariants():
    sc = SC.CASES["room_720"]()
    pts = ol.oracle_laser_obstacles(sc)
    assert 0 < len(pts) < len(sc["ranges"])
    # nothing survives inside a person's disc, nothing beyond max_obstacle_dist
    for px, py in sc["people"]:
        assert (np.hypot(pts[:, 0] - px, pts[:, 1] - py) > 0.35 - 1e-6).all()
    assert (np.hypot(pts[:, 0], pts[:, 1]) < 3.0 + 1e-6).all()
    # people do remove beams
    no_ppl = dict(sc, people=np.zeros((0, 2)))
    assert len(ol.oracle_laser_obstacles(no_ppl)) > len(pts)
    assert len(ol.oracle_laser_obstacles(SC.CASES["all_rejected"]())) == 0
    assert len(ol.oracle_laser_obstacles(SC.CASES["empty"]())) == 0


import os

import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sensor_golden.npz"))


@pytest.mark.parametrize("name", list(SC.CASES))
def test_oracle_laser_equals_reference_laserCb(name):
    """The oracle's laserCb restatement against what the reference's own src/sensor_interface.cpp produced
    (committed fixture, tests/golden/make_sensor_golden.py): same points, bit for bit."""
    got = ol.oracle_laser_obstacles(SC.CASES[name]())
    want = GOLD[name + "/obstacles"]
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.skipif(not ol.have_ref_sensor(), reason="oracle/_ref/libsfw_ref_sensor.so not built")
def test_sensor_golden_is_what_the_reference_does():
    for k, (name, mk) in enumerate(SC.CASES.items()):
        sc = mk()
        agents, obs = ol.ref_sensor_run(sc, SC.people_records(sc, k), SC.ODOM)
        assert np.array_equal(agents, GOLD[name + "/agents"]) and np.array_equal(obs, GOLD[name + "/obstacles"])
