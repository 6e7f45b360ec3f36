"""Synthetic laser scans shared by the CPU and GPU sensor-interface tests (reference
src/sensor_interface.cpp:103-229 has no fixtures of its own)."""
import math

import numpy as np

from social_force_window_planner_b200.scenes import SplitMix64


def make_scan(seed: int, n_beams: int = 720, n_people: int = 6, tf=None, fov=(-2.3561945, 2.3561945)):
    """A room-like scan: walls 1.5-6 m away, NaN / inf / out-of-range returns, and people standing IN
    FRONT of the walls so that some beams end on them (those are the points laserCb removes)."""
    rng = SplitMix64(7000 + seed)
    inc = np.float32((fov[1] - fov[0]) / max(n_beams - 1, 1))
    people = np.array([[rng.uniform(0.6, 2.8) * math.cos(a), rng.uniform(0.6, 2.8) * math.sin(a)]
                       for a in [rng.uniform(fov[0], fov[1]) for _ in range(n_people)]]).reshape(-1, 2)
    ranges = np.empty(n_beams, dtype=np.float32)
    ang = np.float32(fov[0])
    for i in range(n_beams):
        a = float(ang)
        r = 2.5 + 1.5 * math.sin(3.0 * a) + 0.8 * math.cos(11.0 * a) + rng.uniform(-0.02, 0.02)
        # a person between the sensor and the wall: the beam stops on the person's disc
        for px, py in people:
            along = px * math.cos(a) + py * math.sin(a)
            perp = abs(-px * math.sin(a) + py * math.cos(a))
            if along > 0 and perp < 0.25:
                r = min(r, along - math.sqrt(0.25 ** 2 - perp ** 2))
        u = rng.uniform()
        if u < 0.02:
            r = float("nan")
        elif u < 0.04:
            r = float("inf")
        elif u < 0.06:
            r = 3.0  # exactly max_obstacle_dist: rejected (strict <)
        ranges[i] = r
        ang = np.float32(ang + inc)
    if tf is not None:
        # people are reported in the controller frame
        c, s = math.cos(tf[2]), math.sin(tf[2])
        people = np.stack([c * people[:, 0] - s * people[:, 1] + tf[0], s * people[:, 0] + c * people[:, 1] + tf[1]], 1)
    sc = {"ranges": ranges, "angle_min": float(np.float32(fov[0])), "angle_increment": float(inc), "people": people}
    if tf is not None:
        sc["tf"] = tf
    return sc


CASES = {
    "room_720": lambda: make_scan(0),
    "room_tf": lambda: make_scan(1, tf=(12.5, -3.25, 0.8)),
    "no_people": lambda: make_scan(2, n_people=0),
    "ragged_257": lambda: make_scan(3, n_beams=257, n_people=11),
    "single_beam": lambda: make_scan(4, n_beams=1, n_people=1),
    "dense_1440": lambda: make_scan(5, n_beams=1440, n_people=20, tf=(-1.0, 2.0, -2.9), fov=(-math.pi, math.pi)),
    "all_rejected": lambda: {"ranges": np.full(64, np.inf, dtype=np.float32), "angle_min": -1.0,
                             "angle_increment": 0.03, "people": np.zeros((0, 2))},
    "empty": lambda: {"ranges": np.zeros(0, dtype=np.float32), "angle_min": 0.0, "angle_increment": 0.01},
}


def people_records(scan: dict, seed: int, message_frame_tf=None):
    """people_msgs/People rows {x, y, yaw, vx, vy, wz, id, group} for the people of a scan.  ``scan['people']``
    holds controller-frame positions; with ``message_frame_tf`` the message is expressed in another frame
    (the inverse of that planar transform is applied) so that peopleCb has to bring it back."""
    rng = SplitMix64(9000 + seed)
    xy = np.asarray(scan.get("people", np.zeros((0, 2))), dtype=np.float64).reshape(-1, 2)
    rows = np.zeros((len(xy), 8))
    for i, (x, y) in enumerate(xy):
        speed = rng.uniform(0.0, 1.4) if i % 3 else rng.uniform(0.0, 0.08)  # every third one almost still
        hd = rng.uniform(-math.pi, math.pi)
        vx, vy = speed * math.cos(hd), speed * math.sin(hd)
        yaw = rng.uniform(-math.pi, math.pi)
        if message_frame_tf is not None:
            tx, ty, tyaw = message_frame_tf
            c, s = math.cos(tyaw), math.sin(tyaw)
            x, y = c * (x - tx) + s * (y - ty), -s * (x - tx) + c * (y - ty)
            vx, vy = c * vx + s * vy, -s * vx + c * vy
            yaw -= tyaw
        rows[i] = (x, y, yaw, vx, vy, rng.uniform(-0.5, 0.5), 10 + i, (i // 2) if i % 4 < 2 else -1)
    return rows


ODOM = (0.25, -0.4, 0.6, 0.35, 0.02, -0.15)
