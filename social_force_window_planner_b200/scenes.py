"""Deterministic synthetic planning scenes of the shapes BASELINE.json names.

The reference ships no fixtures or benchmark inputs; the scene recipe below is the one SURVEY.md
section 8(d) fixes (robot at the origin, 0.05 m costmap with lethal boxes + inflation rings,
pedestrians in an annulus with naive goals ``pos + 2 s * vel`` as
reference src/sensor_interface.cpp:493-503 builds them, M = 32 obstacle points as the laser
callback would leave them, 16-gon footprint as nav2 synthesises from ``robot_radius``).
Everything is seeded by ``1000 + scene_index`` through a splitmix64 stream implemented here, so
oracle, reference harness and the CUDA path read byte-identical inputs on any machine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from ._abi import PED_DTYPE, SfwParams, default_params

_MASK = (1 << 64) - 1


class SplitMix64:
    """Tiny portable PRNG (public-domain splitmix64 recurrence)."""

    def __init__(self, seed: int):
        self.s = seed & _MASK

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        return lo + (hi - lo) * ((self.next_u64() >> 11) * (1.0 / 9007199254740992.0))


@dataclass
class Scene:
    """One planning scene in host (double) precision; mirrors ``SfwScene``."""

    # x, y, theta, vx, vy, vtheta, wpx, wpy, agent_x, agent_y, agent_vx, agent_vy, agent_radius
    robot: tuple
    costmap: np.ndarray  # uint8 [size_y, size_x]
    resolution: float
    origin_x: float
    origin_y: float
    peds: np.ndarray  # PED_DTYPE [P]
    obstacles: np.ndarray  # float64 [M, 2]
    footprint: np.ndarray  # float64 [F, 2]
    meta: dict = field(default_factory=dict)


@dataclass
class Workload:
    """A named BASELINE.json configuration."""

    name: str
    n_v: int
    n_w: int
    steps: int
    n_peds: int
    map_w: int
    map_h: int
    n_scenes: int = 1
    n_obstacles: int = 32
    ped_r_max: float | None = None
    ped_sep: float = 0.8

    @property
    def samples(self) -> int:
        return self.n_v * self.n_w

    def params(self) -> SfwParams:
        p = default_params()
        p.sim_granularity = 0.025
        p.sim_time = self.steps * 0.025
        return p

    def sample_arrays(self, max_vel_x: float = 0.7, max_vel_th: float = 0.5):
        return sample_arrays(self.n_v, self.n_w, max_vel_x, max_vel_th)


# BASELINE.json configs[0..4] (C0..C4)
WORKLOADS = {
    "C0": Workload("C0", 21, 21, 20, 5, 200, 200),
    "C1": Workload("C1", 256, 256, 64, 20, 400, 400),
    "C2": Workload("C2", 128, 128, 128, 500, 400, 400, ped_r_max=9.5, ped_sep=0.5),
    "C3": Workload("C3", 64, 64, 32, 10, 200, 200, n_scenes=4096),
    "C4": Workload("C4", 1024, 1024, 32, 10, 800, 800),
}


def sample_arrays(n_v: int, n_w: int, max_vel_x: float = 0.7, max_vel_th: float = 0.5):
    """Generalised explicit (linvel, angvel) sample sets.

    The reference keeps explicit arrays too (sfw_planner.hpp:370-371, built at
    sfw_planner.cpp:65-85 as 5 x 9); the named configs only change their length.
    """
    if n_v > 1:
        lin = np.array([max_vel_x * i / (n_v - 1) for i in range(n_v)], dtype=np.float64)
    else:
        lin = np.array([max_vel_x], dtype=np.float64)
    if n_w > 1:
        ang = np.array([-max_vel_th + 2.0 * max_vel_th * j / (n_w - 1) for j in range(n_w)],
                       dtype=np.float64)
    else:
        ang = np.array([0.0], dtype=np.float64)
    return lin, ang


def reference_sample_arrays(max_vel_x: float = 0.7, max_vel_th: float = 0.5):
    """The reference's shipped 5 x 9 sets, in its order (sfw_planner.cpp:65-85)."""
    lstep = max_vel_x / 4
    lin = np.array([i * lstep for i in range(5)], dtype=np.float64)
    astep = max_vel_th / 4
    ang = [0.0]
    for i in range(1, 5):
        ang.append(i * astep)
        ang.append(i * (-astep))
    return lin, np.array(ang, dtype=np.float64)


def circle_footprint(radius: float = 0.35, n: int = 16) -> np.ndarray:
    """nav2's makeFootprintFromRadius shape: n points on a circle [external]."""
    return np.array([[radius * math.cos(2 * math.pi * k / n), radius * math.sin(2 * math.pi * k / n)]
                     for k in range(n)], dtype=np.float64)


def _make_costmap(rng: SplitMix64, w: int, h: int, res: float, ox: float, oy: float):
    cm = np.zeros((h, w), dtype=np.uint8)
    n_boxes = max(1, w // 25)
    ring = 0.5
    xs = ox + (np.arange(w) + 0.5) * res
    ys = oy + (np.arange(h) + 0.5) * res
    for _ in range(n_boxes):
        for _try in range(100):
            sx = rng.uniform(0.2, 1.0)
            sy = rng.uniform(0.2, 1.0)
            cx = rng.uniform(ox + 1.0, ox + w * res - 1.0)
            cy = rng.uniform(oy + 1.0, oy + h * res - 1.0)
            # nearest point of the box to the robot (origin) must be >= 1.2 m away
            nx = min(max(0.0, cx - sx / 2), cx + sx / 2)
            ny = min(max(0.0, cy - sy / 2), cy + sy / 2)
            if math.hypot(nx, ny) >= 1.2:
                break
        else:
            continue
        x0, x1, y0, y1 = cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2
        ix0 = max(0, int((x0 - ring - ox) / res) - 1)
        ix1 = min(w, int((x1 + ring - ox) / res) + 2)
        iy0 = max(0, int((y0 - ring - oy) / res) - 1)
        iy1 = min(h, int((y1 + ring - oy) / res) + 2)
        gx = xs[ix0:ix1][None, :]
        gy = ys[iy0:iy1][:, None]
        dx = np.maximum(np.maximum(x0 - gx, gx - x1), 0.0)
        dy = np.maximum(np.maximum(y0 - gy, gy - y1), 0.0)
        d = np.hypot(dx, dy)
        cost = np.where(d <= 0.0, 254, np.where(d <= ring, np.floor(252.0 * np.exp(-6.0 * d)), 0))
        sub = cm[iy0:iy1, ix0:ix1]
        np.maximum(sub, cost.astype(np.uint8), out=sub)
    cm[0, :] = 255
    cm[-1, :] = 255
    cm[:, 0] = 255
    cm[:, -1] = 255
    return cm


def _make_peds(rng: SplitMix64, n: int, r_max: float, sep: float) -> np.ndarray:
    peds = np.zeros(n, dtype=PED_DTYPE)
    pos = []
    k = 0
    tries = 0
    cur_sep = sep
    while k < n:
        tries += 1
        if tries > 20000:  # crowded: relax the separation a little rather than spin forever
            cur_sep *= 0.9
            tries = 0
        # uniform over the annulus area
        r = math.sqrt(rng.uniform(1.0, r_max * r_max))
        a = rng.uniform(-math.pi, math.pi)
        x, y = r * math.cos(a), r * math.sin(a)
        if any((x - px) ** 2 + (y - py) ** 2 < cur_sep * cur_sep for px, py in pos):
            continue
        speed = rng.uniform(0.3, 1.3)
        hd = rng.uniform(-math.pi, math.pi)
        vx, vy = speed * math.cos(hd), speed * math.sin(hd)
        pos.append((x, y))
        p = peds[k]
        p["x"], p["y"], p["vx"], p["vy"] = x, y, vx, vy
        # naive goal = pos + naive_goal_time * vel (reference sensor_interface.cpp:493-500)
        p["goal_x"], p["goal_y"] = x + 2.0 * vx, y + 2.0 * vy
        p["goal_radius"] = 0.35
        p["desired_velocity"] = 1.0
        p["radius"] = 0.35
        p["has_goal"] = 1
        p["group_id"] = -1
        p["id"] = k + 1
        k += 1
    return peds


def _make_obstacles(cm: np.ndarray, res: float, ox: float, oy: float, m: int) -> np.ndarray:
    """Centres of the m lethal cells nearest the robot within 3 m, padded on a 2.5 m circle."""
    if m == 0:
        return np.zeros((0, 2), dtype=np.float64)
    iy, ix = np.nonzero(cm == 254)
    x = ox + (ix + 0.5) * res
    y = oy + (iy + 0.5) * res
    d = np.hypot(x, y)
    keep = d <= 3.0
    x, y, d = x[keep], y[keep], d[keep]
    order = np.lexsort((ix[keep], iy[keep], d))[:m]
    pts = [(float(x[i]), float(y[i])) for i in order]
    k = 0
    while len(pts) < m:
        a = 2 * math.pi * k / m
        pts.append((2.5 * math.cos(a), 2.5 * math.sin(a)))
        k += 1
    return np.array(pts, dtype=np.float64)


def _stamp_box(cm, res, ox, oy, x0, x1, y0, y1, ring=0.3):
    """Lethal box [x0,x1]x[y0,y1] with an exponential inflation ring, max-merged into ``cm``."""
    h, w = cm.shape
    xs = ox + (np.arange(w) + 0.5) * res
    ys = oy + (np.arange(h) + 0.5) * res
    dx = np.maximum(np.maximum(x0 - xs[None, :], xs[None, :] - x1), 0.0)
    dy = np.maximum(np.maximum(y0 - ys[:, None], ys[:, None] - y1), 0.0)
    d = np.hypot(dx, dy)
    cost = np.where(d <= 0.0, 254, np.where(d <= ring, np.floor(252.0 * np.exp(-6.0 * d)), 0))
    np.maximum(cm, cost.astype(np.uint8), out=cm)


def make_scene(workload: Workload, scene_index: int = 0, *, n_peds: int | None = None,
               n_obstacles: int | None = None, footprint: np.ndarray | None = None,
               robot_xy=(0.0, 0.0), robot_theta: float = 0.0, hazards: bool = False) -> Scene:
    """Build scene ``scene_index`` of a workload (seed = 1000 + scene_index).

    ``robot_xy``/``robot_theta`` translate/rotate nothing but the robot start pose and the whole
    world with it (used by tests to exercise large odom coordinates).
    """
    rng = SplitMix64(1000 + scene_index)
    res = 0.05
    w, h = workload.map_w, workload.map_h
    ox, oy = -w * res / 2, -h * res / 2
    cm = _make_costmap(rng, w, h, res, ox, oy)
    p = workload.n_peds if n_peds is None else n_peds
    r_max = workload.ped_r_max if workload.ped_r_max is not None else min(4.0, 0.45 * w * res)
    peds = _make_peds(rng, p, r_max, workload.ped_sep)
    if hazards:
        # a lethal box the faster straight rollouts drive into, an unknown (255) patch on the right
        # and (when there are pedestrians) one that walks across the robot's path: exercises every
        # rejection branch of scoreTrajectory (sfw_planner.cpp:545-573, :613-627)
        _stamp_box(cm, res, ox, oy, 0.62, 0.85, 0.05, 0.40)
        cm[int((-0.62 - oy) / res):int((-0.50 - oy) / res), int((0.30 - ox) / res):int((0.45 - ox) / res)] = 255
        if p > 0:
            q = peds[0]
            q["x"], q["y"], q["vx"], q["vy"] = 0.45, -0.75, 0.0, 0.8
            q["goal_x"], q["goal_y"] = 0.45, -0.75 + 2.0 * 0.8
    m = workload.n_obstacles if n_obstacles is None else n_obstacles
    obs = _make_obstacles(cm, res, ox, oy, m)
    fp = circle_footprint() if footprint is None else np.asarray(footprint, dtype=np.float64)
    tx, ty = robot_xy
    if tx != 0.0 or ty != 0.0:
        peds["x"] += tx
        peds["y"] += ty
        peds["goal_x"] += tx
        peds["goal_y"] += ty
        obs = obs + np.array([tx, ty])
    f32 = lambda v: float(np.float32(v))  # noqa: E731  (caller-side narrowing, sfw_planner.cpp:145-152)
    rx, ry, rt = f32(tx), f32(ty), f32(robot_theta)
    robot = (rx, ry, rt, f32(0.3), 0.0, 0.0, tx + 3.0, ty + 0.5, tx, ty, 0.3, 0.0, 0.35)
    return Scene(robot=robot, costmap=cm, resolution=res, origin_x=ox + tx, origin_y=oy + ty,
                 peds=peds, obstacles=obs, footprint=fp,
                 meta={"workload": workload.name, "scene_index": scene_index,
                       "seed": 1000 + scene_index})


def make_scenes(workload: Workload, n: int | None = None, first: int = 0):
    n = workload.n_scenes if n is None else n
    return [make_scene(workload, first + i) for i in range(n)]
