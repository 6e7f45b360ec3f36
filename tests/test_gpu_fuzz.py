"""-m gpu: randomized shapes against the oracle — pedestrian / obstacle counts (incl. odd crowds, the
64-pedestrian switch between the two kernels), step counts, footprints, group tags, map sizes, robot states.
Same bar as tests/test_gpu_parity.py."""
import dataclasses

import numpy as np
import pytest

import parity
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scenes import SplitMix64

pytestmark = pytest.mark.gpu


def _random_case(seed):
    rng = SplitMix64(424242 + seed)
    pick = lambda lo, hi: lo + int(rng.uniform() * (hi - lo + 1))  # noqa: E731
    n_peds = [0, 1, 2, 3, 7, 12, 19, 20, 33, 63, 64, 65, 66, 97][pick(0, 13)]
    steps = pick(6, 48) if n_peds < 60 else pick(6, 20)
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=pick(2, 9), n_w=pick(3, 11), steps=steps, n_peds=n_peds,
                             map_w=pick(100, 260), map_h=pick(100, 260), ped_r_max=5.5 if n_peds > 30 else None,
                             ped_sep=0.5 if n_peds > 30 else 0.8)
    fp = None
    k = pick(0, 3)
    if k == 1:
        fp = np.zeros((0, 2))
    elif k == 2:
        fp = np.array([[0.32, 0.22], [-0.28, 0.22], [-0.28, -0.22], [0.32, -0.22]])
    elif k == 3:
        fp = S.circle_footprint(rng.uniform(0.2, 0.5), pick(5, 24))
    sc = S.make_scene(wl, seed, n_obstacles=pick(0, 48), footprint=fp, hazards=rng.uniform() < 0.5,
                      robot_xy=(rng.uniform(-50, 50), rng.uniform(-50, 50)) if rng.uniform() < 0.3 else (0.0, 0.0),
                      robot_theta=rng.uniform(-3.1, 3.1))
    # group tags on some pedestrians (ids 0..2; lone tags stay inert), a few without a goal
    for j in range(n_peds):
        u = rng.uniform()
        if u < 0.35:
            sc.peds[j]["group_id"] = pick(0, 2)
        if rng.uniform() < 0.1:
            sc.peds[j]["has_goal"] = 0
    p = wl.params()
    p.max_trans_acc = rng.uniform(0.2, 2.0)
    p.max_rot_acc = rng.uniform(0.2, 2.0)
    p.robot_radius = float(np.float32(rng.uniform(0.25, 0.45)))
    lin, ang = wl.sample_arrays(max_vel_x=rng.uniform(0.4, 1.0), max_vel_th=rng.uniform(0.3, 1.2))
    p.max_vel_x = float(lin[-1])
    return wl, p, sc, lin, ang


@pytest.mark.parametrize("seed", range(28))
def test_random_scene_vs_oracle(scorer, seed):
    wl, p, sc, lin, ang = _random_case(seed)
    costs, best = scorer.score(p, [sc], lin, ang)
    st = parity.compare(p, sc, lin, ang, costs[0], best[0])
    print(seed, wl.n_peds, "peds", len(sc.obstacles), "obst", wl.steps, "steps", scorer.last_kernel, st)


@pytest.mark.parametrize("seed", range(0, 28, 2))
def test_random_scene_latency_policy(seed):
    """The same random shapes through the block-per-trajectory kernel (SFW_POLICY_LATENCY), which AUTO only picks
    for small grids and big crowds: empty crowds, single pedestrians, point footprints, no obstacles ..."""
    from social_force_window_planner_b200.scorer import Scorer
    wl, p, sc, lin, ang = _random_case(seed)
    s2 = Scorer(0)
    try:
        s2.set_policy(Scorer.POLICY_LATENCY)
        costs, best = s2.score(p, [sc], lin, ang)
        assert s2.last_kernel == "sfw_score_crowd"
        st = parity.compare(p, sc, lin, ang, costs[0], best[0])
        # and the thread-per-trajectory kernel on the same scene agrees with it (same model, other summation order)
        if wl.n_peds <= 64:
            s2.set_policy(Scorer.POLICY_THROUGHPUT)
            c2, b2 = s2.score(p, [sc], lin, ang)
            clear = (costs[0] >= 0) & (c2[0] >= 0)
            if clear.any():
                assert np.max(np.abs(costs[0][clear] - c2[0][clear]) / np.abs(c2[0][clear])) < 5e-2
    finally:
        s2.close()
    print(seed, wl.n_peds, "peds", st)


@pytest.mark.parametrize("seed", range(10))
def test_random_grid_with_forced_prefix_sharing(seed):
    """Mid-size random grids with rollout prefix sharing forced on (``set_prefix_sharing(2)``), through the kernel
    family AUTO picks and through the thread-per-trajectory family: against the oracle, and bit for bit against the
    same run without sharing.  Random accelerations move the fork steps around; hazards kill shared paths; group
    tags select the thread-per-path writers, their absence the warp-per-path ones."""
    from social_force_window_planner_b200.scorer import Scorer
    rng = SplitMix64(99 + seed)
    pick = lambda lo, hi: lo + int(rng.uniform() * (hi - lo + 1))  # noqa: E731
    n_peds = [0, 1, 5, 12, 19, 20, 33, 64][pick(0, 7)]
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=pick(16, 40), n_w=pick(64, 90), steps=pick(12, 40), n_peds=n_peds,
                             map_w=pick(160, 260), map_h=pick(160, 260), ped_r_max=5.5 if n_peds > 30 else None,
                             ped_sep=0.5 if n_peds > 30 else 0.8)
    sc = S.make_scene(wl, 1000 + seed, n_obstacles=pick(0, 40), hazards=rng.uniform() < 0.4)
    if seed % 3 == 0:
        for j in range(min(n_peds, 6)):
            sc.peds[j]["group_id"] = j // 3
    r = list(sc.robot)
    r[3] = float(np.float32(rng.uniform(0.0, 0.5)))
    r[5] = float(np.float32(rng.uniform(-0.5, 0.5)))
    r[10] = r[3]
    sc.robot = tuple(r)
    p = wl.params()
    p.max_trans_acc = rng.uniform(0.3, 1.5)
    p.max_rot_acc = rng.uniform(0.3, 1.5)
    lin, ang = wl.sample_arrays(max_vel_x=rng.uniform(0.5, 1.0), max_vel_th=rng.uniform(0.5, 1.2))
    p.max_vel_x = float(lin[-1])
    s2 = Scorer(0)
    try:
        for pol in (Scorer.POLICY_AUTO, Scorer.POLICY_THROUGHPUT):
            s2.set_policy(pol)
            s2.set_prefix_sharing(2)
            c_on, b_on = s2.score(p, [sc], lin, ang)
            k_on = s2.last_kernel
            s2.set_prefix_sharing(0)
            c_off, b_off = s2.score(p, [sc], lin, ang)
            assert "share" in k_on and "share" not in s2.last_kernel, (k_on, s2.last_kernel)
            assert np.array_equal(c_on, c_off) and np.array_equal(b_on, b_off), k_on
            if pol == Scorer.POLICY_AUTO:
                st = parity.compare(p, sc, lin, ang, c_on[0], b_on[0])
    finally:
        s2.close()
    print(seed, wl.n_v, "x", wl.n_w, wl.steps, "steps", n_peds, "peds", k_on, st)
