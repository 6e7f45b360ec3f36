"""-m gpu: far-field cutoff of the pedestrians' obstacle force (sfw_set_obstacle_cutoff).

lightsfm sums k/M * exp(-(|p - o| - r)/sigma) over every obstacle point (reached through the
computeForces call at reference src/sfw_planner.cpp:592).  The library stores the points as compact clusters and a
pedestrian pair skips a cluster whose every term is below 2^-24 of the force factor; pedestrians are packed in the
order of the clusters they reach.  Checked here: the oracle bar with the cutoff on (default) AND off, how far the
two cost vectors are from each other, cluster edge cases (counts that are not a multiple of 8, one point, more
than 64 clusters), and that the packed pedestrian order keeps group tags attached to the right pedestrians.
Second half: the block-per-trajectory kernel's force-phase layouts (which warp sums which obstacle clusters and pair
offsets depends on the crowd size) either side of every switch, and its two winner-reduction paths."""
import dataclasses

import numpy as np
import pytest

import golden_cases as G
import parity
from social_force_window_planner_b200 import scenes as S

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fresh():
    from social_force_window_planner_b200.scorer import Scorer
    s = Scorer(0)
    yield s
    s.close()


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(a >= 0, b >= 0)
    both = (a >= 0) & (b >= 0)
    return float(np.max(np.abs(a[both] - b[both]) / np.abs(b[both]))) if both.any() else 0.0


def _wall(n, x0, y0, x1, y1):
    t = np.linspace(0.0, 1.0, n)
    return np.stack([x0 + (x1 - x0) * t, y0 + (y1 - y0) * t], 1)


@pytest.mark.parametrize("name,seed,kw", [("C0", 0, {}), ("C0", 2, {"hazards": True}), ("C1", 0, {}), ("C1", 1, {}),
                                          ("C3", 1, {})])
def test_cutoff_on_and_off_meet_the_oracle_bar(fresh, name, seed, kw):
    wl = dataclasses.replace(S.WORKLOADS[name], n_v=20, n_w=21)
    sc = S.make_scene(wl, seed, **kw)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    c_on, b_on = fresh.score(p, [sc], lin, ang)
    skip = fresh.obstacle_skip_fraction
    st_on = parity.compare(p, sc, lin, ang, c_on[0], b_on[0])
    fresh.set_obstacle_cutoff(0.0)
    c_off, b_off = fresh.score(p, [sc], lin, ang)
    assert fresh.obstacle_skip_fraction == 0.0
    st_off = parity.compare(p, sc, lin, ang, c_off[0], b_off[0])
    d = _rel(c_on[0], c_off[0])
    # the skipped terms are below one FP32 ulp of the force factor: the two runs differ by rounding noise
    assert d <= 2e-5, d
    print(name, seed, "skip fraction at start", round(skip, 3), "on-vs-off", d, st_on["max_rel_clear"],
          st_off["max_rel_clear"])


@pytest.mark.parametrize("m", [1, 7, 8, 9, 33, 100, 523])
def test_cluster_edge_counts(fresh, m):
    """Obstacle counts around the cluster size; 523 points = 66 clusters (reach masks wider than one word)."""
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=9, n_w=10, n_peds=9, steps=24)
    sc = S.make_scene(wl, 5)
    rng = np.random.default_rng(m)
    # a near wall (in reach of everybody), a far wall (out of reach of most), some scattered points
    pts = np.concatenate([_wall(m, 1.2, -2.0, 1.6, 2.0), _wall(m, -9.0, -6.0, -9.0, 6.0),
                          rng.uniform(-7.0, 7.0, (m, 2))])
    sc.obstacles = pts[rng.permutation(len(pts))[:m]].copy()
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = fresh.score(p, [sc], lin, ang)
    st = parity.compare(p, sc, lin, ang, costs[0], best[0])
    fresh.set_policy(fresh.POLICY_THROUGHPUT)  # the thread-per-trajectory kernel on the same scene
    costs_t, best_t = fresh.score(p, [sc], lin, ang)
    st_t = parity.compare(p, sc, lin, ang, costs_t[0], best_t[0])
    print(m, fresh.obstacle_skip_fraction, st["max_rel_clear"], st_t["max_rel_clear"])


def test_every_cluster_out_of_reach(fresh):
    """All obstacle points > 20 m from every pedestrian: the pedestrians' sums skip everything, the robot's
    (never skipped) still sees them; both cutoff settings agree with the oracle."""
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=8, n_w=9, n_peds=6)
    sc = S.make_scene(wl, 1)
    sc.obstacles = _wall(40, 30.0, -5.0, 30.0, 5.0)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = fresh.score(p, [sc], lin, ang)
    assert fresh.obstacle_skip_fraction == 1.0
    print(parity.compare(p, sc, lin, ang, costs[0], best[0]))


def test_packed_pedestrian_order_keeps_group_tags(fresh):
    """Grouped pedestrians on both sides of the reach boundary of a wall: the packed order differs from the
    caller's, the group table must follow it (golden: the oracle on the caller's order)."""
    wl = dataclasses.replace(S.WORKLOADS["C1"], n_v=10, n_w=11, steps=40, n_peds=13)
    p, sc, lin, ang = G._grouped(wl, 2, {3: [0, 5, 11], 8: [2, 12], 9: [7]})
    sc.obstacles = np.concatenate([_wall(24, 4.5, -3.0, 4.5, 3.0), _wall(11, -1.0, 3.5, 1.0, 3.5)])
    costs, best = fresh.score(p, [sc], lin, ang)
    assert 0.0 < fresh.obstacle_skip_fraction < 1.0
    st = parity.compare(p, sc, lin, ang, costs[0], best[0])
    tagged = costs.copy()
    sc.peds["group_id"] = -1
    costs0, _ = fresh.score(p, [sc], lin, ang)
    assert not np.array_equal(costs0, tagged), "group tags must change the result"
    print(st)


def test_batch_with_different_obstacle_counts(fresh):
    """Scenes of one batch with 0 / 5 / 32 / 70 obstacle points and different crowd sizes: same results as one
    call per scene (bit for bit under the throughput policy)."""
    fresh.set_policy(fresh.POLICY_THROUGHPUT)
    wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=12, n_w=12)
    scs = []
    for i, (m, n_peds) in enumerate([(0, 4), (5, 7), (32, 10), (70, 3)]):
        scs.append(S.make_scene(wl, 10 + i, n_obstacles=m, n_peds=n_peds))
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = fresh.score(p, scs, lin, ang)
    for i, sc in enumerate(scs):
        c1, b1 = fresh.score(p, [sc], lin, ang)
        assert np.array_equal(c1[0], costs[i])
        parity.compare(p, sc, lin, ang, costs[i], best[i])


# ---- block-per-trajectory kernel: layout boundaries of its force phase -------------------------
@pytest.mark.parametrize("n_peds", [6, 7, 8, 63, 64, 65, 66, 128, 129, 255, 256, 257, 258])
def test_crowd_kernel_layout_boundaries(fresh, n_peds):
    """Owner layout with obstacle helpers (up to 3 pairs, 33 .. 128 pairs), spread layout (4 .. 32 pairs), plain owner
    layout (more than 128 pairs): crowd sizes either side of every switch, odd and even, against the oracle."""
    fresh.set_policy(fresh.POLICY_LATENCY)
    wl = dataclasses.replace(S.WORKLOADS["C2"], n_v=3, n_w=4, steps=10, n_peds=n_peds,
                             ped_r_max=4.0 if n_peds < 70 else 8.0, ped_sep=0.5)
    sc = S.make_scene(wl, n_peds, n_obstacles=37)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = fresh.score(p, [sc], lin, ang)
    assert fresh.last_kernel == "sfw_score_crowd"
    print(n_peds, parity.compare(p, sc, lin, ang, costs[0], best[0]))


@pytest.mark.parametrize("n_scenes", [1, 5, 64, 65, 70])
def test_crowd_kernel_winner_reduction_paths(fresh, n_scenes):
    """Small batches are reduced by the scorer's last block, larger ones (more than 64 scenes or 8192 cost values) by
    the arg-min kernel: same winners as one call per scene, launch after launch (the counters re-arm themselves)."""
    fresh.set_policy(fresh.POLICY_LATENCY)
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=4, n_w=5, steps=8, n_peds=9)
    scs = [S.make_scene(wl, 40 + i) for i in range(n_scenes)]
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = fresh.score(p, scs, lin, ang)
    for rep in range(3):  # sfw_run again on the staged batch
        fresh.run()
        c2, b2 = fresh.download()
        assert np.array_equal(c2, costs) and np.array_equal(b2, best)
    for i in (0, n_scenes // 2, n_scenes - 1):
        c1, b1 = fresh.score(p, [scs[i]], lin, ang)
        assert np.array_equal(c1[0], costs[i]) and b1[0] == best[i]
        parity.compare(p, scs[i], lin, ang, costs[i], best[i])


def test_crowd_kernel_big_grid_unfused_argmin(fresh):
    """More than 8192 cost values in one scene: arg-min kernel; winner consistent with the cost vector."""
    fresh.set_policy(fresh.POLICY_LATENCY)
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=96, n_w=96, steps=6, n_peds=4)
    sc = S.make_scene(wl, 3)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = fresh.score(p, [sc], lin, ang)
    print(parity.compare(p, sc, lin, ang, costs[0], best[0]))
