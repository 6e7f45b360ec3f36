"""-m gpu: CUDA scorer (through the C ABI) vs the CPU oracle / reference golden vectors on identical inputs.

Tolerance (BASELINE.json north_star): per-trajectory cost within 1e-4 relative (parity.RTOL), validity
identical, arg-min index identical whenever the runner-up differs by more than 1e-3."""
import dataclasses
import os

import numpy as np
import pytest

import golden_cases as G
import parity
from social_force_window_planner_b200 import scenes as S

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))


def _run(scorer, wl, scene_index=0, **kw):
    sc = S.make_scene(wl, scene_index, **kw)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [sc], lin, ang)
    return parity.compare(p, sc, lin, ang, costs[0], best[0])


def _same_costs(a, b, same_kernel):
    """Bit-equal when the same kernel produced both; otherwise (a lone small grid goes to the block-per-trajectory
    kernel) the same model with a different summation order: identical validity, 1e-5 relative."""
    a, b = np.asarray(a), np.asarray(b)
    if same_kernel:
        return np.array_equal(a, b)
    both = (a >= 0) & (b >= 0)
    return np.array_equal(a >= 0, b >= 0) and np.allclose(a[both], b[both], rtol=1e-5, atol=0)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_c0_cpu_ref_config(scorer, seed):
    print(_run(scorer, S.WORKLOADS["C0"], seed))


@pytest.mark.parametrize("seed", [0, 1])
def test_c1_shape_reduced_grid(scorer, seed):
    wl = dataclasses.replace(S.WORKLOADS["C1"], n_v=24, n_w=24)
    print(_run(scorer, wl, seed))


@pytest.mark.parametrize("name", list(G.CASES))
def test_golden_cases_vs_reference_outputs(scorer, name):
    """GPU cost vector against what the reference's own compiled sources produced (committed fixture)."""
    p, sc, lin, ang = G.CASES[name]()
    costs, best = scorer.score(p, [sc], lin, ang)
    gold = GOLD[name + "/costs"]
    g = costs[0].astype(np.float64)
    st = parity.compare(p, sc, lin, ang, costs[0], best[0])
    if st["near"] == 0:
        assert np.array_equal(g >= 0, gold >= 0)
        both = (g >= 0) & (gold >= 0)
        if both.any():
            assert np.max(np.abs(g[both] - gold[both]) / np.abs(gold[both])) <= parity.RTOL
        assert np.array_equal(g == -2.0, gold == -2.0)
    print(name, st)


def test_winner_trajectory_points(scorer):
    """Trajectory::x_pts_/y_pts_/th_pts_ of a sample (reference src/trajectory.cpp:36-40) vs the oracle."""
    import ctypes as C
    import oracle_lib as ol
    from social_force_window_planner_b200._abi import SceneArray
    p, sc, lin, ang = G.CASES["c0_hazards_40steps"]()
    costs, best = scorer.score(p, [sc], lin, ang)
    for idx in (int(best[0]["index"]), 440, 220, 3):
        pts, n = scorer.trajectory_points(0, idx)
        buf = np.zeros((64, 3))
        n_o = C.c_uint32(0)
        sa = SceneArray([sc])
        ol.oracle().sfw_oracle_score_trajectory(C.byref(p), None, sa.ptr(0), lin[idx // len(ang)], 0.0,
                                                ang[idx % len(ang)], p.max_trans_acc, 0.0, p.max_rot_acc,
                                                buf.ctypes.data_as(C.POINTER(C.c_double)), 64, C.byref(n_o), None)
        assert n == n_o.value
        assert np.array_equal(pts, buf[:n]), "rollout poses must be bit-identical (FP64 kinematics)"


def test_batch_of_scenes_equals_single_calls(scorer):
    wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=16, n_w=16)
    scs = S.make_scenes(wl, 12)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, scs, lin, ang)
    batch_kernel = scorer.last_kernel
    for k in (0, 5, 11):
        c1, b1 = scorer.score(p, [scs[k]], lin, ang)
        if scorer.last_kernel == batch_kernel:
            assert np.array_equal(c1[0], costs[k]) and b1[0] == best[k]
        else:
            # a lone small grid is latency bound and goes to the block-per-trajectory kernel:
            # same model, different summation order
            both = (c1[0] >= 0) & (costs[k] >= 0)
            assert np.array_equal(c1[0] >= 0, costs[k] >= 0)
            assert np.allclose(c1[0][both], costs[k][both], rtol=1e-5, atol=0)
        parity.compare(p, scs[k], lin, ang, costs[k], best[k])
        parity.compare(p, scs[k], lin, ang, c1[0], b1[0])


def test_ragged_batch(scorer):
    """Scenes of one batch may differ in pedestrian / obstacle count, map size and footprint."""
    wl = dataclasses.replace(S.WORKLOADS["C0"], steps=24)
    scs = [S.make_scene(wl, 0), S.make_scene(wl, 1, n_peds=0, n_obstacles=0),
           S.make_scene(dataclasses.replace(wl, map_w=120, map_h=90), 2, n_peds=11, n_obstacles=7),
           S.make_scene(wl, 3, footprint=np.zeros((0, 2)), hazards=True)]
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, scs, lin, ang)
    for k, sc in enumerate(scs):
        parity.compare(p, sc, lin, ang, costs[k], best[k])


def test_row_slabs_merge_to_full_grid(scorer):
    """Multi-GPU row-slab sharding on one device: slab winners merged == full-grid winner."""
    from social_force_window_planner_b200 import sharding
    p, sc, lin, ang = G.CASES["c0_hazards_seed5"]()
    full_costs, full_best = scorer.score(p, [sc], lin, ang)
    recs = []
    scorer.upload(p, [sc], lin, ang)
    for r in range(3):
        b, e = sharding.block_partition(len(lin), 3, r)
        scorer.set_row_slab(b, e)
        scorer.run()
        c, best = scorer.download()
        assert np.array_equal(c[0][b * len(ang):e * len(ang)], full_costs[0][b * len(ang):e * len(ang)])
        assert (c[0][:b * len(ang)] == -2.0).all() and (c[0][e * len(ang):] == -2.0).all()
        recs.append(best[0])
    m = sharding.merge_winners(np.array(recs))
    assert m == full_best[0]


def test_full_size_c1_properties(scorer):
    """BASELINE.json configs[1] at full size: properties that need no oracle run —
    (i) the arg-min equals a host arg-min over the returned cost vector under the reference's order,
    (ii) a strided sub-grid scored on its own gives bit-identical costs (trajectories are independent),
    (iii) re-running is deterministic."""
    import ctypes as C
    import oracle_lib as ol
    from social_force_window_planner_b200._abi import SfwBest
    wl = S.WORKLOADS["C1"]
    sc = S.make_scene(wl, 0)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [sc], lin, ang)
    full_kernel = scorer.last_kernel
    costs2, best2 = scorer.score(p, [sc], lin, ang)
    assert np.array_equal(costs, costs2) and best[0] == best2[0]
    sb = SfwBest()
    dp = C.POINTER(C.c_double)
    c64 = np.ascontiguousarray(costs[0], dtype=np.float64)
    ol.oracle().sfw_oracle_argmin(c64.ctypes.data_as(dp), lin.ctypes.data_as(dp), len(lin), ang.ctypes.data_as(dp),
                                  len(ang), C.byref(sb))
    assert (sb.valid, sb.index) == (int(best[0]["valid"]), int(best[0]["index"]))
    ri, ci = np.arange(0, wl.n_v, 17), np.arange(0, wl.n_w, 13)
    sub, _ = scorer.score(p, [sc], lin[ri], np.ascontiguousarray(ang[ci]))
    assert _same_costs(sub[0].reshape(len(ri), len(ci)), costs[0].reshape(wl.n_v, wl.n_w)[np.ix_(ri, ci)],
                       scorer.last_kernel == full_kernel)
    # and the sub-grid against the oracle
    parity.compare(p, sc, lin[ri], np.ascontiguousarray(ang[ci]), sub[0], _[0])


def test_errors_are_reported_not_thrown(scorer):
    from social_force_window_planner_b200.scorer import SfwError
    p, sc, lin, ang = G.CASES["c0_seed0"]()
    with pytest.raises(SfwError) as ei:
        scorer.score(p, [sc], lin[:0], ang)
    assert ei.value.code == -1
    q = S.WORKLOADS["C0"].params()
    q.sim_granularity = 0.0
    with pytest.raises(SfwError):
        scorer.score(q, [sc], lin, ang)
    # the context stays usable
    costs, best = scorer.score(p, [sc], lin, ang)
    assert int(best[0]["valid"]) == 1


# ---- dense crowds: block-per-trajectory kernel (sfw_crowd.cu) ---------------------------------------
@pytest.mark.parametrize("n_peds", [65, 150, 151])
def test_crowd_kernel_parity(scorer, n_peds):
    wl = dataclasses.replace(S.WORKLOADS["C2"], n_v=6, n_w=6, steps=32, n_peds=n_peds, ped_r_max=6.0)
    sc = S.make_scene(wl, 0)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [sc], lin, ang)
    assert scorer.last_kernel == "sfw_score_crowd"
    # with hundreds of agents most rollouts pass within the 5e-6 margin of SOME goal pop / sign flip; they are
    # held to the same 1e-4 on the branch the GPU took (parity.check_samples)
    print(parity.compare(p, sc, lin, ang, costs[0], best[0]))


def test_crowd_kernel_hazards_and_points(scorer):
    """Costmap rejections, collisions and the recorded-point counts through the crowd kernel."""
    import ctypes as C
    import oracle_lib as ol
    from social_force_window_planner_b200._abi import SceneArray
    wl = dataclasses.replace(S.WORKLOADS["C0"], steps=40, n_peds=70, n_v=9, n_w=9)
    sc = S.make_scene(wl, 3, hazards=True)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [sc], lin, ang)
    assert scorer.last_kernel == "sfw_score_crowd"
    st = parity.compare(p, sc, lin, ang, costs[0], best[0])
    assert 0 < st["valid"] < st["n"] - 1, st
    sa = SceneArray([sc])
    for idx in range(0, 81, 7):
        if lin[idx // 9] == 0.0 and ang[idx % 9] == 0.0:
            continue
        pts, n = scorer.trajectory_points(0, idx)
        n_o = C.c_uint32(0)
        ol.oracle().sfw_oracle_score_trajectory(C.byref(p), None, sa.ptr(0), lin[idx // 9], 0.0, ang[idx % 9],
                                                p.max_trans_acc, 0.0, p.max_rot_acc, None, 0, C.byref(n_o), None)
        assert n == n_o.value, (idx, n, n_o.value)


def test_crowd_batch_and_slab(scorer):
    from social_force_window_planner_b200 import sharding
    wl = dataclasses.replace(S.WORKLOADS["C2"], n_v=4, n_w=5, steps=16, n_peds=80, ped_r_max=5.0)
    scs = [S.make_scene(wl, k) for k in range(3)]
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, scs, lin, ang)
    for k in range(3):
        c1, b1 = scorer.score(p, [scs[k]], lin, ang)
        assert np.array_equal(c1[0], costs[k]) and b1[0] == best[k]
    scorer.upload(p, [scs[1]], lin, ang)
    recs = []
    for r in range(2):
        b, e = sharding.block_partition(len(lin), 2, r)
        scorer.set_row_slab(b, e)
        scorer.run()
        c, bb = scorer.download()
        assert np.array_equal(c[0][b * 5:e * 5], costs[1][b * 5:e * 5])
        recs.append(bb[0])
    assert sharding.merge_winners(np.array(recs)) == best[1]


@pytest.mark.parametrize("shift", [(0.0, 0.0), (-0.7, 0.0), (0.3, -1.07), (-0.75, 0.8)])
def test_free_space_shortcut_at_map_edges(scorer, shift):
    """The footprint check skips rasterisation where the whole +-fp_rc neighbourhood is free AND inside
    the map.  A border-less 60x60 map (every edge cell cost 0) shifted so that rollouts leave it: only
    the in-map test of the shortcut stands between a vertex off the map (-3 in the reference,
    src/costmap_model.cpp:36,56-60) and a wrong "free" answer.  Plus isolated cost cells the footprint
    outline may or may not touch."""
    wl = dataclasses.replace(S.WORKLOADS["C0"], steps=40, map_w=60, map_h=60, n_peds=3)
    sc = S.make_scene(wl, 2, n_obstacles=4)
    cm = np.zeros((60, 60), dtype=np.uint8)
    cm[30, 44] = 254   # on the straight path: 0.7 m ahead of the start pose
    cm[37, 36] = 200   # inside the swept disc, high but legal
    cm[22, 33] = 253   # inscribed-inflated value: allowed along footprint edges, rejected in point mode
    cm[12, 12] = 255
    sc.costmap = cm
    sc.origin_x += shift[0]
    sc.origin_y += shift[1]
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [sc], lin, ang)
    st = parity.compare(p, sc, lin, ang, costs[0], best[0])
    if shift != (0.0, 0.0):
        assert st["valid"] < st["n"] - 1, st  # some rollouts really do leave the map
    print(shift, st)


def test_crowd_kernel_with_groups(scorer):
    """Group forces (lightsfm computeGroupForce) through the block-per-trajectory kernel."""
    wl = dataclasses.replace(S.WORKLOADS["C2"], n_v=5, n_w=5, steps=24, n_peds=72, ped_r_max=6.0)
    groups = {g: list(range(4 * g, 4 * g + 4)) for g in range(0, 12, 2)}
    groups[40] = [60, 61]
    groups[41] = [70]
    p, sc, lin, ang = G._grouped(wl, 3, groups)
    costs, best = scorer.score(p, [sc], lin, ang)
    assert scorer.last_kernel == "sfw_score_crowd"
    st = parity.compare(p, sc, lin, ang, costs[0], best[0])
    sc.peds["group_id"] = -1
    costs0, _ = scorer.score(p, [sc], lin, ang)
    assert not np.array_equal(costs0, costs), "group tags must change the result"
    print(st)


@pytest.mark.parametrize("n_peds", [0, 1, 5, 7, 20, 33, 64])
def test_crowd_kernel_small_blocks_parity(n_peds):
    """The 128-thread instantiation of the block-per-trajectory kernel (chosen when it saves a wave: a 21 x 21 grid
    is two waves of 256-thread blocks, one of 128-thread blocks): owner layout (up to 3 pairs), spread layout
    (4 .. 32 pairs), odd crowds, against the oracle; a row slab of the same grid keeps the block size and stays
    bit-identical."""
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C0"], steps=24, n_peds=n_peds)
    sc = S.make_scene(wl, 2, hazards=n_peds >= 5)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    s = Scorer(0)
    try:
        s.set_policy(Scorer.POLICY_LATENCY)
        costs, best = s.score(p, [sc], lin, ang)
        assert s.last_kernel == "sfw_score_crowd"
        assert s.block_threads == 128, s.block_threads
        print(parity.compare(p, sc, lin, ang, costs[0], best[0]))
        s.upload(p, [sc], lin, ang)
        s.set_row_slab(3, 9)  # 126 samples: one wave of either block size
        s.run()
        c2, _ = s.download()
        assert s.block_threads == 128
        assert np.array_equal(c2[0][3 * len(ang):9 * len(ang)], costs[0][3 * len(ang):9 * len(ang)])
    finally:
        s.close()


def test_crowd_kernel_small_blocks_with_groups():
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C0"], steps=24, n_peds=24)
    groups = {0: [0, 1, 2, 3], 2: [8, 9, 10], 5: [20, 21]}
    p, sc, lin, ang = G._grouped(wl, 3, groups)
    s = Scorer(0)
    try:
        s.set_policy(Scorer.POLICY_LATENCY)
        costs, best = s.score(p, [sc], lin, ang)
        assert s.last_kernel == "sfw_score_crowd" and s.block_threads == 128
        print(parity.compare(p, sc, lin, ang, costs[0], best[0]))
    finally:
        s.close()


# ---- BASELINE.json configs[2..4] at FULL size: size-independent properties + oracle spot checks -------
def _spot_check(p, sc, lin, ang, costs, picks):
    """A handful of trajectories of a big grid against the oracle's single-trajectory scorer (same two-branch
    rule and tolerance as parity.compare)."""
    picks = np.asarray(picks, dtype=np.int64)
    return parity.check_samples(p, sc, lin, ang, picks, np.asarray(costs)[picks])[0]


def _host_argmin(costs, lin, ang):
    import ctypes as C
    import oracle_lib as ol
    from social_force_window_planner_b200._abi import SfwBest
    sb = SfwBest()
    dp = C.POINTER(C.c_double)
    c64 = np.ascontiguousarray(costs, dtype=np.float64)
    ol.oracle().sfw_oracle_argmin(c64.ctypes.data_as(dp), lin.ctypes.data_as(dp), len(lin), ang.ctypes.data_as(dp),
                                  len(ang), C.byref(sb))
    return sb.valid, sb.index


def test_full_size_c4_fine_sweep(scorer):
    """configs[4]: 1024x1024 samples, 32 steps, 10 pedestrians, 800x800 costmap.  The full grid's arg-min against
    a host arg-min of its own cost vector, and a 64 x 64 = 4096-sample strided sub-grid of the FULL run's cost
    vector against the oracle (two-branch rule, 1e-4)."""
    wl = S.WORKLOADS["C4"]
    sc = S.make_scene(wl, 0)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [sc], lin, ang)
    full_kernel = scorer.last_kernel
    assert costs.shape == (1, 1024 * 1024)
    assert _host_argmin(costs[0], lin, ang) == (int(best[0]["valid"]), int(best[0]["index"]))
    ri, ci = np.arange(3, wl.n_v, 16), np.arange(5, wl.n_w, 16)
    assert len(ri) * len(ci) >= 4096
    picks = (ri[:, None] * wl.n_w + ci[None, :]).reshape(-1)
    st, _ = parity.check_samples(p, sc, lin, ang, picks, costs[0][picks])
    assert st["valid"] > 1000, st
    # a sub-grid scored on its own gives the same bits (trajectories are independent)
    ri2, ci2 = ri[::6], ci[::5]
    sub, sb = scorer.score(p, [sc], lin[ri2], np.ascontiguousarray(ang[ci2]))
    assert _same_costs(sub[0].reshape(len(ri2), len(ci2)), costs[0].reshape(wl.n_v, wl.n_w)[np.ix_(ri2, ci2)],
                       scorer.last_kernel == full_kernel)
    print(st)


def test_full_size_c3_scene_batch(scorer):
    """configs[3] at one GPU's share of the 8-GPU run: 512 independent scenes, 64x64 samples, 32 steps, 10 peds.
    Batch results must equal single-scene calls bit for bit (scenes are independent), a few against the oracle."""
    wl = S.WORKLOADS["C3"]
    scs = S.make_scenes(wl, 512)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, scs, lin, ang)
    assert costs.shape == (512, 4096)
    for k in (0, 137, 511):
        c1, b1 = scorer.score(p, [scs[k]], lin, ang)
        assert np.array_equal(c1[0], costs[k]) and b1[0] == best[k]
        assert _host_argmin(costs[k], lin, ang) == (int(best[k]["valid"]), int(best[k]["index"]))
    _spot_check(p, scs[300], lin, ang, costs[300], [0, 1, 63, 64, 2047, 2048 + 31, 4095])
    parity.compare(p, scs[77], lin[::9], np.ascontiguousarray(ang[::7]),
                   costs[77].reshape(64, 64)[::9, ::7].reshape(-1),
                   scorer.score(p, [scs[77]], lin[::9], np.ascontiguousarray(ang[::7]))[1][0])


C2_GOLD = os.path.join(os.path.dirname(__file__), "golden", "c2_rows.npz")


def test_full_size_c2_rows_vs_oracle(scorer):
    """configs[2] at FULL size (128x128 samples, 128 steps, 500 pedestrians): 8 whole linvel rows = 1024
    trajectories through the row-slab interface against the oracle's committed values for exactly these
    trajectories (tests/golden/c2_rows.npz: cost + near decisions per trajectory, ~5 core-seconds each, made by
    tests/golden/make_c2_rows.py).  Same two-branch rule and 1e-4 as everywhere; a trajectory that needs another
    branch is re-run through the oracle here.  Rows outside the slab stay SKIPPED."""
    gold = np.load(C2_GOLD)
    wl = S.WORKLOADS["C2"]
    sc = S.make_scene(wl, 0)
    assert G.scene_crc(sc) == int(gold["crc"][0]), "scene generator drifted from the fixture"
    assert np.array_equal(gold["margins"], np.array(parity.MARGINS))
    p = wl.params()
    lin, ang = wl.sample_arrays()
    scorer.upload(p, [sc], lin, ang)
    rows = [int(r) for r in gold["rows"]]
    spans, start = [], rows[0]
    for a, b in zip(rows, rows[1:] + [None]):  # consecutive rows share a launch
        if b != a + 1:
            spans.append((start, a + 1))
            start = b
    n_w = wl.n_w
    for b, e in spans:
        scorer.set_row_slab(b, e)
        scorer.run()
        costs, best = scorer.download()
        assert scorer.last_kernel == "sfw_score_crowd"
        c = costs[0].reshape(wl.n_v, n_w)
        assert (c[:b] == -2.0).all() and (c[e:] == -2.0).all()
        for r in range(b, e):
            k = rows.index(r)
            idx = np.arange(r * n_w, (r + 1) * n_w)
            st, _ = parity.check_samples(p, sc, lin, ang, idx, c[r], oracle_costs=gold["costs"][k],
                                         events=gold["events"][k], n_events=gold["n_events"][k],
                                         label=f"C2 full size, linvel row {r}")
            print(r, st)
        # the slab's winner is the arg-min of the slab's own cost vector under the reference's order
        assert _host_argmin(costs[0], lin, ang) == (int(best[0]["valid"]), int(best[0]["index"]))


def test_kernel_policy(scorer):
    """AUTO sends a lone small grid to the block-per-trajectory kernel, THROUGHPUT pins the thread-per-trajectory
    one (bit-identical to what a large batch computes for the same scene), LATENCY always takes the former."""
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C0"], steps=40, n_peds=20)
    sc = S.make_scene(wl, 0)
    p = wl.params()
    lin, ang = S.reference_sample_arrays()
    s2 = Scorer(0)
    try:
        res = {}
        for name, pol in (("auto", Scorer.POLICY_AUTO), ("throughput", Scorer.POLICY_THROUGHPUT), ("latency", Scorer.POLICY_LATENCY)):
            s2.set_policy(pol)
            res[name] = s2.score(p, [sc], lin, ang) + (s2.last_kernel,)
            parity.compare(p, sc, lin, ang, res[name][0][0], res[name][1][0])
        assert res["auto"][2] == "sfw_score_crowd" and res["latency"][2] == "sfw_score_crowd"
        assert res["throughput"][2].startswith("sfw_score_small")
        assert np.array_equal(res["auto"][0], res["latency"][0])
        # THROUGHPUT: the same scene inside a big batch gives the very same bits
        s2.set_policy(Scorer.POLICY_THROUGHPUT)
        big, _ = s2.score(p, [sc] * 400, lin, ang)
        assert np.array_equal(big[0], res["throughput"][0][0]) and np.array_equal(big[399], big[0])
    finally:
        s2.close()


@pytest.mark.parametrize("policy", ["throughput", "latency"])
def test_long_horizon_without_staged_window(scorer, policy):
    """Reach (max speed x sim_time + footprint radius) beyond what fits the 48 KB shared-memory window: the
    footprint check reads the costmap from global memory and the free-space shortcut is off."""
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C0"], n_v=4, n_w=5, steps=340, n_peds=3, map_w=420, map_h=420)
    sc = S.make_scene(wl, 2)
    p = wl.params()
    lin, ang = wl.sample_arrays(max_vel_x=1.0)
    p.max_vel_x = 1.0
    s2 = Scorer(0)
    try:
        s2.set_policy(Scorer.POLICY_THROUGHPUT if policy == "throughput" else Scorer.POLICY_LATENCY)
        costs, best = s2.score(p, [sc], lin, ang)
        st = parity.compare(p, sc, lin, ang, costs[0], best[0])
        assert 0 < st["valid"] < st["n"] - 1, st  # some rollouts survive 8.5 s, some hit a box
        print(s2.last_kernel, st)
    finally:
        s2.close()


# ---- rollout prefix sharing -------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n_scenes", [("C4", 1), ("C3", 64), ("C1", 1), ("C1", 3)])
def test_prefix_sharing_is_bit_identical(name, n_scenes):
    """Dense grids start every sample from the shared state of its fork point (SfwShareDev): same arithmetic in
    the same order, so the cost vector, the recorded-point counts and the winners must equal the unshared run bit
    for bit — and the shared run must actually have been taken.  C1 x 1 is the one-wave flavour: chunks of
    fork-sorted samples dealt over blocks and schedulers, path records written by the warp-per-path kernel."""
    from social_force_window_planner_b200.scorer import Scorer
    wl = S.WORKLOADS[name]
    scs = S.make_scenes(wl, n_scenes)
    if n_scenes > 1:  # different odometry per scene: per-scene fork tables
        for k, sc in enumerate(scs):
            r = list(sc.robot)
            r[3] = float(np.float32(0.05 + 0.6 * (k % 7) / 6.0))
            r[5] = float(np.float32(-0.4 + 0.8 * (k % 5) / 4.0))
            r[10] = r[3]
            sc.robot = tuple(r)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    s2 = Scorer(0)
    try:
        costs_on, best_on = s2.score(p, scs, lin, ang)
        k_on = s2.last_kernel
        pts_on = [s2.trajectory_points(0, i)[1] for i in (0, 777, wl.samples - 1)]
        s2.set_prefix_sharing(False)
        costs_off, best_off = s2.score(p, scs, lin, ang)
        k_off = s2.last_kernel
        pts_off = [s2.trajectory_points(0, i)[1] for i in (0, 777, wl.samples - 1)]
    finally:
        s2.close()
    assert "share" in k_on and "share" not in k_off, (k_on, k_off)
    assert np.array_equal(costs_on, costs_off)
    assert np.array_equal(best_on, best_off)
    assert pts_on == pts_off
    print(name, k_on, "valid", float((costs_on >= 0).mean()))


@pytest.mark.parametrize("n_v,n_w,n_scenes,n_peds,groups", [(100, 77, 1, 20, False), (37, 41, 5, 7, False),
                                                           (64, 96, 2, 12, True), (33, 1000, 1, 3, False)])
def test_prefix_sharing_forced_on_ragged_grids(n_v, n_w, n_scenes, n_peds, groups):
    """Sharing forced (mode 2) through the thread-per-trajectory kernel on shapes the cost model would not pick:
    sample counts that are no multiple of 32 (a partial last chunk), several scenes in one wave, odd pedestrian
    counts (a dummy pair half), pedestrian groups (thread-per-path writers instead of warp-per-path)."""
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C1"], n_v=n_v, n_w=n_w, n_peds=n_peds)
    # the last scene has a lethal box ahead and a pedestrian crossing the robot's path: shared paths die too
    scs = [S.make_scene(wl, i, hazards=(n_scenes > 1 and i == n_scenes - 1)) for i in range(n_scenes)]
    for k, sc in enumerate(scs):
        r = list(sc.robot)
        r[3] = float(np.float32(0.1 + 0.1 * k))
        r[5] = float(np.float32(-0.3 + 0.2 * k))
        r[10] = r[3]
        sc.robot = tuple(r)
        if groups:
            sc.peds["group_id"][:9] = np.arange(9) // 3
    p = wl.params()
    lin, ang = wl.sample_arrays()
    s2 = Scorer(0)
    try:
        s2.set_policy(Scorer.POLICY_THROUGHPUT)
        s2.set_prefix_sharing(2)
        costs_on, best_on = s2.score(p, scs, lin, ang)
        k_on = s2.last_kernel
        s2.set_prefix_sharing(0)
        costs_off, best_off = s2.score(p, scs, lin, ang)
        k_off = s2.last_kernel
    finally:
        s2.close()
    assert "share" in k_on and "share" not in k_off, (k_on, k_off)
    assert np.array_equal(costs_on, costs_off)
    assert np.array_equal(best_on, best_off)
    assert (costs_on >= 0).any()
    if n_scenes > 1:
        assert (costs_on[-1] == -1.0).any(), "the hazards scene should kill trajectories (and shared paths)"
    _spot_check(p, scs[-1], lin, ang, costs_on[-1], [0, n_w + 1, n_v * n_w - 1])


def test_prefix_sharing_crowd_kernel_is_bit_identical():
    """The block-per-trajectory kernel shares rollout prefixes the same way (records = pedestrian state +
    social work + collision bookkeeping; rollout and footprint are recomputed per item)."""
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C2"], n_v=64, n_w=64, steps=40, n_peds=200, ped_r_max=7.0)
    sc = S.make_scene(wl, 1)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    s2 = Scorer(0)
    try:
        costs_on, best_on = s2.score(p, [sc], lin, ang)
        k_on = s2.last_kernel
        n_on = [s2.trajectory_points(0, i)[1] for i in (1, 2000, 4095)]
        s2.set_prefix_sharing(False)
        costs_off, best_off = s2.score(p, [sc], lin, ang)
        k_off = s2.last_kernel
        n_off = [s2.trajectory_points(0, i)[1] for i in (1, 2000, 4095)]
    finally:
        s2.close()
    assert k_on == "sfw_score_crowd,share" and k_off == "sfw_score_crowd", (k_on, k_off)
    assert np.array_equal(costs_on, costs_off) and np.array_equal(best_on, best_off) and n_on == n_off
    frac = float((costs_on >= 0).mean())
    assert 0.02 < frac < 0.98, frac  # collisions and survivors both present: dead records are exercised
    _spot_check(p, sc, lin, ang, costs_on[0], [5, 700, 2080, 4000])


def test_may_i_stop_matches_reference(scorer):
    """sfw_may_i_stop against what the reference's own (private, unreachable) SFWPlanner::mayIStop returned on
    the hazards scene (tests/golden/may_i_stop_golden.json, from oracle/_ref)."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "may_i_stop_golden.json")))
    p, sc, lin, ang = G.CASES["c0_hazards_40steps"]()
    scorer.upload(p, [sc], lin, ang)
    seen = set()
    for g in gold:
        ok, steps = scorer.may_i_stop(0, *g["args"], g["dt"])
        assert int(ok) == g["can_stop"], g
        seen.add(g["can_stop"])
        if g["args"][0] > 0 and ok:  # braking from v at 1 m/s^2 takes ceil(v / (a dt)) steps (+1 when the
            n = int(np.ceil(g["args"][0] / (p.max_trans_acc * g["dt"]) - 1e-9))  # subtractions leave a residue)
            assert steps in (n, n + 1), (g, steps)
    assert seen == {0, 1}


def test_parallel_packing_is_bit_identical():
    """A batch is packed by the context's host workers (sfw_set_host_threads), one scene per work item, and shipped
    in pieces.  Ragged scenes (pedestrian / obstacle counts, map sizes, footprints, group tags in some scenes only):
    1, 3 and 8 workers must produce the same bits, and the same as single-scene calls."""
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=40, n_w=40)  # 1600 samples: sharing tables are built too
    scs = []
    for k in range(37):
        w2 = dataclasses.replace(wl, map_w=120 + 8 * (k % 5), map_h=100 + 12 * (k % 4))
        sc = S.make_scene(w2, k, n_peds=(k * 7) % 23, n_obstacles=(k * 5) % 41, hazards=(k % 6 == 0),
                          footprint=None if k % 3 else S.circle_footprint(0.3, 7 + k % 9))
        if k % 4 == 1 and len(sc.peds) >= 6:
            sc.peds["group_id"][:6] = np.arange(6) // 3
        scs.append(sc)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    res = {}
    for n in (1, 3, 8):
        s2 = Scorer(0)
        try:
            s2.set_policy(Scorer.POLICY_THROUGHPUT)
            s2.set_host_threads(n)
            res[n] = s2.score(p, scs, lin, ang)
            if n == 1:
                for k in (0, 5, 13, 36):
                    c1, b1 = s2.score(p, [scs[k]], lin, ang)
                    assert np.array_equal(c1[0], res[1][0][k]) and b1[0] == res[1][1][k], k
        finally:
            s2.close()
    for n in (3, 8):
        assert np.array_equal(res[n][0], res[1][0]) and np.array_equal(res[n][1], res[1][1]), n
    parity.compare(p, scs[13], lin, ang, res[8][0][13], res[8][1][13])


@pytest.mark.parametrize("n_v,n_w,n_scenes,world", [(100, 77, 1, 3), (64, 64, 5, 2), (96, 40, 2, 8)])
def test_prefix_sharing_on_row_slabs_is_bit_identical(n_v, n_w, n_scenes, world):
    """Row slabs (multi-GPU strong scaling of one batch) keep rollout prefix sharing: the sample launch walks the
    slab's own fork order (row tables rebuilt per slab), the path launches are the whole grid's.  Every slab's cost
    rows must equal the unshared full-grid run bit for bit, the merged winner must be the full grid's, and going
    back to the full grid afterwards must still work."""
    from social_force_window_planner_b200 import sharding
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C1"], n_v=n_v, n_w=n_w, n_peds=9, steps=40)
    scs = [S.make_scene(wl, i, hazards=(i == n_scenes - 1 and n_scenes > 1)) for i in range(n_scenes)]
    for k, sc in enumerate(scs):
        r = list(sc.robot)
        r[3] = float(np.float32(0.1 + 0.1 * k))
        r[5] = float(np.float32(-0.3 + 0.2 * k))
        r[10] = r[3]
        sc.robot = tuple(r)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    s2 = Scorer(0)
    try:
        s2.set_policy(Scorer.POLICY_THROUGHPUT)
        s2.set_prefix_sharing(0)
        full_costs, full_best = s2.score(p, scs, lin, ang)
        assert "share" not in s2.last_kernel
        s2.set_prefix_sharing(2)
        s2.upload(p, scs, lin, ang)
        recs = []
        for r in range(world):
            b, e = sharding.block_partition(n_v, world, r)
            s2.set_row_slab(b, e)
            s2.run()
            c, best = s2.download()
            assert "share" in s2.last_kernel, (r, s2.last_kernel)
            c3 = c.reshape(n_scenes, n_v, n_w)
            f3 = full_costs.reshape(n_scenes, n_v, n_w)
            assert np.array_equal(c3[:, b:e], f3[:, b:e]), r
            assert (c3[:, :b] == -2.0).all() and (c3[:, e:] == -2.0).all()
            recs.append(best)
        for k in range(n_scenes):
            assert sharding.merge_winners(np.array([rec[k] for rec in recs])) == full_best[k]
        s2.set_row_slab(0, n_v)
        s2.run()
        c, best = s2.download()
        assert "share" in s2.last_kernel and np.array_equal(c, full_costs) and np.array_equal(best, full_best)
    finally:
        s2.close()


def test_crowd_kernel_at_its_pedestrian_limit(scorer):
    """SFW_MAX_PEDS_CROWD = 2048 pedestrians (csrc/sfw_dev.h; the crowd of a trajectory lives in one block's shared
    memory): a 2048-pedestrian scene scores and meets the oracle, 2049 is refused with SFW_ERR_UNSUPPORTED — never
    a wrong answer or a crash."""
    from social_force_window_planner_b200.scorer import SfwError
    wl = dataclasses.replace(S.WORKLOADS["C2"], n_v=2, n_w=3, steps=6, n_peds=2048, ped_r_max=16.0, ped_sep=0.45,
                             map_w=800, map_h=800)
    sc = S.make_scene(wl, 0)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [sc], lin, ang)
    assert scorer.last_kernel == "sfw_score_crowd"
    st = parity.compare(p, sc, lin, ang, costs[0], best[0])
    assert st["valid"] >= 1, st
    wl2 = dataclasses.replace(wl, n_peds=2049)
    with pytest.raises(SfwError) as ei:
        scorer.score(p, [S.make_scene(wl2, 0)], lin, ang)
    assert ei.value.code == -3
    print(st)


def test_pipelined_batch_equals_staged_batch():
    """sfw_score_batch launches a big batch piece by piece while the rest of it is still being packed and copied
    (pieces of 8, 16, 32 ... MB of costmaps); upload + run + download stages everything first.  Same bits either
    way — with sharing on and off, and for every scene against a single-scene call."""
    from social_force_window_planner_b200.scorer import Scorer
    wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=40, n_w=40)   # 640 scenes x 41.6 KB = 26.6 MB of costmaps
    scs = S.make_scenes(wl, 640)
    for k, sc in enumerate(scs):
        r = list(sc.robot)
        r[3] = float(np.float32(0.05 + 0.6 * (k % 7) / 6.0))
        r[5] = float(np.float32(-0.4 + 0.8 * (k % 5) / 4.0))
        r[10] = r[3]
        sc.robot = tuple(r)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    s2 = Scorer(0)
    try:
        for sharing in (1, 0):
            s2.set_prefix_sharing(sharing)
            c_pipe, b_pipe = s2.score(p, scs, lin, ang)
            k_pipe = s2.last_kernel
            s2.upload(p, scs, lin, ang)
            s2.run()
            c_st, b_st = s2.download()
            assert s2.last_kernel == k_pipe and ("share" in k_pipe) == bool(sharing)
            assert np.array_equal(c_pipe, c_st) and np.array_equal(b_pipe, b_st), sharing
        for k in (0, 200, 201, 402, 639):  # first / last scenes of the pieces
            c1, b1 = s2.score(p, [scs[k]], lin, ang)
            if s2.last_kernel == k_pipe:
                assert np.array_equal(c1[0], c_pipe[k]) and b1[0] == b_pipe[k], k
        parity.compare(p, scs[402], lin, ang, c_pipe[402], b_pipe[402])
    finally:
        s2.close()
