"""-m gpu: the winner exchange fused into the scorer's epilogue (csrc/sfw_exchange.cu).  On one GPU the ring has
a single member (the kernel stores into its own gather buffer through the same code path); the multi-rank
check is scripts/exchange_check.py (torchrun, >= 2 GPUs; bench.py --gpus N uses the same calls)."""
import dataclasses

import numpy as np
import pytest

from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer, SfwError

pytestmark = pytest.mark.gpu


def test_single_rank_exchange_matches_download():
    wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=12, n_w=12)
    scs = S.make_scenes(wl, 5)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    sc = Scorer(0)
    try:
        h = sc.exchange_export(8)
        assert len(h) == 64
        sc.exchange_connect(0, 1, [h])
        sc.upload(p, scs, lin, ang)
        for tick in range(3):  # both epoch parities
            sc.run()
            sc.exchange_sync()
            got = sc.exchange_fetch()
            costs, best = sc.download()
            assert got.shape == (1, 5) and np.array_equal(got[0], best), tick
        # the crowd path exports from its arg-min kernel
        wl2 = dataclasses.replace(S.WORKLOADS["C2"], n_v=3, n_w=3, steps=8, n_peds=70, ped_r_max=5.0)
        sc.upload(wl2.params(), [S.make_scene(wl2, 0)], *wl2.sample_arrays())
        sc.run()
        assert sc.last_kernel == "sfw_score_crowd"
        sc.exchange_sync()
        got = sc.exchange_fetch()
        _, best = sc.download()
        assert np.array_equal(got[0], best)
        # more scenes than the exported buffer holds: refused, not overrun
        with pytest.raises(SfwError):
            sc.upload(p, S.make_scenes(wl, 9), lin, ang)
            sc.run()
    finally:
        sc.close()


def test_exchange_call_order_is_checked():
    sc = Scorer(0)
    try:
        with pytest.raises(SfwError):
            sc.exchange_connect(0, 1, [b"\0" * 64])
        with pytest.raises(SfwError):
            sc.exchange_sync()
    finally:
        sc.close()


# ---- several ranks on ONE device: contexts of this process connected with sfw_exchange_connect_local ----------
def _ring(world, max_scenes):
    scs = [Scorer(0) for _ in range(world)]
    for s in scs:
        s.exchange_export(max_scenes)
    Scorer.exchange_connect_local(scs)
    return scs


def test_scene_batch_exchange_three_ranks_with_ragged_counts():
    """Scene-batch sharding (configs[3]): ranks stage 3 / 2 / 4 scenes of one list; after the kernels every rank
    holds all 9 winners in global scene order, equal to what one context computes for the whole list — without a
    collective.  Two ticks (both epoch slots)."""
    wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=12, n_w=12)
    scenes = S.make_scenes(wl, 9)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    ring = _ring(3, 4)
    ref = Scorer(0)
    try:
        ref.set_policy(Scorer.POLICY_THROUGHPUT)
        _, want = ref.score(p, scenes, lin, ang)
        counts = [3, 2, 4]
        parts = [scenes[0:3], scenes[3:5], scenes[5:9]]
        for s, part in zip(ring, parts):
            s.set_policy(Scorer.POLICY_THROUGHPUT)
            s.exchange_expect(counts)
            s.upload(p, part, lin, ang)
        for tick in range(2):
            for s in ring:
                s.run()
            for s in ring:
                got = s.exchange_fetch()
                assert got.shape == (9,) and np.array_equal(got, want), tick
    finally:
        for s in ring + [ref]:
            s.close()


@pytest.mark.parametrize("policy", [Scorer.POLICY_THROUGHPUT, Scorer.POLICY_LATENCY])
@pytest.mark.parametrize("name", ["no_zero_w", "all_costs_equal", "duplicated_rows", "big_grid", "all_invalid"])
def test_row_slab_merge_on_the_device(name, policy):
    """Row slabs of one scene over 3 ranks (one of them may get an EMPTY slab): every rank's slab winner goes to
    every rank's gather buffer from the scorer's epilogue, the wait kernel merges them with the reference's
    tie-break order, and every rank ends up with the winner the reference's findBestAction picked for the whole
    grid — on scenes built to tie across slabs."""
    import os
    import tie_cases as T
    from social_force_window_planner_b200 import sharding
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "tie_golden.npz"))[name + "/best"]
    p, sc, lin, ang = T.CASES[name]()
    world = 3
    ring = _ring(world, 1)
    try:
        for r, s in enumerate(ring):
            s.set_policy(policy)
            s.upload(p, [sc], lin, ang)
            s.set_row_slab(*sharding.block_partition(len(lin), world, r))
        for tick in range(2):
            for s in ring:
                s.run()
            for s in ring:
                m = s.exchange_merge()[0]
                assert int(m["valid"]) == int(gold[0]), (name, tick, m)
                if gold[0]:
                    assert (int(m["index"]), float(m["v"]), float(m["w"])) == (int(gold[1]), gold[2], gold[3]), (name, m)
    finally:
        for s in ring:
            s.close()


def test_empty_slab_still_delivers_and_merge_matches_full_grid():
    """5 linvel rows over 8 ranks: three ranks own no row.  They still deliver (invalid) records, nobody hangs,
    and the merged winner equals the single-context winner."""
    import golden_cases as G
    from social_force_window_planner_b200 import sharding
    p, sc, lin, ang = G.CASES["ref_5x9_samples_40steps"]()
    ring = _ring(8, 1)
    ref = Scorer(0)
    try:
        ref.set_policy(Scorer.POLICY_THROUGHPUT)
        _, want = ref.score(p, [sc], lin, ang)
        for r, s in enumerate(ring):
            s.set_policy(Scorer.POLICY_THROUGHPUT)
            s.upload(p, [sc], lin, ang)
            s.set_row_slab(*sharding.block_partition(len(lin), 8, r))
            s.run()
        for s in ring:
            assert s.exchange_merge()[0] == want[0]
    finally:
        for s in ring + [ref]:
            s.close()


def test_missing_peer_is_an_error_not_a_hang():
    """Rank 1 skips a tick: rank 0's fetch gives up after the timeout and names the rank that did not deliver
    (SFW_ERR_STATE); once rank 1 catches up the ring works again."""
    import time
    wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=8, n_w=8)
    scenes = S.make_scenes(wl, 2)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    ring = _ring(2, 1)
    try:
        for s, scn in zip(ring, scenes):
            s.exchange_set_timeout(0.3)
            s.upload(p, [scn], lin, ang)
        ring[0].run()  # rank 1 does not run
        t0 = time.perf_counter()
        with pytest.raises(SfwError) as ei:
            ring[0].exchange_fetch()
        assert ei.value.code == -4 and "rank 1" in str(ei.value)
        assert 0.25 < time.perf_counter() - t0 < 5.0
        ring[1].run()  # late, but it delivers: the tick completes
        a, b = ring[0].exchange_fetch(), ring[1].exchange_fetch()
        assert np.array_equal(a, b) and a.shape == (2, 1)
        _, b0 = ring[0].download()
        _, b1 = ring[1].download()
        assert a[0, 0] == b0[0] and a[1, 0] == b1[0]
    finally:
        for s in ring:
            s.close()


def test_pipelined_batches_in_a_ring():
    """Two ranks, each scoring a batch big enough to be launched piece by piece from inside sfw_score_batch: the
    exchange bookkeeping of a run happens once, before the first piece, and every rank still ends up with all
    winners."""
    wl = dataclasses.replace(S.WORKLOADS["C3"], n_v=16, n_w=16)
    scenes = S.make_scenes(wl, 900)  # 450 per rank x 41.6 KB = 18.7 MB > the 16 MB piece threshold
    p = wl.params()
    lin, ang = wl.sample_arrays()
    ring = _ring(2, 450)
    ref = Scorer(0)
    try:
        ref.set_policy(Scorer.POLICY_THROUGHPUT)
        _, want = ref.score(p, scenes, lin, ang, want_costs=False)
        for tick in range(2):
            for r, s in enumerate(ring):
                s.set_policy(Scorer.POLICY_THROUGHPUT)
                _, best = s.score(p, scenes[450 * r:450 * (r + 1)], lin, ang, want_costs=False)
                assert np.array_equal(best, want[450 * r:450 * (r + 1)])
            for s in ring:
                got = s.exchange_fetch()
                assert np.array_equal(got.reshape(-1), want), tick
    finally:
        for s in ring + [ref]:
            s.close()
