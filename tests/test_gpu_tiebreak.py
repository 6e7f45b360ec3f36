"""-m gpu: arg-min TIE-BREAKS on the device, both kernel families (reference src/sfw_planner.cpp:344,394-414).

Scenes with exact ties (tests/tie_cases.py): cost(v, +w) == cost(v, -w), duplicated linvel rows, every cost equal,
cost == 10000 with linvel == 0 vs > 0, no valid sample.  The winner INDEX must be the one the reference's own
findBestAction picked (tests/golden/tie_golden.npz, from oracle/_ref) — under AUTO, THROUGHPUT (thread per
trajectory: warp -> block -> last-block reduction), LATENCY (block per trajectory: fused last-block reduction or
the stand-alone arg-min kernel), and through row slabs merged on the host and on the device."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import parity
import tie_cases as T
from social_force_window_planner_b200 import sharding
from social_force_window_planner_b200.scorer import Scorer

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tie_golden.npz"))
POLICIES = {"auto": Scorer.POLICY_AUTO, "throughput": Scorer.POLICY_THROUGHPUT, "latency": Scorer.POLICY_LATENCY}


def _expect(name):
    valid, index, v, w = GOLD[name + "/best"]
    return int(valid), int(index), float(v), float(w)


def _check_winner(best, name):
    valid, index, v, w = _expect(name)
    assert int(best["valid"]) == valid, (name, best)
    if valid:
        assert (int(best["index"]), float(best["v"]), float(best["w"])) == (index, v, w), (name, best)
    else:
        assert float(best["v"]) == 0.0 and float(best["w"]) == 0.0


@pytest.mark.parametrize("policy", list(POLICIES))
@pytest.mark.parametrize("name", list(T.CASES))
def test_winner_index_matches_reference_on_exact_ties(name, policy):
    p, sc, lin, ang = T.CASES[name]()
    s = Scorer(0)
    try:
        s.set_policy(POLICIES[policy])
        costs, best = s.score(p, [sc], lin, ang)
        kernel = s.last_kernel
    finally:
        s.close()
    c2 = costs[0].reshape(len(lin), len(ang))
    # the ties the scene was built for must exist in the GPU's own cost vector, bit for bit
    for a, b in T.mirror_pairs(ang):
        assert np.array_equal(c2[:, a], c2[:, b]), (name, policy, kernel, "cost(v,+w) != cost(v,-w)")
    if name in ("duplicated_rows", "big_grid"):
        dup = [(i, j) for i in range(len(lin)) for j in range(i + 1, len(lin)) if lin[i] == lin[j]]
        assert dup and all(np.array_equal(c2[i], c2[j]) for i, j in dup)
    if name == "all_costs_equal":
        assert (costs[0][costs[0] != -2.0] == 0.0).all()
    if name.startswith("cost_10000"):
        assert (c2[-1][c2[-1] != -2.0] == 10000.0).all()
    parity.compare(p, sc, lin, ang, costs[0], best[0])
    _check_winner(best[0], name)
    print(name, policy, kernel, best[0])


@pytest.mark.parametrize("policy", ["throughput", "latency"])
@pytest.mark.parametrize("name", ["no_zero_w", "all_costs_equal", "duplicated_rows", "big_grid"])
def test_three_slab_merge_keeps_the_reference_winner(name, policy):
    """Row slabs (multi-GPU strong scaling of one scene) on one device: per-slab winners merged on the host with
    sharding.merge_winners."""
    p, sc, lin, ang = T.CASES[name]()
    s = Scorer(0)
    try:
        s.set_policy(POLICIES[policy])
        s.upload(p, [sc], lin, ang)
        recs = []
        for r in range(3):
            b, e = sharding.block_partition(len(lin), 3, r)
            s.set_row_slab(b, e)
            s.run()
            _, best = s.download()
            recs.append(best[0])
    finally:
        s.close()
    _check_winner(sharding.merge_winners(np.array(recs)), name)


def test_many_tied_scenes_in_one_batch():
    """The same tied scene 300 times in one launch (several scenes per wave, last-block reductions racing): every
    scene must report the reference's winner; repeated launches too."""
    p, sc, lin, ang = T.CASES["no_zero_w"]()
    s = Scorer(0)
    try:
        s.set_policy(Scorer.POLICY_THROUGHPUT)
        for _ in range(3):
            costs, best = s.score(p, [sc] * 300, lin, ang)
            for k in range(300):
                _check_winner(best[k], "no_zero_w")
            assert (costs == costs[0]).all()
    finally:
        s.close()
