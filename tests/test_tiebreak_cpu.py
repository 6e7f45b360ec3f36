"""Arg-min tie-breaks (reference src/sfw_planner.cpp:344,394-414) on the CPU: the oracle's arg-min and cost vectors
against what the reference's own findBestAction produced on the exact-tie scenes of tests/tie_cases.py
(tests/golden/tie_golden.npz, from oracle/_ref), and the host-side slab merge used by the multi-GPU path."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import tie_cases as T
from social_force_window_planner_b200 import sharding

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tie_golden.npz"))


@pytest.mark.parametrize("name", list(T.CASES))
def test_oracle_reproduces_reference_ties_and_winner(name):
    p, sc, lin, ang = T.CASES[name]()
    costs, best, _ = ol.oracle_score(p, sc, lin, ang)
    gold = GOLD[name + "/costs"]
    assert np.array_equal(costs, gold), "oracle cost vector must equal the reference's bit for bit"
    valid, index, v, w = GOLD[name + "/best"]
    assert best.valid == int(valid)
    if valid:
        assert (best.index, best.v, best.w) == (int(index), v, w)
    c2 = gold.reshape(len(lin), len(ang))
    pairs = T.mirror_pairs(ang)
    assert pairs and all(np.array_equal(c2[:, a], c2[:, b]) for a, b in pairs), "the scene must hold exact +-w ties"


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", list(T.CASES))
def test_fixture_is_current(name):
    p, sc, lin, ang = T.CASES[name]()
    costs, best = ol.ref_score(p, sc, lin, ang)
    assert np.array_equal(costs, GOLD[name + "/costs"])
    assert [best.valid, best.index, best.v, best.w] == list(GOLD[name + "/best"])


@pytest.mark.parametrize("name", ["no_zero_w", "all_costs_equal", "duplicated_rows", "big_grid", "cost_10000_with_linvel"])
@pytest.mark.parametrize("world", [2, 3, 7])
def test_slab_winners_merge_to_the_reference_winner(name, world):
    """Row slabs of one scene (multi-GPU strong scaling): each rank's winner over its rows, merged with
    sharding.merge_winners, must be the reference's winner of the whole grid — ties across slabs included."""
    from social_force_window_planner_b200._abi import BEST_DTYPE
    p, sc, lin, ang = T.CASES[name]()
    gold = GOLD[name + "/costs"].reshape(len(lin), len(ang))
    recs = np.zeros(world, dtype=BEST_DTYPE)
    for r in range(world):
        b, e = sharding.block_partition(len(lin), world, r)
        slab = np.full_like(gold, -2.0)
        slab[b:e] = gold[b:e]
        sb = ol.oracle_argmin(slab.reshape(-1), lin, ang)
        recs[r] = (sb.valid, sb.index, sb.cost, 0.0, sb.v, sb.w)
    m = sharding.merge_winners(recs)
    valid, index, v, w = GOLD[name + "/best"]
    assert int(m["valid"]) == int(valid) and int(m["index"]) == int(index) and float(m["v"]) == v and float(m["w"]) == w
