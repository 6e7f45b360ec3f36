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
