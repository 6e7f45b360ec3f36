"""-m gpu: CUDA scorer (through the C ABI) vs the CPU oracle on identical seeded scenes."""
import dataclasses

import numpy as np
import pytest

from social_force_window_planner_b200 import scenes as S

import parity

pytestmark = pytest.mark.gpu


def _run(scorer, wl, scene_index=0, **kw):
    sc = S.make_scene(wl, scene_index, **kw)
    p = wl.params()
    lin, ang = wl.sample_arrays()
    costs, best = scorer.score(p, [sc], lin, ang)
    return parity.compare(p, sc, lin, ang, costs[0], best[0])


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_c0_cpu_ref_config(scorer, seed):
    st = _run(scorer, S.WORKLOADS["C0"], seed)
    print(st)


@pytest.mark.parametrize("seed", [0, 1])
def test_c1_shape_reduced_grid(scorer, seed):
    wl = dataclasses.replace(S.WORKLOADS["C1"], n_v=24, n_w=24)
    st = _run(scorer, wl, seed)
    print(st)
