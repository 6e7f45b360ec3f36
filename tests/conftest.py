import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the gpu-marked tests instead of failing in sfw_create
    (the product has no CPU fallback; `-m gpu` on a GPU box runs them all)."""
    if any(it.get_closest_marker("gpu") for it in items) and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
        for it in items:
            if it.get_closest_marker("gpu"):
                it.add_marker(skip)


@pytest.fixture(scope="session")
def scorer():
    from social_force_window_planner_b200.scorer import Scorer
    s = Scorer(0)
    yield s
    s.close()


@pytest.fixture(autouse=True)
def _parity_label(request):
    """Every parity.compare / check_samples call of a test is filed under the test's id."""
    import parity
    parity._LABEL[0] = request.node.nodeid
    yield
    parity._LABEL[0] = None


def pytest_sessionfinish(session, exitstatus):
    """Parity statistics of the run (near-discontinuity share, worst relative error on each side, branch
    resolutions) -> gpurun_out/parity_stats.json; the tracked copy is profiles/r2_parity_stats.json."""
    try:
        import parity
    except Exception:
        return
    if not parity.STATS:
        return
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    tot = dict(calls=len(parity.STATS))
    for k in ("n", "valid", "near", "base_ok", "branch_resolved", "branch_runs", "unresolved", "validity_flips_near"):
        tot[k] = int(sum(s[k] for s in parity.STATS))
    for k in ("max_rel_clear", "max_rel_near", "max_flips", "max_events_per_traj"):
        tot[k] = max(s[k] for s in parity.STATS)
    with open(os.path.join(out, "parity_stats.json"), "w") as f:
        json.dump(dict(rtol=parity.RTOL, margins=dict(goal=parity.GOAL_MARGIN, collision=parity.COLLISION_MARGIN,
                                                      theta=parity.THETA_MARGIN,
                                                      theta_min_weight=parity.THETA_MIN_WEIGHT),
                       exitstatus=int(exitstatus), totals=tot, calls=parity.STATS), f, indent=1)
