"""B200-native DWA + social-force trajectory sampler/scorer.

One hot path of robotics-upo/social_force_window_planner — the (v, w) sampling/scoring loop of
``SFWPlanner::findBestAction`` and everything it calls per sample — as hand-written sm_100a CUDA
behind a C ABI (``include/sfw_b200.h``).  This package is the thin Python host side: ctypes
bindings, the synthetic-scene generator and a mirror of the reference's planner interface.
There is no CPU fallback: importing :mod:`._lib` fails if ``libsfw_b200.so`` is not built.
"""
from . import _abi, scenes  # noqa: F401

__all__ = ["_abi", "scenes"]
