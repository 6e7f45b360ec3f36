"""Regenerates tests/golden/marker_golden.npz: the visualization MarkerArray the REFERENCE's own
findBestAction (oracle/_ref, compiled unmodified) leaves behind after one grid tick
(src/sfw_planner.cpp:345-417,435-441) on cases of tests/golden_cases.py.

    make -C oracle ref && python tests/golden/make_marker_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import golden_cases as G  # noqa: E402
import oracle_lib as ol  # noqa: E402

MARKER_CASES = ["ref_5x9_samples_40steps", "c0_hazards_40steps", "point_footprint_hazards", "c0_seed1"]
out = {}
for name in MARKER_CASES:
    p, sc, lin, ang = G.CASES[name]()
    ok, rgba, npts, xyz = ol.ref_markers(p, sc, lin, ang, max_points=64)
    out[name + "/ok"] = np.array([ok])
    out[name + "/rgba"] = rgba
    out[name + "/npts"] = npts
    out[name + "/xyz"] = xyz
    print(name, ok, "markers", len(npts), "red", int((rgba[:, 0] == 1).sum()), "blue", int((rgba[:, 2] == 1).sum()),
          "green", int((rgba[:, 1] == 1).sum()))
np.savez_compressed(os.path.join(HERE, "marker_golden.npz"), **out)
