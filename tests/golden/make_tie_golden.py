"""Regenerates tests/golden/tie_golden.npz: the tie-break scenes of tests/tie_cases.py through the REFERENCE's own
findBestAction loop (oracle/_ref, src/sfw_planner.cpp:338-468 compiled unmodified): cost vector + the winner it
picked (index recovered from the marker it paints green, :435-441).

    make -C oracle ref && python tests/golden/make_tie_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import oracle_lib as ol  # noqa: E402
import tie_cases as T  # noqa: E402

out = {}
for name, mk in T.CASES.items():
    p, sc, lin, ang = mk()
    costs, best = ol.ref_score(p, sc, lin, ang)
    out[name + "/costs"] = costs
    out[name + "/best"] = np.array([best.valid, best.index, best.v, best.w], dtype=np.float64)
    n_w = len(ang)
    c2 = costs.reshape(len(lin), n_w)
    ties = sum(int((c2[:, a] == c2[:, b]).sum()) for a, b in T.mirror_pairs(ang))
    print(f"{name}: {len(costs)} samples, {(costs >= 0).sum()} valid, exact +-w ties {ties}/{len(lin) * (n_w // 2)}, "
          f"best valid={best.valid} index={best.index} v={best.v} w={best.w}")
np.savez_compressed(os.path.join(HERE, "tie_golden.npz"), **out)
