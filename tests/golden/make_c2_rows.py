"""Regenerates tests/golden/c2_rows.npz: BASELINE.json configs[2] (128 x 128 samples, 128 steps, 500 pedestrians)
at FULL size — 8 whole linvel rows = 1024 trajectories — scored by the CPU oracle with its branch probe on
(oracle/sfw_oracle.h: per trajectory the cost and the decisions taken within parity.MARGINS of a discontinuity).

    python tests/golden/make_c2_rows.py            # ~4.8 s per trajectory per core: about 10 min on 8 cores

One oracle trajectory of this scene is 32 M pair-force evaluations, so the GPU box replays the committed values
instead of spending charged GPU-box minutes on them (tests/test_gpu_parity.py::test_full_size_c2_rows_vs_oracle);
trajectories that need the other branch of a near decision are re-run there.  The scene is rebuilt from
scenes.make_scene(WORKLOADS["C2"], 0) on both sides; its CRC is stored and checked.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import golden_cases as G  # noqa: E402
import oracle_lib as ol  # noqa: E402
import parity  # noqa: E402
from social_force_window_planner_b200 import scenes as S  # noqa: E402

ROWS = [0, 17, 40, 41, 64, 90, 111, 127]

wl = S.WORKLOADS["C2"]
sc = S.make_scene(wl, 0)
p = wl.params()
lin, ang = wl.sample_arrays()
out = {"rows": np.array(ROWS, dtype=np.int64), "crc": np.array([G.scene_crc(sc)], dtype=np.uint64),
       "margins": np.array(parity.MARGINS)}
costs, events, n_events = [], [], []
for r in ROWS:
    t = time.time()
    c, ev, n = ol.oracle_probe_grid(p, sc, lin, ang, parity.MARGINS, first=r * wl.n_w, count=wl.n_w,
                                    max_events=parity.MAX_EVENTS)
    costs.append(c)
    events.append(ev)
    n_events.append(n)
    print(f"row {r}: {time.time() - t:.0f} s, valid {(c >= 0).sum()}/{len(c)}, near {(n > 0).sum()}, "
          f"max decisions {n.max()}, cost range {c[c >= 0].min() if (c >= 0).any() else -1:.3f} .. {c.max():.3f}",
          flush=True)
out["costs"] = np.stack(costs)
out["events"] = np.stack(events)
out["n_events"] = np.stack(n_events)
np.savez_compressed(os.path.join(HERE, "c2_rows.npz"), **out)
