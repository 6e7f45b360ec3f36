"""Which kernel family serves one tick of the reference's 5 x 9 sample set faster, by crowd size: end-to-end
sfw_score latency under SFW_POLICY_THROUGHPUT (thread per trajectory) and SFW_POLICY_LATENCY (block per trajectory).
Feeds the crossover rule in make_plan (csrc/sfw_abi.cu).  Not the bench contract."""
import os, sys, time, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
from social_force_window_planner_b200._abi import SceneArray

grids = [None] + [int(a) for a in sys.argv[1:]]
for g in grids:
    for n_peds in (0, 1, 2, 3, 4, 5, 8):
        wl = dataclasses.replace(S.WORKLOADS["C0"], steps=40, n_peds=n_peds)
        sc = S.make_scene(wl, 0)
        p = wl.params()
        lin, ang = S.reference_sample_arrays() if g is None else dataclasses.replace(wl, n_v=g, n_w=g).sample_arrays()
        sa = SceneArray([sc])
        out = []
        for pol in (Scorer.POLICY_THROUGHPUT, Scorer.POLICY_LATENCY, Scorer.POLICY_AUTO):
            s = Scorer(0)
            s.set_policy(pol)
            for _ in range(20):
                s.score(p, sa, lin, ang)
            ts = []
            for _ in range(200):
                t0 = time.perf_counter()
                s.score(p, sa, lin, ang)
                ts.append(time.perf_counter() - t0)
            out.append("%s %4.0f us" % (s.last_kernel, np.median(ts) * 1e6))
            s.close()
        print(f"{len(lin)}x{len(ang)} samples, {n_peds} peds: throughput {out[0]} | latency {out[1]} | auto {out[2]}")
