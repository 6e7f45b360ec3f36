"""Where the end-to-end time of ONE C1 tick goes beyond the kernels: python scripts/c1_e2e_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
from social_force_window_planner_b200._abi import SceneArray

for name in ("C1", "C0"):
    wl = S.WORKLOADS[name]
    sa = SceneArray(S.make_scenes(wl, 1))
    p = wl.params(); lin, ang = wl.sample_arrays()
    sc = Scorer(0)
    out = None
    for _ in range(5):
        out = sc.score(p, sa, lin, ang, out=out)
    K = 200
    t = np.zeros(5)
    for _ in range(K):
        t0 = time.perf_counter(); sc.upload(p, sa, lin, ang); t1 = time.perf_counter()
        sc.sync(); t2 = time.perf_counter()
        sc.run(); t3 = time.perf_counter()
        sc.sync(); t4 = time.perf_counter()
        sc.download(); t5 = time.perf_counter()
        t += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4])
    t0 = time.perf_counter()
    for _ in range(K):
        out = sc.score(p, sa, lin, ang, out=out)
    e2e = (time.perf_counter() - t0) / K
    t *= 1e6 / K
    print(f"{name}: upload call {t[0]:.1f} us (+ {t[1]:.1f} us until the H2D has landed), run call {t[2]:.1f} us, kernels {t[3]:.1f} us, "
          f"download {t[4]:.1f} us; one sfw_score_batch {e2e * 1e6:.1f} us")
    sc.close()
