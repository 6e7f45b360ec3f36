"""Where a batched tick's end-to-end time goes (BASELINE configs[3], one GPU's share): pack + H2D (sfw_upload),
kernels (sfw_run), D2H + copy-out (sfw_download), for 1 .. 16 host workers.

    python scripts/c3_e2e_probe.py [n_scenes]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
from social_force_window_planner_b200._abi import SceneArray

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
wl = S.WORKLOADS["C3"]
sa = SceneArray(S.make_scenes(wl, n))
p = wl.params(); lin, ang = wl.sample_arrays()
for threads in (1, 2, 4, 8, 16):
    sc = Scorer(0)
    sc.set_host_threads(threads)
    for _ in range(3):
        sc.score(p, sa, lin, ang)
    K = 10
    t = np.zeros(4)
    for _ in range(K):
        t0 = time.perf_counter(); sc.upload(p, sa, lin, ang); t1 = time.perf_counter()
        sc.sync(); t2 = time.perf_counter()
        sc.run(); sc.sync(); t3 = time.perf_counter()
        sc.download(); t4 = time.perf_counter()
        t += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3])
    out = sc.score(p, sa, lin, ang)
    t0 = time.perf_counter()
    for _ in range(K):
        out = sc.score(p, sa, lin, ang, out=out)  # output buffers reused, as a C/C++ caller's would be
    e2e = (time.perf_counter() - t0) / K
    t *= 1e3 / K
    print(f"{n} scenes, {threads:2d} host workers: upload call {t[0]:.2f} ms (+ {t[1]:.2f} ms until the H2D has landed), "
          f"kernels {t[2]:.2f} ms, download {t[3]:.2f} ms; one sfw_score_batch {e2e * 1e3:.2f} ms = {e2e * 1e3 / t[2]:.3f} x kernels")
    sc.close()
