"""Where the end-to-end latency of one 5 x 9 tick goes: sfw_upload / sfw_run / sfw_download timed separately."""
import os, sys, time, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
from social_force_window_planner_b200._abi import SceneArray
wl = dataclasses.replace(S.WORKLOADS["C0"], steps=40, n_peds=20)
sc = S.make_scene(wl, 0); p = wl.params(); lin, ang = S.reference_sample_arrays(); sa = SceneArray([sc])
s = Scorer(0)
for _ in range(20): s.score(p, sa, lin, ang)
T = {"upload": [], "upload+sync": [], "run+sync": [], "download": [], "score": []}
for _ in range(200):
    t0 = time.perf_counter(); s.upload(p, sa, lin, ang); t1 = time.perf_counter(); s.sync(); t2 = time.perf_counter()
    s.run(); s.sync(); t3 = time.perf_counter(); s.download(); t4 = time.perf_counter()
    s.score(p, sa, lin, ang); t5 = time.perf_counter()
    T["upload"].append(t1 - t0); T["upload+sync"].append(t2 - t0); T["run+sync"].append(t3 - t2); T["download"].append(t4 - t3); T["score"].append(t5 - t4)
for k, v in T.items():
    print(f"{k:12s} median {np.median(v) * 1e6:7.1f} us")
s.close()
