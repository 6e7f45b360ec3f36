"""End-to-end latency of ONE control tick at the reference's shipped configuration (5 x 9 samples, sim_time 1.0 s
-> 40 steps, 200x200 costmap) through sfw_score (host buffers in, winner + cost vector out), next to the
reference's own CPU path (oracle/_ref) on one core.  Not the bench contract."""
import os, sys, time, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
import oracle_lib as ol

for n_peds in (1, 5, 20, 40):
    wl = dataclasses.replace(S.WORKLOADS["C0"], steps=40, n_peds=n_peds)
    sc = S.make_scene(wl, 0)
    p = wl.params()
    lin, ang = S.reference_sample_arrays()
    if len(sys.argv) > 1:
        wl2 = dataclasses.replace(wl, n_v=int(sys.argv[1]), n_w=int(sys.argv[1]))
        lin, ang = wl2.sample_arrays()
    from social_force_window_planner_b200._abi import SceneArray
    sa = SceneArray([sc])  # marshalled once: the timed call is the C ABI call, not Python object building
    s = Scorer(0)
    for _ in range(20):
        s.score(p, sa, lin, ang)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        s.score(p, sa, lin, ang)
        ts.append(time.perf_counter() - t0)
    s_kernel = s.last_kernel
    s.close()
    cpu = "n/a"
    if ol.have_ref():
        ol.ref_score(p, sc, lin, ang)
        t0 = time.perf_counter()
        for _ in range(5):
            ol.ref_score(p, sc, lin, ang, want_best=False)
        cpu = f"{(time.perf_counter() - t0) / 5 * 1e3:.2f} ms"
    print(f"{len(lin)}x{len(ang)} samples {s_kernel}, 40 steps, {n_peds:2d} peds: sfw_score e2e median {np.median(ts) * 1e6:.0f} us  p99 {np.percentile(ts, 99) * 1e6:.0f} us"
          f"   reference CPU (1 core) {cpu}", flush=True)
