"""Ablation probe (not the bench contract): device time of the C1 tick with parts of the scene removed,
to see how the phases of the step (pair forces / obstacle sums / footprint) add up or overlap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer

def timeit(tag, wl, scs, reps=8):
    p = wl.params(); lin, ang = wl.sample_arrays()
    st = torch.cuda.Stream(); s = Scorer(0, st.cuda_stream)
    with torch.cuda.stream(st):
        s.upload(p, scs, lin, ang)
        for _ in range(2): s.run()
        s.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record(st)
        for i in range(reps):
            s.run(); ev[i + 1].record(st)
        s.sync()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
        costs, best = s.download()
    n = len(scs) * wl.samples
    print(f"{tag:44s} {np.median(ts):9.3f} ms  {n / (np.median(ts) * 1e-3):.3e} traj/s  valid {(costs >= 0).mean():.3f}  {s.last_kernel}", flush=True)
    s.close()

wl = S.WORKLOADS["C1"]
timeit("C1 full", wl, [S.make_scene(wl, 0)])
timeit("C1 no obstacles (M=0)", wl, [S.make_scene(wl, 0, n_obstacles=0)])
timeit("C1 no pedestrians (P=0)", wl, [S.make_scene(wl, 0, n_peds=0)])
timeit("C1 point footprint (F=0)", wl, [S.make_scene(wl, 0, footprint=np.zeros((0, 2)))])
timeit("C1 no peds, no obstacles", wl, [S.make_scene(wl, 0, n_peds=0, n_obstacles=0)])
timeit("C1 no peds, no obstacles, point footprint", wl, [S.make_scene(wl, 0, n_peds=0, n_obstacles=0, footprint=np.zeros((0, 2)))])
timeit("C1 M=0, point footprint (pairs only)", wl, [S.make_scene(wl, 0, n_obstacles=0, footprint=np.zeros((0, 2)))])
for name, nsc in (("C4", 1), ("C3", 512), ("C0", 1)):
    w = S.WORKLOADS[name]
    timeit(f"{name} x{nsc}", w, S.make_scenes(w, nsc))
w = S.WORKLOADS["C2"]
timeit("C2", w, S.make_scenes(w, 1), reps=2)
