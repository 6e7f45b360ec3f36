"""C2-shaped crowd tick on a reduced sample grid (for ncu captures of sfw_score_crowd)."""
import os, sys, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wl = dataclasses.replace(S.WORKLOADS["C2"], n_v=n, n_w=n)
sc = S.make_scene(wl, 0)
p = wl.params(); lin, ang = wl.sample_arrays()
st = torch.cuda.Stream(); s = Scorer(0, st.cuda_stream)
with torch.cuda.stream(st):
    s.upload(p, [sc], lin, ang)
    s.run(); s.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record(st)
    for i in range(2):
        s.run(); ev[i + 1].record(st)
    s.sync()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(2)]
    costs, best = s.download()
print(f"C2 {n}x{n}: {np.median(ts):.2f} ms  {n * n / np.median(ts) * 1e3:.3e} traj/s  valid {(costs >= 0).mean():.3f} {s.last_kernel}")
s.close()
