"""How the C1 tick scales with the number of warps per scheduler: the same scene with 1/4 .. 4/4 of the
linvel rows (each row = 256 samples = 8 warps).  Not the bench contract."""
import os, sys, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
wl = S.WORKLOADS["C1"]
sc = S.make_scene(wl, 0)
p = wl.params(); lin, ang = wl.sample_arrays()
st = torch.cuda.Stream(); s = Scorer(0, st.cuda_stream)
with torch.cuda.stream(st):
    for rows in (18, 37, 74, 111, 148, 185, 222, 256, 259 if False else 256):
        ri = np.unique(np.linspace(0, wl.n_v - 1, rows).round().astype(int))
        s.upload(p, [sc], lin[ri], ang)
        for _ in range(2): s.run()
        s.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        ev[0].record(st)
        for i in range(6):
            s.run(); ev[i + 1].record(st)
        s.sync()
        ts = np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(6)])
        warps = len(ri) * wl.n_w / 32
        print(f"rows {len(ri):4d}  warps {warps:6.0f}  warps/scheduler {warps / 592:5.2f}  {ts:7.3f} ms  {len(ri) * wl.n_w / ts / 1e3:.3e} traj/s  {s.last_kernel}", flush=True)
s.close()
