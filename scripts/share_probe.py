"""Device time of a tick with and without rollout prefix sharing (not the bench contract)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
for name, nsc in (("C4", 1), ("C3", 512), ("C1", 1), ("C1", 6)):
    wl = S.WORKLOADS[name]; scs = S.make_scenes(wl, nsc); p = wl.params(); lin, ang = wl.sample_arrays()
    st = torch.cuda.Stream(); s = Scorer(0, st.cuda_stream)
    with torch.cuda.stream(st):
        for on in (True, False):
            s.set_prefix_sharing(on)
            s.upload(p, scs, lin, ang)
            for _ in range(2): s.run()
            s.sync()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            ev[0].record(st)
            for i in range(6):
                s.run(); ev[i + 1].record(st)
            s.sync()
            ts = np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(6)])
            print(f"{name} x{nsc} sharing {'on ' if on else 'off'}: {ts:8.3f} ms  {nsc * wl.samples / ts / 1e3:.3e} traj/s  {s.last_kernel}", flush=True)
    s.close()
