"""Device time of one tick of every named workload (medians of CUDA-event times): python scripts/kernel_times.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
for name, n, reps in (("C1", 1, 30), ("C4", 1, 20), ("C3", 512, 15), ("C0", 1, 30), ("C2", 1, 3)):
    wl = S.WORKLOADS[name]
    st = torch.cuda.Stream()
    s = Scorer(0, st.cuda_stream)
    with torch.cuda.stream(st):
        s.upload(wl.params(), S.make_scenes(wl, n), *wl.sample_arrays()); s.sync()
        for _ in range(3 if name != "C2" else 1): s.run()
        s.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record(st)
        for i in range(reps):
            s.run(); ev[i + 1].record(st)
        s.sync()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
        print(f"{name} x{n}: median {np.median(ts):.4f} ms, min {min(ts):.4f} ms  ({s.last_kernel})")
    s.close()
