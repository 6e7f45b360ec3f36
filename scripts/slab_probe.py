"""Device time of one C4 tick restricted to a row slab (what one rank of N runs): python scripts/slab_probe.py [N ...]
(A/B of builds through SFW_B200_LIB.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
wl = S.WORKLOADS["C4"]
st = torch.cuda.Stream()
s = Scorer(0, st.cuda_stream)
with torch.cuda.stream(st):
    s.upload(wl.params(), S.make_scenes(wl, 1), *wl.sample_arrays()); s.sync()
    for n in [1] + [int(a) for a in sys.argv[1:] if int(a) != 1] if len(sys.argv) > 1 else [1, 2, 4, 8]:
        rows = wl.n_v // n
        s.set_row_slab(0, rows)
        for _ in range(3): s.run()
        s.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        ev[0].record(st)
        for i in range(20):
            s.run(); ev[i + 1].record(st)
        s.sync()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(20)]
        print(f"C4 slab 1/{n} ({rows} rows): median {np.median(ts):.4f} ms  -> strong-scaling efficiency {np.median(ts) and (1.0 / n) / (np.median(ts) / T1) if n > 1 else 1.0:.3f}" if n > 1 else f"C4 full grid: median {np.median(ts):.4f} ms")
        if n == 1:
            T1 = float(np.median(ts))
s.close()
