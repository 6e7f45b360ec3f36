"""A/B of programmatic dependent launch on the prefix-sharing launches: run twice, with and without
SFW_B200_NO_PDL=1 in the environment (read at sfw_create).  python scripts/pdl_probe.py"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from social_force_window_planner_b200 import scenes as S
    from social_force_window_planner_b200.scorer import Scorer
    for name, n in (("C1", 1), ("C4", 1), ("C3", 512)):
        wl = S.WORKLOADS[name]
        st = torch.cuda.Stream()
        s = Scorer(0, st.cuda_stream)
        with torch.cuda.stream(st):
            s.upload(wl.params(), S.make_scenes(wl, n), *wl.sample_arrays()); s.sync()
            for _ in range(5): s.run()
            s.sync()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
            ev[0].record(st)
            for i in range(30):
                s.run(); ev[i + 1].record(st)
            s.sync()
            ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(30)]
            print(f"{'no PDL' if os.environ.get('SFW_B200_NO_PDL') else 'PDL   '} {name} x{n}: median {np.median(ts):.4f} ms, min {min(ts):.4f} ms  ({s.last_kernel})")
        s.close()
else:
    for env in ({}, {"SFW_B200_NO_PDL": "1"}, {}, {"SFW_B200_NO_PDL": "1"}):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e)
