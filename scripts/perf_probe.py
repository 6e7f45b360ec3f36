"""Quick device-time probe of sfw_run on a named workload (not the bench contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
name = sys.argv[1] if len(sys.argv) > 1 else "C1"
nsc = int(sys.argv[2]) if len(sys.argv) > 2 else None
wl = S.WORKLOADS[name]
t0 = time.time(); scs = S.make_scenes(wl, nsc); print("scene gen %.2fs" % (time.time() - t0))
p = wl.params(); lin, ang = wl.sample_arrays()
st = torch.cuda.Stream()
s = Scorer(0, st.cuda_stream)
with torch.cuda.stream(st):
    t0 = time.time(); s.upload(p, scs, lin, ang); s.sync(); print("upload %.4fs" % (time.time() - t0))
    for _ in range(3): s.run()
    s.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record(st)
    for i in range(10):
        s.run(); ev[i + 1].record(st)
    s.sync()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(10)]
    n = len(scs) * wl.samples
    print(name, "kernel ms", np.median(ts), "traj/s %.3e" % (n / (np.median(ts) * 1e-3)), s.last_kernel)
    costs, best = s.download()
    print("valid", (costs >= 0).sum(), "of", costs.size, "best", best[0])
    from social_force_window_planner_b200._abi import SceneArray
    sa = SceneArray(scs)  # marshalled once: the timed call is the C ABI call with host buffers
    s.score(p, sa, lin, ang, want_costs=True)
    t0 = time.time()
    for _ in range(5): s.score(p, sa, lin, ang, want_costs=True)
    print("e2e ms", (time.time() - t0) / 5 * 1e3)
    t0 = time.time()
    for _ in range(5): s.upload(p, sa, lin, ang); s.sync()
    print("  upload+sync ms", (time.time() - t0) / 5 * 1e3)
    s.run(); s.sync()
    t0 = time.time()
    for _ in range(5): s.download()
    print("  download ms", (time.time() - t0) / 5 * 1e3)
