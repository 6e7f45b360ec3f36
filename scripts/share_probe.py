"""Device time of a tick with and without rollout prefix sharing, with the library's own decision trace
(SFW_B200_TRACE_SHARING); not the bench contract."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
st = torch.cuda.Stream()
os.environ['SFW_B200_TRACE_SHARING'] = '1'
def timeit(s):
    for _ in range(2): s.run()
    s.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    ev[0].record(st)
    for i in range(6):
        s.run(); ev[i + 1].record(st)
    s.sync()
    return np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(6)])
for name, nsc in (("C1", 1), ("C1", 2), ("C1", 6), ("C3", 40), ("C3", 60), ("C3", 100), ("C3", 512), ("C4", 1), ("C2", 1)):
    wl = S.WORKLOADS[name]; scs = S.make_scenes(wl, nsc); p = wl.params(); lin, ang = wl.sample_arrays()
    with torch.cuda.stream(st):
        res = []
        for on in (False, True):
            s = Scorer(0, st.cuda_stream); s.set_prefix_sharing(on)
            t0 = time.time(); s.upload(p, scs, lin, ang); s.sync(); tu = time.time() - t0
            t0 = time.time(); s.upload(p, scs, lin, ang); s.sync(); tu = time.time() - t0
            t = timeit(s); c, b = s.download(); res.append((t, c, s.last_kernel, tu)); s.close()
        print(name, nsc, "off %.3f ms (%s, upload %.2f ms)  on %.3f ms (%s, upload %.2f ms) identical %s" % (
            res[0][0], res[0][2], res[0][3] * 1e3, res[1][0], res[1][2], res[1][3] * 1e3, np.array_equal(res[0][1], res[1][1])), flush=True)
