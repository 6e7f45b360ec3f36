"""Prefix sharing on batches whose scenes have different odometry (every scene walks its own fork order); not the bench contract."""
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
for name, nsc in (("C3", 256), ("C1", 6)):
    wl = S.WORKLOADS[name]; p = wl.params(); lin, ang = wl.sample_arrays()
    for hetero in (False, True):
        scs = S.make_scenes(wl, nsc)
        if hetero:
            for k, sc in enumerate(scs):
                r = list(sc.robot)
                r[3] = float(np.float32(0.05 + 0.6 * (k % 7) / 6.0)); r[5] = float(np.float32(-0.4 + 0.8 * (k % 5) / 4.0)); r[10] = r[3]
                sc.robot = tuple(r)
        with torch.cuda.stream(st):
            res = []
            for on in (False, True):
                s = Scorer(0, st.cuda_stream); s.set_prefix_sharing(on)
                s.upload(p, scs, lin, ang); s.sync()
                t = timeit(s); c, b = s.download(); res.append((t, c, s.shared_prefix_steps)); s.close()
            print(name, nsc, "hetero" if hetero else "same", "off %.3f on %.3f ms shared steps %.2f identical %s" % (res[0][0], res[1][0], res[1][2], np.array_equal(res[0][1], res[1][1])), flush=True)
