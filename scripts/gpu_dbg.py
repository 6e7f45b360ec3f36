import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer
wl = S.WORKLOADS["C0"]
sc = S.make_scene(wl, 0)
p = wl.params(); lin, ang = wl.sample_arrays()
s = Scorer(0)
costs, best = s.score(p, [sc], lin, ang)
print(costs[0][:10], best)
