"""A few ticks of one named workload for ncu (no timing here): python scripts/prof_workload.py C1|C2|C3|C4 [rows_begin rows_end] [n_scenes]
C2 is profiled on a row slab of the FULL grid (the whole grid is 0.4 s per launch; ncu replays every launch ~40 times)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from social_force_window_planner_b200 import scenes as S
from social_force_window_planner_b200.scorer import Scorer

name = sys.argv[1] if len(sys.argv) > 1 else "C1"
wl = S.WORKLOADS[name]
slab = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None
n_scenes = int(sys.argv[4]) if len(sys.argv) > 4 else (64 if name == "C3" else 1)
sc = Scorer(0)
sc.upload(wl.params(), S.make_scenes(wl, n_scenes), *wl.sample_arrays())
if slab:
    sc.set_row_slab(*slab)
for _ in range(4):
    sc.run()
sc.sync()
print(name, sc.last_kernel, "launches", sc.kernel_launches)
sc.close()
