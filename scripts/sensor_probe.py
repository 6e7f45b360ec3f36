"""Device-time probe of sfw_laser_obstacles on a batch of scans (one per scene of BASELINE configs[3]) next to
the oracle's CPU restatement of SFMSensorInterface::laserCb on one host core.  Not the bench contract."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol, sensor_cases as SC
from social_force_window_planner_b200.scorer import Scorer

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scans = [SC.make_scan(100 + (k % 64), n_beams=720, n_people=10) for k in range(64)]
scans = [scans[k % 64] for k in range(n_scans)]
s = Scorer(0)
outs = s.laser_obstacles(scans)
t0 = time.perf_counter()
for _ in range(5):
    outs = s.laser_obstacles(scans)
dt = (time.perf_counter() - t0) / 5
beams = sum(len(sc["ranges"]) for sc in scans)
kept = sum(len(o) for o in outs)
print(f"GPU  e2e (host buffers in, points out): {n_scans} scans, {beams} beams, {kept} kept: {dt * 1e3:.3f} ms "
      f"= {beams / dt:.3e} beams/s  ({(4 * beams + 16 * kept + 160 * n_scans) / dt / 1e9:.2f} GB/s algorithmic)")
t0 = time.perf_counter()
ref = [ol.oracle_laser_obstacles(sc) for sc in scans[:256]]
dtc = (time.perf_counter() - t0) / 256 * n_scans
print(f"CPU  oracle laserCb, 1 core (256 scans timed, scaled): {dtc * 1e3:.1f} ms = {beams / dtc:.3e} beams/s")
assert all(np.array_equal(a.shape, b.shape) for a, b in zip(outs[:256], ref))
s.close()
