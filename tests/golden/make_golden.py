"""Regenerates tests/golden/ref_golden.npz by running the REFERENCE's own sources — compiled unmodified
by oracle/Makefile into oracle/_ref/libsfw_ref.so — on the cases of tests/golden_cases.py.

    make -C oracle ref && python tests/golden/make_golden.py

Only runnable where /root/reference exists (the build container); the .npz is committed so the GPU box and
CI replay it without the reference.  lightsfm is not vendored by the reference: the golden values pin
everything EXCEPT lightsfm's internals (restated in oracle/stubs/lightsfm).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import golden_cases as G  # noqa: E402
import oracle_lib as ol  # noqa: E402
from social_force_window_planner_b200._abi import SceneArray  # noqa: E402

out = {}
for name, mk in G.CASES.items():
    p, sc, lin, ang = mk()
    costs, best = ol.ref_score(p, sc, lin, ang)
    out[name + "/costs"] = costs
    out[name + "/best"] = np.array([best.valid, best.index, best.v, best.w], dtype=np.float64)
    out[name + "/crc"] = np.array([G.scene_crc(sc)], dtype=np.uint64)
    print(f"{name}: {len(costs)} samples, {(costs >= 0).sum()} valid, best={best.valid}/{best.index}")

p, sc, lin, ang = G.CASES["c0_hazards_40steps"]()
sa = SceneArray([sc])
out["footprint/poly"] = np.array([ol.ref().sfw_ref_footprint_cost(sa.ptr(0), *pose) for pose in G.FOOTPRINT_POSES])
p, sc, lin, ang = G.CASES["point_footprint_hazards"]()
sa = SceneArray([sc])
out["footprint/point"] = np.array([ol.ref().sfw_ref_footprint_cost(sa.ptr(0), *pose) for pose in G.FOOTPRINT_POSES])
print("footprint poly ", out["footprint/poly"])
print("footprint point", out["footprint/point"])
np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
