"""Regenerates tests/golden/node_golden.json: what the REFERENCE's whole plugin — SFWPlannerNode with its own
SFWPlanner, SFMSensorInterface, CostmapModel and Trajectory, every source compiled unmodified into
oracle/_ref/libsfw_ref_node.so — returns from computeVelocityCommands on the scenarios of tests/node_cases.py.

    make -C oracle ref && python tests/golden/make_node_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import node_cases as N  # noqa: E402
import oracle_lib as ol  # noqa: E402

out = {}
for name in N.NAMES:
    cmd, status, left, reached = ol.ref_node_run(**N.make(name))
    out[name] = {"cmd": cmd.tolist(), "status": status.tolist(), "plan_left": left.tolist(), "goal_reached": reached.tolist()}
    print(name, out[name])
json.dump(out, open(os.path.join(HERE, "node_golden.json"), "w"), indent=1)
