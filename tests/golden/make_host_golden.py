"""Regenerates tests/golden/host_golden.json: the reference's own findBestAction (oracle/_ref) on the
scenarios of tests/host_cases.py.  Build-container only (needs /root/reference)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import host_cases as H  # noqa: E402

out = {}
for name in H.CASES:
    ok, cmd, wp, running = H.run_reference(name)
    out[name] = {"ok": ok, "cmd": list(cmd), "wp_index": wp, "running": running}
    print(name, out[name])
json.dump(out, open(os.path.join(HERE, "host_golden.json"), "w"), indent=1)
