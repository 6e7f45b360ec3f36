"""Regenerates tests/golden/sensor_golden.npz by running the REFERENCE's own SFMSensorInterface —
src/sensor_interface.cpp compiled unmodified by oracle/Makefile into oracle/_ref/libsfw_ref_sensor.so — on
the message sets of tests/sensor_cases.py (callback order odom, people, laser, people, odom).

    make -C oracle ref && python tests/golden/make_sensor_golden.py

Only runnable where /root/reference exists; the .npz is committed.  External to the reference and therefore
NOT pinned by these vectors: tf2 (one planar rigid transform) and lightsfm's Agent container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import oracle_lib as ol  # noqa: E402
import sensor_cases as SC  # noqa: E402

out = {}
for k, (name, mk) in enumerate(SC.CASES.items()):
    sc = mk()
    for ptf in (False, True):
        if ptf and sc.get("tf") is None:
            continue
        ppl = SC.people_records(sc, k, sc.get("tf") if ptf else None)
        agents, obs = ol.ref_sensor_run(sc, ppl, SC.ODOM, people_has_tf=ptf)
        key = name + ("/people_tf" if ptf else "")
        out[key + "/agents"] = agents
        out[key + "/obstacles"] = obs
        print(f"{key}: {len(sc['ranges'])} beams -> {len(obs)} obstacle points, {len(agents) - 1} people")
np.savez_compressed(os.path.join(HERE, "sensor_golden.npz"), **out)
