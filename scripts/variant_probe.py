"""A/B of kernel-experiment builds (build/variants/libsfw_<name>.so, scripts/build_variants.sh): device time of one
tick of the named workloads per variant and a hash of the cost vector (bit-identity across variants).

    python scripts/variant_probe.py base new ...            # every workload
    SFW_PROBE_WL=C1,C2 python scripts/variant_probe.py ...   # a subset
"""
import hashlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("SFW_PROBE_CHILD"):
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from social_force_window_planner_b200 import scenes as S
    from social_force_window_planner_b200.scorer import Scorer
    want = os.environ.get("SFW_PROBE_WL", "C1,C4,C3,C0,C2").split(",")
    for name, n, reps in (("C1", 1, 30), ("C4", 1, 20), ("C3", 512, 15), ("C0", 1, 30), ("C2", 1, 3)):
        if name not in want:
            continue
        wl = S.WORKLOADS[name]
        st = torch.cuda.Stream()
        s = Scorer(0, st.cuda_stream)
        with torch.cuda.stream(st):
            s.upload(wl.params(), S.make_scenes(wl, n), *wl.sample_arrays()); s.sync()
            for _ in range(3 if name != "C2" else 1): s.run()
            s.sync()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
            ev[0].record(st)
            for i in range(reps):
                s.run(); ev[i + 1].record(st)
            s.sync()
            ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
            costs, best = s.download()
            h = hashlib.sha256(np.ascontiguousarray(costs).tobytes()).hexdigest()[:12]
            print(f"  {name} x{n}: median {np.median(ts):.4f} ms, min {min(ts):.4f} ms  costs {h} valid {(costs >= 0).mean():.3f} ({s.last_kernel})", flush=True)
        s.close()
    sys.exit(0)
for v in sys.argv[1:]:
    lib = os.path.join(ROOT, "build", "variants", f"libsfw_{v}.so")
    print(f"== {v}", flush=True)
    env = dict(os.environ, SFW_B200_LIB=lib, SFW_PROBE_CHILD="1")
    subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
