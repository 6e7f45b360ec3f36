"""Group scripts/ncu_attrib.py output by code region of sfw_kernels.cu (function the line belongs to)."""
import re, subprocess, sys, os
rep = sys.argv[1]
src = open("social_force_window_planner_b200/csrc/sfw_kernels.cu").read().splitlines()
# region boundaries by scanning for function starts
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:template.*)?\s*(?:__device__ __forceinline__|extern \"C\" __global__).*?\b(\w+)\(", l)
    if m:
        marks.append((i, m.group(1)))
    if "-- social force step" in l:
        marks.append((i, "step:sfm"))
    if "-- legality of the current pose" in l:
        marks.append((i, "step:rollout"))
    if "---- terminal costs" in l:
        marks.append((i, "terminal+argmin"))
    if "// obstacle force" == l.strip():
        marks.append((i, "step:ped-update"))
marks.sort()
def region(ln):
    r = "?"
    for i, n in marks:
        if i <= ln:
            r = n
    return r
out = subprocess.run([sys.executable, "scripts/ncu_attrib.py", rep, "sfw_score_small", "social_force_window_planner_b200/libsfw_b200.so", "0.0"], capture_output=True, text=True).stdout
agg = {}
for l in out.splitlines():
    m = re.match(r'(\S+):(\d+)\s+inst\s+([\d.]+)%\s+thr/inst\s+([\d.]+)\s+samples\s+([\d.]+)%', l)
    if not m:
        continue
    fn, ln, p, t, s = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
    key = region(ln) if fn == "sfw_kernels.cu" else "lib:" + fn
    a = agg.setdefault(key, [0, 0])
    a[0] += p; a[1] += s
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] > 0.2 or v[1] > 0.2:
        print(f"{k:28s} inst {v[0]:5.1f}%  samples {v[1]:5.1f}%")
