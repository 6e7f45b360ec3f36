"""Attribute an ncu SASS source page to CUDA source lines (no GPU needed).

    python scripts/ncu_attrib.py <report.ncu-rep> <kernel name> [lib.so] [min_pct]

Joins `ncu --page source --csv` (per-SASS-instruction counts, in address order) with
`nvdisasm -g` line markers of the same kernel in the in-tree library; prints the share of executed
warp instructions, average active threads and stall samples per source line."""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "social_force_window_planner_b200/libsfw_b200.so"
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
lines = None
for f in os.listdir(tmp):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if f".text.{kern}" in out:
        lines = out.splitlines()
        break
assert lines, "kernel not found"
# walk the kernel's section, collect (line, inlined-at chain) per instruction
insts = []
cur = None
on = False
for ln in lines:
    if ln.startswith("//--------------------- .text."):
        on = f".text.{kern} " in ln or ln.rstrip().endswith(f".text.{kern}") or f".text.{kern}\t" in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        insts.append((cur, ln.strip()))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
ix = {n: k for k, n in enumerate(hdr)}
sass = [r for r in rows[h + 1:] if len(r) >= len(hdr)]
print(f"# sass rows {len(sass)}, disasm insts {len(insts)}")
n = min(len(sass), len(insts))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot = sum(float(r[ix["Instructions Executed"]]) for r in sass)
tots = sum(float(r[ix["# Samples"]]) for r in sass)
for k in range(n):
    r = sass[k]
    key = insts[k][0][:2] if insts[k][0] else ("?", 0)
    a = agg[key]
    a[0] += float(r[ix["Instructions Executed"]])
    a[1] += float(r[ix["Thread Instructions Executed"]])
    a[2] += float(r[ix["# Samples"]])
print(f"# total warp insts {tot:.3e}, samples {tots:.0f}")
src = {}
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    pct = a[0] / tot * 100
    if pct < minpct:
        continue
    fn, l = key
    if fn not in src:
        try:
            src[fn] = open(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", fn)).read().splitlines()
        except Exception:
            src[fn] = []
    text = src[fn][l - 1].strip() if 0 < l <= len(src[fn]) else ""
    print(f"{fn}:{l:<5} inst {pct:5.1f}%  thr/inst {a[1] / max(a[0], 1):5.1f}  samples {a[2] / max(tots, 1) * 100:5.1f}%  {text[:90]}")
