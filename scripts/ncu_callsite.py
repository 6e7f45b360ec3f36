"""Attribute an ncu SASS source page to the OUTERMOST call site in the kernel's own .cu file.

    python scripts/ncu_callsite.py <report.ncu-rep> <mangled kernel name> <file.cu> [lib.so] [min_pct]

Like scripts/ncu_attrib.py, but follows nvdisasm's inline chain (`nvdisasm -gi`) up to the line of the
kernel source file that (transitively) inlined the instruction: tells how the executed warp
instructions and the stall samples split over pair forces / obstacle sums / footprint / bookkeeping.
Also splits each call site by pipe class (MUFU, FP32x2, FP64, LDS/STS, other)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern, cu = sys.argv[1], sys.argv[2], sys.argv[3]
lib = sys.argv[4] if len(sys.argv) > 4 else "social_force_window_planner_b200/libsfw_b200.so"
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    out = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if f".text.{kern}" in out:
        lines = out.splitlines()
        break
assert lines, "kernel not found"
insts = []
chain = []
on = False
fresh = True
for ln in lines:
    if ln.startswith("//--------------------- .text."):
        on = ln.split()[1] == f".text.{kern}"
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        insts.append((list(chain), ln.strip()))
        fresh = True
# a report with several kernels: NCU_FILTER="--launch-skip 1 --launch-count 1" picks the launch
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] +
                                      os.environ.get("NCU_FILTER", "").split(), capture_output=True,
                                      text=True).stdout.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
ix = {n: k for k, n in enumerate(hdr)}
sass = []
for r in rows[h + 1:]:
    if "Source" in r and "Address" in r:  # newer ncu prints the table once per view: keep the first
        break
    if len(r) >= len(hdr):
        sass.append(r)
assert len(sass) == len(insts), (len(sass), len(insts))


def klass(op):
    op = op.split()[0] if not op.startswith("@") else op.split()[1]
    if op.startswith("MUFU"):
        return "mufu"
    if op.startswith(("FFMA2", "FMUL2", "FADD2")):
        return "fp32x2"
    if op.startswith(("FFMA", "FMUL", "FADD", "FMNMX", "FSEL", "FSETP", "FSET", "FCHK")):
        return "fp32"
    if op.startswith(("DADD", "DMUL", "DFMA", "DSETP", "F2F", "I2F", "F2I", "D2", "DMNMX")):
        return "fp64/cvt"
    if op.startswith(("LDS", "STS", "LDG", "STG", "LD.", "ST.", "LDC")):
        return "mem"
    return "other"


tot = sum(float(r[ix["Instructions Executed"]]) for r in sass)
tots = sum(float(r[ix["# Samples"]]) for r in sass)
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
for (ch, text), r in zip(insts, sass):
    site = next((c for c in reversed(ch) if c[0] == os.path.basename(cu)), ch[-1] if ch else ("?", 0))
    op = re.sub(r"^/\*[0-9a-f]+\*/\s+", "", text)
    a = agg[site]
    n = float(r[ix["Instructions Executed"]])
    a[0] += n
    a[1] += float(r[ix["# Samples"]])
    a[2][klass(op)] += n
src = open(cu).read().splitlines()
print(f"# {kern}: {len(sass)} SASS instructions, {tot:.4e} warp instructions executed, {tots:.0f} samples")
for site, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    pct = a[0] / tot * 100
    if pct < minpct:
        continue
    text = src[site[1] - 1].strip() if 0 < site[1] <= len(src) else ""
    mix = " ".join(f"{k}={v / tot * 100:.1f}" for k, v in a[2].most_common(4))
    print(f"{site[0]}:{site[1]:<4} inst {pct:5.1f}%  samples {a[1] / tots * 100:5.1f}%  [{mix}]  {text[:70]}")
allmix = collections.Counter()
for a in agg.values():
    allmix.update(a[2])
print("# mix: " + " ".join(f"{k}={v / tot * 100:.1f}%" for k, v in allmix.most_common()))
