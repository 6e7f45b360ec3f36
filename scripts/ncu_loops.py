"""Loop-level reading of an ncu report's SASS source page (no GPU needed):

    python scripts/ncu_loops.py <report.ncu-rep> [kernel substring] [min share %]

For every loop (a backward branch) that holds at least `min share` of the kernel's stall samples: its share of the
samples (warp time) and of the executed warp instructions, the trips, the instruction mix of the body, the
register-file operand reads of one trip (32-bit reads: a packed F32x2 operand counts 2, a broadcast / scalar 1,
immediates, constants and uniform registers 0) and the cycles one SM sub-partition spends per trip,

    cycles per trip = share of samples x elapsed cycles / (trips / (SMs x 4)),

next to the two lower bounds of the body: FP32-pipe cycles (a packed instruction holds the pipe for 2) and
register-file cycles (reads / 2: the register file of a sub-partition delivers two 32-bit operands per lane and
cycle — measured, scripts/dbg/mix_bench.cu).
"""
import collections, csv, re, subprocess, sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
SMS = 148


def ncu(page):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True,
                                          text=True).stdout.splitlines()))


raw = ncu("raw")
rh = raw[0]
launches = [dict(zip(rh, r)) for r in raw[2:]]
src = ncu("source")
# split the source page per launch
blocks, cur = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)


def reads_of(s):
    s = re.sub(r"^@!?U?P\d\s+", "", s)
    op = s.split()[0]
    args = s[len(op):].split(",")
    srcs = args if op.startswith(("ST", "BRA", "RED", "ATOM")) else args[1:]
    n = 0
    for t in srcs:
        t = t.strip()
        if not re.search(r"(?<![U\w])R\d+", t):
            continue
        w = 2 if ("F32x2" in t or ".64" in t) else 1
        if "[" in t:
            w = 1
        elif op.startswith("STS.128") or op.startswith("ST.E.128"):
            w = 4
        elif op.startswith(("STS.64", "DADD", "DMUL", "DFMA", "DSETP")):
            w = 2
        n += w
    return op.split(".")[0], n


# (a report with several results prints every result's page once per result: drop the repeats)
uniq = []
for b in blocks:
    if not uniq or uniq[-1]["name"] != b["name"] or uniq[-1]["rows"][:50] != b["rows"][:50]:
        uniq.append(b)
blocks = uniq
for li, b in enumerate(blocks):
    if want and want not in b["name"]:
        continue
    hdr = b["rows"][0]
    ix = {n: k for k, n in enumerate(hdr)}
    data = [r for r in b["rows"][1:] if len(r) >= len(hdr)]
    L = launches[li] if li < len(launches) else {}
    cyc = float(L.get("sm__cycles_elapsed.max", "0").replace(",", "") or 0)
    samp = [int(r[ix["# Samples"]] or 0) for r in data]
    ie = [float(r[ix["Instructions Executed"]] or 0) for r in data]
    addr = [int(r[ix["Address"]], 16) for r in data]
    a2i = {a: i for i, a in enumerate(addr)}
    tot_s, tot_i = sum(samp), sum(ie)
    print(f"==== {b['name']}   elapsed {cyc:.0f} cycles, {tot_i:.3e} warp instructions, {tot_s} samples")
    loops = []
    for i, r in enumerate(data):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", r[ix["Source"]])
        if m and int(m.group(1), 16) in a2i and a2i[int(m.group(1), 16)] <= i:
            loops.append((a2i[int(m.group(1), 16)], i))
    for a, e in sorted(loops):
        s = sum(samp[a:e + 1])
        if tot_s == 0 or 100.0 * s / tot_s < min_share or ie[e] == 0 or tot_s < 1000:
            continue  # (a back edge that never ran, or a launch too short to have samples)
        inner = [(x, y) for (x, y) in loops if x >= a and y <= e and (x, y) != (a, e)]
        mix, reads, fp_cyc = collections.Counter(), 0, 0
        for k in range(a, e + 1):
            op, n = reads_of(data[k][ix["Source"]])
            mix[op] += 1
            reads += n
            if op in ("FFMA2", "FMUL2", "FADD2"):
                fp_cyc += 2
            elif op in ("FFMA", "FMUL", "FADD"):
                fp_cyc += 1
        trips = ie[e]
        line = (f"loop [{a}..{e}] {e - a + 1:4d} instr  samples {100.0 * s / tot_s:5.1f} %  instructions "
                f"{100.0 * sum(ie[a:e + 1]) / tot_i:5.1f} %  trips {trips:.4g}")
        if not inner and cyc and trips:
            per = (s / tot_s) * cyc / (trips / (SMS * 4))
            line += (f"\n      leaf: {per:6.1f} cycles / trip / sub-partition;  bounds: FP32 pipe {fp_cyc}, register file "
                     f"{reads / 2:.0f} ({reads} reads), MUFU {8 * mix['MUFU']}, issue {e - a + 1}"
                     f"\n      mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common(12)))
        print(line)
