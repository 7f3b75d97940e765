#!/usr/bin/env python3
"""Per basic-block view of one kernel of an .ncu-rep: contiguous SASS runs with the same execution count, their share of the
stall samples and samples per executed instruction; optionally the annotated listing.
usage: ncu_blocks.py report.ncu-rep kernel-regex [listing.txt]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "sass", "-k", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
unit = min(int(r[ie]) for r in data if int(r[ie] or 0) > 0)
blocks, cur = [], None
lst = open(sys.argv[3], "w") if len(sys.argv) > 3 else None
agg = collections.Counter()
for k, r in enumerate(data):
    ex = int(r[ie] or 0) / unit
    s = int(r[isamp] or 0)
    toks = r[1].split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    op = toks[0] if toks else "?"
    st = {hdr[i][6:]: int(r[i] or 0) for i in stall if int(r[i] or 0) > 0}
    for a, b in st.items():
        agg[a] += b
    if lst:
        lst.write("%4d %7.2f %5d  %-64s %s\n" % (k, ex, s, r[1].strip()[:64], sorted(st.items(), key=lambda kv: -kv[1])[:3]))
    if cur is None or abs(cur["ex"] - ex) > 1e-9:
        cur = {"ex": ex, "k0": k, "n": 0, "s": 0, "fp": 0, "st": collections.Counter()}
        blocks.append(cur)
    cur["n"] += 1
    cur["s"] += s
    cur["k1"] = k
    cur["st"].update(st)
    if op[0] == "D" and op[:4] in ("DADD", "DMUL", "DFMA", "DSET", "DMNM"):
        cur["fp"] += 1
tot = sum(b["s"] for b in blocks)
print("samples", tot, "unit", unit, {k: "%.1f%%" % (100 * v / tot) for k, v in agg.most_common(8)})
for b in blocks:
    if b["s"] > tot * 0.005:
        print("%4d-%4d ex %7.2f n %4d fp64 %4d  %5.1f%% of samples, %6.2f samples/exec-inst  %s" % (
            b["k0"], b["k1"], b["ex"], b["n"], b["fp"], 100 * b["s"] / tot, b["s"] / (b["n"] * b["ex"]),
            [(a, "%.0f%%" % (100 * c / b["s"])) for a, c in b["st"].most_common(3)]))
