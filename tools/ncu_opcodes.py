#!/usr/bin/env python3
"""Executed warp instructions per SASS opcode of one kernel of an .ncu-rep (source page).
usage: ncu_opcodes.py report.ncu-rep kernel-regex"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "sass", "-k", "regex:" + rx],
                     capture_output=True, text=True).stdout
per, samp, tot, hdr = collections.Counter(), collections.Counter(), 0, None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        src = r[1].strip()
        toks = src.split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        op = toks[0].rstrip(";") if toks else "?"
        c = int(r[hdr.index("Instructions Executed")] or 0)
        per[op] += c
        samp[op] += int(r[hdr.index("# Samples")] or 0)
        tot += c
print("executed warp instructions:", tot)
fam = collections.Counter()
for op, c in per.items():
    fam[op.split(".")[0]] += c
for op, c in fam.most_common(40):
    print("%5.1f%%  %-12s %d" % (100.0 * c / max(tot, 1), op, c))
print("--- by full opcode")
for op, c in per.most_common(50):
    print("%5.1f%%  %-28s %12d  samples %d" % (100.0 * c / max(tot, 1), op, c, samp[op]))
