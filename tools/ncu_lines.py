#!/usr/bin/env python3
"""Executed warp instructions per CUDA source line of one kernel of an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep kernel-regex [top N]"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "cuda,sass", "-k", "regex:" + rx],
                     capture_output=True, text=True).stdout
per, tot, cur, hdr = collections.Counter(), 0, "", None
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        c = r[hdr.index("Instructions Executed")]
        if c.isdigit():
            per[(cur, int(r[0]), r[1].strip()[:110])] += int(c)
            tot += int(c)
print("executed warp instructions:", tot)
for (f, ln, src), c in per.most_common(top):
    print("%5.1f%%  %s:%d  %s" % (100.0 * c / max(tot, 1), f, ln, src))
