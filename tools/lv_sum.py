#!/usr/bin/env python3
"""lv_sum.py TAG...: one line per gpurun_out/lv_TAG.csv (tools/lv.sh): per-kernel durations of one CCpol gradient pass, ms"""
import csv, sys, os
for t in sys.argv[1:]:
    f = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "lv_%s.csv" % t)
    if not os.path.exists(f):
        print(t, "missing"); continue
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    st = [i for i, r in enumerate(rows) if "setup" in r[4]]
    rows = rows[st[0]:(st[1] if len(st) > 1 else len(rows))]
    print(t, " ".join("%s=%.3f" % (r[4].split("ccpol_")[1].split("_kernel")[0], int(r[-1]) / 1e6) for r in rows),
          "total=%.3f" % (sum(int(r[-1]) for r in rows) / 1e6))
