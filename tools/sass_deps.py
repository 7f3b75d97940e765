#!/usr/bin/env python3
"""sass_deps.py object.o kernel-substring: static look at how well ptxas interleaved independent FP64 chains in one kernel:
for every FP64 instruction the distance (in instructions) to the producer of its nearest FP64-produced source register;
prints the histogram share of distances 1, 2, 3, 4+ (distance 1 = the instruction waits a full FP64 latency)."""
import re, subprocess, sys, collections
obj, sub = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
on, ins = False, []
for l in out.splitlines():
    if "Function :" in l:
        on = sub in l
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", l)
    if on and m:
        ins.append(m.group(1))
last = {}   # register -> index of the FP64 instruction that wrote it
hist = collections.Counter()
nfp = 0
for i, t in enumerate(ins):
    t = re.sub(r"^@!?U?P\d\s+", "", t)
    op = t.split()[0]
    regs = re.findall(r"\bR(\d+)\b", t)
    if op.startswith(("DADD", "DMUL", "DFMA")):
        nfp += 1
        dst, srcs = int(regs[0]), [int(r) for r in regs[1:]]
        d = [i - last[r] for r in srcs if r in last]
        if d:
            hist[min(min(d), 4)] += 1
        last[dst] = i
    elif regs and not op.startswith(("ST", "BRA", "ISETP", "DSETP")):
        last.pop(int(regs[0]), None); last.pop(int(regs[0]) + 1, None)
print(sub, "FP64 instructions", nfp, " producer distance 1/2/3/4+:", " ".join("%.1f%%" % (100.0 * hist[k] / max(1, sum(hist.values()))) for k in (1, 2, 3, 4)))
