#!/usr/bin/env python3
"""Summarise an .ncu-rep that holds one `ncu --set full` capture of each kernel of the CCpol PES-gradient
pipeline (read here, no GPU needed): per-kernel headline metrics, stall reasons, DRAM traffic and the SASS-level
FP64 rate (2*DFMA + DMUL + DADD thread instructions, from the source page) against the DFMA peak.
usage: ncu_pipeline_summary.py report.ncu-rep nbeads [out.md] [--mode strict|fast|analytic --capture NAME]
With --mode the per-bead counters (DRAM bytes, executed FP64 flop) are also written into profiles/ccpol_counters.json together
with the hash of the kernel sources in the tree (bench.py reads them from there and refuses a stale entry)."""
import json
import os
import re
import csv
import io
import subprocess
import sys
from collections import Counter

PEAK_FLOP_PER_CYCLE = 148 * 64 * 2  # 148 SMs x 64 FP64 lanes x 2 (fma)


def page(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = {}
    av = sys.argv[1:]
    for i, a in enumerate(av):
        if a.startswith("--") and i + 1 < len(av):
            opts[a[2:]] = av[i + 1]
    args = [a for a in args if a not in opts.values()]
    rep, nbeads = args[0], int(args[1])
    raw = page(rep, "--page", "raw")
    h = raw[0]
    kern = [dict(zip(h, r)) for r in raw[2:]]
    prefix = "agrad" if opts.get("mode") == "analytic" else "ccpol"
    kern = [k for k in kern if re.search(prefix + r"_(\w+?)_kernel", k["Kernel Name"])]
    names = [re.search(prefix + r"_(\w+?)_kernel", k["Kernel Name"]).group(1) for k in kern]
    rows = [
        ("duration [ms]", "gpu__time_duration.sum"),
        ("registers/thread", "launch__registers_per_thread"),
        ("grid", "launch__grid_size"),
        ("block", "launch__block_size"),
        ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("FP64 pipe active %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("warp instructions executed", "smsp__inst_executed.sum"),
        ("DRAM read [MB]", "dram__bytes_read.sum"),
        ("DRAM write [MB]", "dram__bytes_write.sum"),
        ("local loads (warp inst)", "sass__inst_executed_local_loads"),
        ("stall: wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("stall: long scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("stall: short scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("stall: no instruction", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        ("stall: math pipe throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("stall: barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("stall: not selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ]
    units = dict(zip(h, raw[1]))
    out = ["| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
    for label, key in rows:
        vals = []
        for k in kern:
            v = k.get(key, "")
            try:
                f = float(v)
                if key.startswith("dram__bytes"):
                    u = units.get(key, "")
                    f = f * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
                v = "%.4g" % f
            except ValueError:
                pass
            vals.append(v)
        out.append("| %s | %s |" % (label, " | ".join(vals)))
    # SASS-level FP64 flop per kernel from the source page
    tot_ms = sum(float(k["gpu__time_duration.sum"]) for k in kern)
    flops, fracs, dram = [], [], 0.0
    for k, nm in zip(kern, names):
        src = page(rep, "--page", "source", "--print-source", "sass", "-k", "regex:%s_%s_kernel" % (prefix, nm))
        hh = src[1]
        te, so = hh.index("Predicated-On Thread Instructions Executed"), hh.index("Source")
        c = Counter()
        for r in src[2:]:
            if len(r) > te and r[te].isdigit():
                t = r[so].split()
                op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                c[op] += int(r[te])
        fl = 2 * c["DFMA"] + c["DMUL"] + c["DADD"]
        cyc = float(k["sm__cycles_elapsed.avg"])
        flops.append(fl)
        fracs.append(fl / cyc / PEAK_FLOP_PER_CYCLE)
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram += float(k[key]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units.get(key, ""), 1.0)
    out.append("| SASS FP64 flop (2 DFMA + DMUL + DADD) | %s |" % " | ".join("%.4g" % f for f in flops))
    out.append("| ... as %% of the DFMA peak over the kernel | %s |" % " | ".join("%.1f" % (100 * f) for f in fracs))
    shares = ["%s %.1f %%" % (n, 100 * float(k["gpu__time_duration.sum"]) / tot_ms) for n, k in zip(names, kern)]
    cyc_tot = sum(float(k["sm__cycles_elapsed.avg"]) for k in kern)
    text = "\n".join(out)
    text += "\n\nPass of %d beads = %d energies: %.3f ms in total (ncu, serialised) -> %.3g bead-gradients/s; stage shares: %s\n" % (
        nbeads, nbeads * 36, tot_ms, nbeads / tot_ms * 1e3, ", ".join(shares))
    text += "\nSASS-level FP64 flop per bead-gradient: %.0f; time-weighted over the pass: %.1f %% of the DFMA peak (%d flop/cycle)\n" % (
        sum(flops) / nbeads, 100 * sum(flops) / cyc_tot / PEAK_FLOP_PER_CYCLE, PEAK_FLOP_PER_CYCLE)
    text += "\nDRAM traffic of the pass (read + write, all stages): %.3f GB = %.0f bytes per bead-gradient\n" % (dram / 1e9, dram / nbeads)
    print(text)
    if len(args) > 2:
        open(args[2], "w").write(text)
    if "mode" in opts:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        from bench import ccpol_source_hash

        path = os.path.join(root, "profiles", "ccpol_counters.json")
        try:
            d = json.load(open(path))
        except Exception:
            d = {}
        d[opts["mode"]] = {"dram_bytes_per_bead": dram / nbeads, "sass_flop_per_bead": sum(flops) / nbeads,
                           "pass_ms_ncu": tot_ms, "beads": nbeads, "capture": opts.get("capture", os.path.basename(rep)),
                           "source_sha256": ccpol_source_hash(opts["mode"]),
                           "how": "ncu --set full --clock-control none --import-source on, one pass; dram__bytes_read.sum + "
                                  "dram__bytes_write.sum and 2*DFMA+DMUL+DADD thread instructions (source page), all kernels of the pipeline"}
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
        print("wrote", path)


if __name__ == "__main__":
    main()
