#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, opcode mix weighted by
executed count, and executed instructions per CUDA source line range.  usage: ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def page(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    lines = []
    raw = page(rep, "--page", "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
            "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
    lines.append("| metric | value | unit |\n|---|---|---|")
    for k in keys:
        if k in m:
            lines.append("| %s | %s | %s |" % (k, m[k], u.get(k, "")))
    sass = page(rep, "--page", "source", "--print-source", "sass")
    h = sass[1]
    ie, src = h.index("Instructions Executed"), h.index("Source")
    data = [r for r in sass[2:] if len(r) > ie and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in data)
    c = Counter()
    for r in data:
        t = r[src].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[ie])
    lines.append("\nstatic SASS instructions: %d; executed warp instructions: %d\n" % (len(data), tot))
    lines.append("| opcode | % of executed |\n|---|---|")
    for op, n in c.most_common(22):
        lines.append("| %s | %.2f |" % (op, 100.0 * n / tot))
    fp64 = sum(c[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
    lines.append("\nFP64 arithmetic share of executed instructions: %.1f %%" % (100.0 * fp64 / tot))
    cu = page(rep, "--page", "source", "--print-source", "cuda,sass")
    # per source line
    try:
        h = None
        per = Counter()
        cur = None
        for r in cu:
            if len(r) >= 2 and r[0] == "File Name":
                cur = r[1].split("/")[-1]
            if "Instructions Executed" in r and "Source" in r and "Line No" in r[0:2] + r:
                h = r
                continue
            if h and cur and len(r) == len(h):
                ln, cnt = r[h.index("#")] if "#" in h else r[0], r[h.index("Instructions Executed")]
                if cnt.isdigit() and ln.isdigit():
                    per[(cur, int(ln))] += int(cnt)
        if per:
            lines.append("\n| file:line | % of executed |\n|---|---|")
            for (f, ln), n in per.most_common(25):
                lines.append("| %s:%d | %.2f |" % (f, ln, 100.0 * n / tot))
    except Exception as e:  # layout differences between ncu versions
        lines.append("(per-line table unavailable: %s)" % e)
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
