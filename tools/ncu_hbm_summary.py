#!/usr/bin/env python3
"""Summarise an .ncu-rep holding `ncu --set full` captures of the HBM-shaped kernels of the path (read here, no GPU needed):
one row per captured launch with duration, DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), achieved DRAM GB/s
against the measured copy bandwidth of MEASURED_PEAKS.json, and the pipes that compete with it.
usage: ncu_hbm_summary.py report.ncu-rep [out.md] [--alg name=bytes ...]   (algorithmic bytes per launch, optional)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def page(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(v, unit):
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)


def to_ms(v, unit):
    return float(v) * {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    outp = None
    alg = {}
    for a in sys.argv[2:]:
        if a.startswith("--alg"):
            continue
        if "=" in a:
            k, v = a.split("=")
            alg[k] = float(v)
        else:
            outp = a
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6500.0
    raw = page(rep, "--page", "raw")
    h, units = raw[0], dict(zip(raw[0], raw[1]))
    rows = [dict(zip(h, r)) for r in raw[2:]]
    out = ["| kernel | grid x block | regs | duration [ms] | DRAM read [MB] | DRAM write [MB] | DRAM GB/s | %% of %.0f GB/s | algorithmic MB | traffic / algorithmic | FP64 pipe %% | issue active %% | warps active %% | L2 hit %% |" % peak,
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for k in rows:
        name = k["Kernel Name"].split("(")[0].split("::")[-1]
        ms = to_ms(k["gpu__time_duration.sum"], units.get("gpu__time_duration.sum", "ms"))
        rd = to_bytes(k["dram__bytes_read.sum"], units.get("dram__bytes_read.sum", "byte"))
        wr = to_bytes(k["dram__bytes_write.sum"], units.get("dram__bytes_write.sum", "byte"))
        gbs = (rd + wr) / (ms * 1e-3) / 1e9
        ab = None
        for key, v in alg.items():
            if key in name:
                ab = v
        out.append("| %s | %s x %s | %s | %.4f | %.1f | %.1f | %.0f | %.1f | %s | %s | %s | %s | %s | %s |" % (
            name, k.get("launch__grid_size", ""), k.get("launch__block_size", ""), k.get("launch__registers_per_thread", ""), ms, rd / 1e6, wr / 1e6,
            gbs, 100 * gbs / peak, "%.1f" % (ab / 1e6) if ab else "", "%.2f" % ((rd + wr) / ab) if ab else "",
            k.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", ""),
            k.get("smsp__issue_active.avg.pct_of_peak_sustained_active", ""),
            k.get("sm__warps_active.avg.pct_of_peak_sustained_active", ""),
            k.get("lts__t_sector_hit_rate.pct", "")))
    text = "\n".join(out)
    print(text)
    if outp:
        open(outp, "w").write(text + "\n")


if __name__ == "__main__":
    main()
