#!/usr/bin/env python
"""Key counters of `ncu --set full` reports -> JSON (the per-kernel summaries committed under profiles/).

    python scripts/ncu_summary.py name=path.ncu-rep [name=path.ncu-rep ...] > profiles/rN_ncu_full_summary.json
Reads each report with `ncu -i <rep> --page raw --csv`; stall_* = warp-state samples relative to `selected`.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        rec = dict(zip(head, r))
        k = {"Kernel Name": rec.get("Kernel Name", "")}
        for m in KEEP:
            if m in rec:
                k[m] = f"{rec[m]} {units[head.index(m)]}".strip()
        stalls = {h[len(STALL):]: float(rec[h].replace(",", "")) for h in head
                  if h.startswith(STALL) and not h.endswith("_not_issued") and rec[h] not in ("", "n/a")}
        sel = stalls.get("selected", 0.0)
        if sel > 0:
            for name, v in sorted(stalls.items()):
                if v / sel >= 0.3:
                    k["stall_" + name] = round(v / sel, 3)
        out.append(k)
    return out


def main():
    res = {}
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        res[name] = summarize(path)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
