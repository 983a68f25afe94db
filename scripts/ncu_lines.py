#!/usr/bin/env python
"""Rank CUDA source lines of an `ncu --page source --csv --print-source cuda,sass` dump by warp
instructions executed and by stall samples (first launch in the report only).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > p.csv; python scripts/ncu_lines.py p.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, fname, hdr, seen = [], None, None, set()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        hdr = None
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r if fname not in seen else None
        seen.add(fname)
    elif hdr is not None and len(r) >= len(hdr) - 2 and r[0] not in ("", "Line No"):
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            out.append((float(r[ie] or 0), float(r[ss] or 0), fname, int(r[0]), r[1]))
        except ValueError:
            pass
ti, ts = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"total warp-instructions {ti:.0f}, stall samples {ts:.0f}")
for o in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{o[0]:10.0f} {o[0] / max(ti, 1):6.1%} inst | {o[1] / max(ts, 1):6.1%} stall | {o[2]}:{o[3]:<4d} {o[4].strip()[:95]}")
