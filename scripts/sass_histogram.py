#!/usr/bin/env python
"""Instruction histogram of libparc_b200.so's sm_100a SASS, per kernel (`cuobjdump -sass`), for profiles/.

    python scripts/sass_histogram.py > profiles/r2_sass_histogram.txt
Columns: kernel (demangled, shortened), instructions, then the mnemonics that say which hardware paths the kernel
uses: packed fp32x2 (FMUL2 / FADD2 / FFMA2), shuffles, global / shared loads and stores, conversions (F2I / I2F), MUFU,
atomics / reductions, TMA / bulk copies (UBLKCP / UTMA*), tensor-core ops (HMMA / UTC*MMA) -- the last two are expected
to be absent or near-absent: these kernels are gathers and short serial chains, not GEMMs.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parc_b200", "libparc_b200.so")
GROUPS = [("FMUL2/FADD2/FFMA2", r"^(FMUL2|FADD2|FFMA2)"), ("FFMA/FMUL/FADD", r"^(FFMA|FMUL|FADD)(\.|$)"),
          ("SHFL", r"^SHFL"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"),
          ("F2I/I2F", r"^(F2I|I2F|F2F)"), ("MUFU", r"^MUFU"), ("ATOM/RED", r"^(ATOM|RED|ATOMS|ATOMG|REDG)"),
          ("BAR", r"^BAR"), ("TMA/bulk", r"^(UBLKCP|UTMA|SYNCS)"), ("tensor", r"^(HMMA|IMMA|UTC.*MMA|QGMMA|UTCQMMA)"),
          ("ACQBULK/PDL", r"^(ACQBULK|PREEXIT)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    kernels, cur = [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    print("# cuobjdump -sass parc_b200/libparc_b200.so (sm_100a); one row per kernel")
    print("# " + " | ".join(["kernel", "instructions"] + [g for g, _ in GROUPS]))
    total = collections.Counter()
    for name, c in sorted(zip(names, kernels), key=lambda x: -sum(x[1].values())):
        short = re.sub(r"\((bool|int)\)", "", name)
        short = re.sub(r"\(.*", "", short).replace("parc::", "").replace("void ", "")
        row = [short[:70], str(sum(c.values()))]
        for g, pat in GROUPS:
            n = sum(v for k, v in c.items() if re.match(pat, k))
            total[g] += n
            row.append(str(n))
        print(" | ".join(row))
    print("# totals: " + ", ".join(f"{g} {total[g]}" for g, _ in GROUPS))


if __name__ == "__main__":
    main()
