#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libboss_b200.so, the count of the instructions that prove which hardware
path it uses -- DMMA.8x8x4 (FP64 tensor core; tcgen05 / TMEM have no FP64 kind), UBLKCP (1-D TMA bulk copy),
SYNCS (mbarrier), and, as a negative control, UTCMMA / LDTM / UTMALDG (tcgen05 / TMEM / tensor-map TMA: expected 0).

    python tools/sass_counts.py > profiles/r02_sass_counts.txt        (needs only cuobjdump, no GPU)
"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "boss.jl_b200", "lib", "libboss_b200.so")
PATS = ["DMMA", "DFMA", "UBLKCP", "SYNCS", "MUFU", "UTCMMA|UTCHMMA|UTCQMMA", "LDTM|STTM", "UTMALDG", "STL|LDL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    arch = set()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {p: 0 for p in PATS}
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        ins = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not ins:
            continue
        op = ins.group(1)
        for p in PATS:
            if re.match(r"(?:%s)(?:\.|$)" % p, op):
                kernels[cur][p] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# %s   arch: %s" % (os.path.relpath(SO, ROOT), ", ".join(sorted(arch))))
    print("# columns: " + "  ".join(PATS))
    tot = {p: 0 for p in PATS}
    rows = []
    for (k, c), name in zip(kernels.items(), demangle):
        name = re.sub(r"\(.*", "", name).replace("boss::", "")
        rows.append((name, c))
        for p in PATS:
            tot[p] += c[p]
    for name, c in sorted(rows):
        if c["DMMA"] or c["UBLKCP"] or "kernel" in name:
            print("%-64s %s" % (name[:64], " ".join("%6d" % c[p] for p in PATS)))
    print("%-64s %s" % ("TOTAL (%d kernels)" % len(rows), " ".join("%6d" % tot[p] for p in PATS)))


if __name__ == "__main__":
    main()
