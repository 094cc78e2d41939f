#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove what the kernels use (B200_PROFILING.md): UTMALDG (TMA bulk tensor
load), SYNCS (mbarrier), REDUX (warp reduce), IDP (dp4a / dp2a), MUFU, ATOMS, LDL/STL (local-memory spills), plus the register
count. Reads the in-tree library with cuobjdump; no GPU needed.

    python tools/sass_markers.py > profiles/r2_sass_markers.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200", "libofb200.so")
MARKS = ["UTMALDG", "SYNCS", "REDUX", "IDP", "MUFU", "ATOMS", "ATOMG", "SHFL", "LDL", "STL", "UCGABAR", "DFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for l in res.split("\n"):
        m = re.search(r"Function (\S+):", l)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", l)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    cur = None
    for l in sass.split("\n"):
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MARKS:
                if op.startswith(k):
                    counts[cur][k] += 1
    try:
        demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.split("\n")
    except OSError:
        demangle = []
    if len(demangle) < len(counts):
        demangle = list(counts)
    print("# cuobjdump -sass / -res-usage of libofb200.so (sm_100a): instructions per kernel and marker mnemonics")
    print("# %-70s %6s %5s %7s  %s" % ("kernel", "instr", "regs", "smem", "markers"))
    for (name, c), dm in zip(counts.items(), demangle):
        short = re.sub(r"^void ", "", dm).replace("<unnamed>::", "").replace("(bool)", "").replace("(int)", "")
        m = re.match(r"([A-Za-z_0-9:]+(?:<[^()]*>)?)", short)          # name + template arguments, parameter list dropped
        short = m.group(1) if m else short
        r = regs.get(name, ("?", "?"))
        marks = " ".join("%s=%d" % (k, c[k]) for k in MARKS if c[k])
        print("%-72s %6d %5s %7s  %s" % (short[:72], c["_total"], r[0], r[1], marks))


if __name__ == "__main__":
    sys.exit(main())
