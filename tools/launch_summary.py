#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel.
    python tools/launch_summary.py gpurun_out/bench_launches.csv "<command that was profiled>" > profiles/<name>.md"""
import collections
import csv
import re
import sys


def main(path, what):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("=="))]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) != len(hdr) or r is hdr or r[iv] == "Metric Value":
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").strip()
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
        try:
            us = float(r[iv].replace(",", "")) * scale
        except ValueError:
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("# ncu launch list of `%s` (gpu__time_duration.sum, --clock-control none)\n" % what)
    print("Cold-cache, serialised per-launch times: compare SHARES.\n")
    print("| kernel | launches | total us | share | us/launch |\n|---|---|---|---|---|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f%% | %.2f |" % (name, n, us, 100 * us / tot, us / n))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")
