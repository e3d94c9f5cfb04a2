#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share.  Usage: python scripts/summarize_launches.py gpurun_out/launches.csv [steps]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)
        t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r[ui]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches over {steps} step(s), "
          f"{tot / steps:.1f} us of kernel time per step (ncu: cold-cache, serialised -> compare SHARES)")
    print(f"{'kernel':90s} {'launches/step':>13s} {'us/step':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:90]:90s} {n / steps:13.1f} {t / steps:10.1f} {100 * t / tot:6.1f}%")


if __name__ == "__main__":
    main()
