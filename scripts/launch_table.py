#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
launches, total / mean duration, share of the summed kernel time."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, frames=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        name = re.sub(r"\(.*", "", r[ik])
        rows.append((name, v * scale))
    agg = OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f %% |" % (n, c, t, t / c, 100 * t / tot))
    print("| all | %d | %.1f | | |" % (len(rows), tot))


if __name__ == "__main__":
    main(sys.argv[1])
