#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, mean and last duration (us)."""
import collections
import csv
import sys


def summarize(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        d.setdefault(r[ki].split("(")[0][:44], []).append(v)
    return d


if __name__ == "__main__":
    for path in sys.argv[1:]:
        print(path)
        d = summarize(path)
        tot = sum(sum(v) for v in d.values())
        for k, v in d.items():
            print("  %-46s n=%4d mean=%8.1f us last=%8.1f  share=%5.1f%%" % (k, len(v), sum(v) / len(v), v[-1], 100 * sum(v) / tot))
