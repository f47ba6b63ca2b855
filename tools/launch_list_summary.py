#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/launch_list_summary.py launches.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    n = r["Kernel Name"].replace("void ", "").split("(")[0][-80:]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"])
tot = sum(a[1] for a in agg.values())
print("| launches | total us | share | avg us | kernel |\n|---|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| %d | %.1f | %.1f%% | %.1f | `%s` |" % (c, t / 1e3, 100 * t / tot, t / 1e3 / c, n))
print("\n%d launches, %.1f us of kernel time in total (cold-cache, serialised by ncu)" % (len(rows), tot / 1e3))
