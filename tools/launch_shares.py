#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/launch_shares.py launches.csv > shares.txt
Two tables: the whole-batch launches (largest grid of each kernel = the device-resident steps of bench.py) and all launches."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
recs = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    n = 1
    for v in re.findall(r"\d+", r[gs]):
        n *= int(v)
    recs.append((re.sub(r"\(.*", "", r[kn]), n, float(r[mv]) / 1e3))
mx = collections.defaultdict(int)
for n, g, t in recs:
    mx[n] = max(mx[n], g)


def table(sel):
    agg = collections.OrderedDict()
    for n, g, t in recs:
        if sel(n, g):
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += t
    tot = sum(a[1] for a in agg.values())
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-70s launches=%3d total=%10.1f us share=%5.1f%% avg=%8.1f us" % (n[:70], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))


print("(1) whole-batch launches (largest grid of each kernel)")
table(lambda n, g: g == mx[n])
print("\n(2) every launch")
table(lambda n, g: True)
