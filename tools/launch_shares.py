#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: total time, share, launches.
   python tools/launch_shares.py gpurun_out/launches_X.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
    a = agg.setdefault(r[ki].split('(')[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'total ms':>10} {'share':>6} {'n':>5}  kernel   (per-launch times are cold-cache and serialised: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / 1000:10.3f} {100 * t / tot:5.1f}% {n:5d}  {k}")
