#!/usr/bin/env python3
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> time and share per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        if d.get('Metric Unit', 'ns') in ('us', 'usecond'):
            v *= 1e3
        a = agg.setdefault(d['Kernel Name'][:80], [0, 0.0])
        a[0] += 1
        a[1] += v
tot = sum(a[1] for a in agg.values()) or 1
for k, a in agg.items():
    print(f"{a[1] / 1e3:12.1f} us  {a[0]:4d} launches  {100 * a[1] / tot:5.1f}%  {k}")
