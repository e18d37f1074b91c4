#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_summary.py file.csv [skip-kernels-regex]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
ix = {n: i for i, n in enumerate(rows[h])}
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(rows[h]): continue
    k = r[ix['Kernel Name']].split('(')[0]
    if skip and skip.search(k): continue
    v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
    v = v / 1000 if u in ('us', 'usecond') else (v / 1e6 if u in ('ns', 'nsecond') else (v * 1000 if u in ('s', 'second') else v))
    a = agg.setdefault(k, [0, 0]); a[0] += v; a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"| {k[:70]} | {a[1]} | {a[0]:.3f} | {a[0] / a[1]:.4f} | {100 * a[0] / tot:.1f}% |")
