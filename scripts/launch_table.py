#!/usr/bin/env python3
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches / total / mean (us)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
take = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
agg = collections.OrderedDict()
seq = []
for r in rows[1 + skip: 1 + skip + take]:
    name = r[ki].split('(')[0].replace('void ', '').replace('kvx::', '')[:48]
    t = float(r[vi].replace(',', '')) / 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    seq.append((name, t))
tot = sum(a[1] for a in agg.values())
for k, (c, t) in agg.items():
    print(f"{k:50s} n={c:4d} total={t:9.1f} us  mean={t / c:8.1f} us  share={t / tot:6.1%}")
print(f"{'sum':50s} total={tot:9.1f} us")
if '-v' in sys.argv:
    for n, t in seq: print(f"  {n:48s} {t:8.1f}")
