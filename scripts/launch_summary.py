#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares, averages)."""
import csv, re, sys
path = sys.argv[1]
lines = open(path).readlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
agg, tot = {}, 0.0
for r in csv.DictReader(lines[start:]):
    try:
        t = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    u = r['Metric Unit']
    t = t / 1e3 if u == 'ns' else t * 1e3 if u == 'ms' else t
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f"| `{k[:80]}` | {n} | {t/1e3:.2f} | {100*t/tot:.1f}% | {t/n:.1f} |")
print(f"\ntotal {tot/1e3:.1f} ms over {sum(a[0] for a in agg.values())} launches")
