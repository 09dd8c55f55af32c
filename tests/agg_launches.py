"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg, tot, n = collections.OrderedDict(), 0.0, 0
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    val = float(row['Metric Value'].replace(',', ''))
    val *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1e-3)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += val; tot += val; n += 1
print(f"total {tot:.1f} us over {n} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={c:4d}  avg {t/c:8.1f} us  {k[:100]}")
