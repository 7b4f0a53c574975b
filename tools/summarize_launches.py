"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.3f} ms {c:5d} launches {100*t/tot:5.1f}%  {k[:80]}")
print(f"{tot:10.3f} ms total")
