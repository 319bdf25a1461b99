"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family and the first launches in order."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg, seq = collections.OrderedDict(), []
for row in csv.DictReader(lines):
    t = float(row['Metric Value'].replace(',', ''))
    t = {'ns': t / 1e3, 'us': t, 'usecond': t, 'ms': t * 1e3, 'msecond': t * 1e3, 'nsecond': t / 1e3}.get(row['Metric Unit'], t)
    name = row['Kernel Name']
    short = re.sub(r'void |\(anonymous namespace\)::', '', name)
    short = re.sub(r'\(.*', '', short)
    seq.append((short, t, row.get('Grid Size', ''), row.get('Block Size', '')))
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {len(seq)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100*v[1]/tot:5.1f}% n={v[0]:4d} avg={v[1]/v[0]:8.1f}  {k[:120]}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for s in seq[:n]:
    print(f"{s[1]:9.1f} {s[2]:>16} {s[0][:110]}")
