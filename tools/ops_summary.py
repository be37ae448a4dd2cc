"""Summarise a bench.py --ops-csv table by (kind, impl, shape)."""
import collections
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
tot = sum(float(r['ms']) for r in rows)
print('total ms', round(tot, 3), 'ops', len(rows))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    a = agg[(r['kind'], r['impl'], r['shape'])]
    a[0] += 1; a[1] += float(r['ms']); a[2] += float(r['gflop'])
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]:>2} {k[1]:>2} {k[2]:45s} n={a[0]:3d} ms={a[1]:7.3f} ({100*a[1]/tot:4.1f}%) per={a[1]/a[0]*1000:7.1f}us TF/s={a[2]/max(a[1],1e-9):7.1f}")
