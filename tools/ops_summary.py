"""Summarise a bench.py --ops-csv table by (kernel, shape); second table: backbone / lifter sections."""
import collections
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tot = sum(float(r['ms']) for r in rows)
print('total ms', round(tot, 3), 'ops', len(rows))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in rows:
    a = agg[(r['kernel'].split('[')[0], r['shape'])]
    a[0] += 1; a[1] += float(r['ms']); a[2] += float(r['gflop']); a[3] += float(r['mbytes'])
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]:32s} {k[1]:38s} n={a[0]:3d} ms={a[1]:7.3f} ({100*a[1]/tot:4.1f}%) per={a[1]/a[0]*1000:7.1f}us TF/s={a[2]/max(a[1],1e-9):7.1f} GB/s={a[3]/max(a[1],1e-9):7.1f}")
sec = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    t = r['tag']
    if t.startswith('backbone.'):
        parts = t.split('.')
        key = 'backbone.' + parts[1] + ('.branches.' + parts[4] if len(parts) > 4 and parts[3] == 'branches' else ('.fuse' if 'fuse_layers' in t else ''))
        if parts[1].startswith('stage'):
            key = 'backbone.stage*' + key.split(parts[1], 1)[1]
    elif t.startswith('volume_net.'):
        parts = t.split('.')
        key = 'lifter.' + parts[1]
    else:
        key = t
    s = sec[key]
    s[0] += 1; s[1] += float(r['ms']); s[2] += float(r['gflop'])
print()
for k, s in sorted(sec.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} n={s[0]:3d} ms={s[1]:7.3f} ({100*s[1]/tot:4.1f}%) TF/s={s[2]/max(s[1],1e-9):7.1f}")
