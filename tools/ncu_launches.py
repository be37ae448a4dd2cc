"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total time, share.
   python tools/ncu_launches.py gpurun_out/launches.csv [launches_per_step] > profiles/rN_launches_summary.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
ix = {k: i for i, k in enumerate(rows[h])}
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[h + 1:]:
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    u = r[ix["Metric Unit"]]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
n = sum(a[0] for a in agg.values())
print(f"# ncu launch list summary ({sys.argv[1]})\n")
print(f"{n} launches, {tot / 1e3:.2f} ms total device time (cold-cache, serialised by ncu: compare SHARES, not absolutes)\n")
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % |")
