"""Top stall locations from `ncu -i rep --page source --csv` (SASS view): address, samples, dominant stall, SASS."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    tot += n
    st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    out.append((n, r[ix["Address"]], r[ix["Source"]], st, r[ix["Instructions Executed"]]))
print("total samples", tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for n, a, s, st, ex in sorted(out, reverse=True)[:top]:
    print(f"{n:7d} {100.0 * n / max(tot, 1):5.1f}% {a[-5:]} exec={ex:>8s} {st[0][1][6:]:>12s}:{st[0][0]:<6d} {st[1][1][6:]:>10s}:{st[1][0]:<5d} {s[:90]}")
