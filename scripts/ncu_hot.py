"""Top SASS instructions by stall samples from `ncu --page source --csv` output (first kernel section)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < 7 or not r[ci["# Samples"]].isdigit():
        if data:
            break
        continue
    data.append(r)
tot = sum(int(r[ci["# Samples"]]) for r in data)
stall = [h for h in hdr if h.startswith("stall_") and not h.endswith("_not_issued")]
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for c in stall:
        agg[c] = agg.get(c, 0) + int(r[ci[c]] or 0)
print("stall mix:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for idx, r in enumerate(data):
    n = int(r[ci["# Samples"]])
    if n > tot * thr:
        top = sorted([(int(r[ci[c]] or 0), c) for c in stall], reverse=True)[:2]
        print(f"{idx:5d} {n:6d} {100 * n / tot:5.1f}%  {r[ci['Source']].strip()[:64]:64s} {top}")
