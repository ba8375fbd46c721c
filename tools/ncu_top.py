"""Top stall locations from `ncu -i X.ncu-rep --page source --csv` (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
k = ci["# Samples"]
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((float(r[k]), i, r[ci["Source"]].strip()))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
top = sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for v, i, s in top:
    print(f"{v:8.0f} {100*v/tot:5.1f}%  line {i:5d}  {s[:100]}")
