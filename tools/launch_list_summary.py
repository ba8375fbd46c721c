"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised launches: compare SHARES).
    python tools/launch_list_summary.py gpurun_out/r02h_launches.csv > profiles/r02h_launch_list_summary.md"""
import collections
import csv
import io
import sys

txt = open(sys.argv[1], errors="replace").read()
start = txt.find('"ID"')
rows = list(csv.reader(io.StringIO(txt[start:]))) if start >= 0 else []
hdr = rows[0] if rows else []
ci = {h: i for i, h in enumerate(hdr)}
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ci["Kernel Name"]]
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + v)
total = sum(t for _, t in tot.values()) or 1.0
print("| kernel | launches | total ns | share of window |\n|---|---|---|---|")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"| {name[:90]} | {n} | {t:.0f} | {100 * t / total:.1f}% |")
print(f"\nwindow: {sum(n for n, _ in tot.values())} launches, {total / 1e6:.3f} ms")
