"""profiles/traffic.json <- DRAM bytes per launch from an ncu summary (tools/ncu_summary.py output of one `ncu --set full` capture).
    python tools/update_traffic.py c3d_cips_fwd gpurun_out/r02ae_ncu_cips_summary.md profiles/r02ae_ncu_cips_summary.md "B=16, round 2, CTA-pair kernel" """
import json, os, re, sys
key, md, cite, note = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0.0
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    m = re.search(r"^- " + re.escape(name) + r": ([0-9.eE+-]+) (\w+)", open(md).read(), re.M)
    if not m:
        sys.exit(f"{name} not found in {md}")
    tot += float(m.group(1)) * UNIT[m.group(2)]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
d = json.load(open(path))
d[key]["dram_bytes_per_launch"] = tot
d[key]["source"] = f"{cite} (ncu --set full, one launch, {note})"
json.dump(d, open(path, "w"), indent=1)
print(key, tot)
