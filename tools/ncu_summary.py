"""Markdown summary of one kernel from `ncu -i X.ncu-rep --page raw --csv` (first matching launch).

usage: python tools/ncu_summary.py raw.csv [kernel-substring] > section.md
"""
import csv, sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]
_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    row = next(r for r in rows[2:] if pat in r[ci["Kernel Name"]])
    print(f"## {row[ci['Kernel Name']]}\n")
    dram = 0.0
    for k in KEYS:
        if k not in ci:
            continue
        v, u = row[ci[k]], units[ci[k]]
        print(f"- {k}: {v} {u}")
        if k.startswith("dram__bytes_"):
            dram += float(v.replace(",", "")) * _SCALE.get(u, 1.0)
    print(f"- DRAM traffic per launch (read+write): {dram / 1e6:.1f} MB")


if __name__ == "__main__":
    main()
