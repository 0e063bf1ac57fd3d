"""Condense an `ncu --page raw --csv` export into a per-kernel table (markdown).
usage: python profiles/summarize_ncu.py gpurun_out/prof_raw.csv > profiles/rNN_ncu_summary.md"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram R"),
    ("dram__bytes_write.sum", "dram W"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
             and "not_issued" not in h]
    cols = [k for k, _ in KEYS if k in idx]
    print("| kernel | " + " | ".join(f"{n} [{units[idx[k]]}]" if units[idx[k]] else n for k, n in KEYS if k in idx) + " | top stalls |")
    print("|---|" + "---|" * (len(cols) + 1))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("dfb::", "")
        vals = []
        for k in cols:
            v = r[idx[k]]
            try:
                vals.append(f"{float(v):.4g}")
            except ValueError:
                vals.append(v)
        st = []
        for h in stall:
            try:
                st.append((float(r[idx[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        st.sort(reverse=True)
        print(f"| {name} | " + " | ".join(vals) + " | " + ", ".join(f"{n} {v:.2f}" for v, n in st[:3]) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
