"""Condense an `ncu --page raw --csv` export into a per-kernel table (markdown).
usage: python profiles/summarize_ncu.py gpurun_out/prof_raw.csv [--traffic-json out.json --key MODEL:BxS --source TEXT]
       > profiles/rNN_ncu_summary.md
--traffic-json additionally writes {key: {bench kernel name: mean DRAM bytes per launch}} (dram__bytes_read.sum +
dram__bytes_write.sum), which bench.py reports as `roofline.traffic`."""
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


def bench_name(kernel: str, grid: str) -> str:
    """ncu kernel name -> the name the library's event profiler / bench.py uses."""
    k = kernel.split("<")[0]
    if k == "k_conv_in":
        return "k_conv_in[df_conv0]" if "<2>" in kernel else "k_conv_in[erb_conv0]"
    if k == "k_gemm_bf16x3":
        return "k_gemm_bf16x3[gru_proj]"
    if k == "k_gru_tc":
        return "k_gru_tc512" if "512" in kernel else "k_gru_tc"
    return k


def traffic(path, out_json, key, source):
    import json
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("dfb::", "")
        try:
            b = float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]])
        except (ValueError, KeyError):
            continue
        unit_r = rows[1][idx["dram__bytes_read.sum"]].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit_r, 1.0)
        n = bench_name(name, "")
        s_, c_ = acc.get(n, (0.0, 0))
        acc[n] = (s_ + b * scale, c_ + 1)
    try:
        d = json.load(open(out_json))
    except Exception:
        d = {}
    d[key] = {k: v[0] / v[1] for k, v in acc.items()}
    d["_source"] = source
    json.dump(d, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--traffic-json")
    ap.add_argument("--key", default="DeepFilterNet3:128x10")
    ap.add_argument("--source", default="ncu --set full --clock-control none, one forward")
    a = ap.parse_args()
    main(a.raw)
    if a.traffic_json:
        traffic(a.raw, a.traffic_json, a.key, a.source)
