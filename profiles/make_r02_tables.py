"""Tables of profiles/README.md (round 2) from the committed bench lines.
usage: python profiles/make_r02_tables.py > /tmp/tables.md"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    p = os.path.join(HERE, name)
    return json.load(open(p)) if os.path.isfile(p) else None


def main():
    d = load("r02_bench_n1.json")
    if d:
        print("### Headline line (`r02_bench_n1.json`, default `python bench.py`)\n")
        print(f"* device resident: **{d['value'] / 1e3:.1f} k audio-s/s**, {d['ms_per_step']:.2f} ms / step, {d['gpu_launches'] // d['steps']} launches / step")
        print(f"* end to end (host tensors): **{d['e2e']['value'] / 1e3:.1f} k audio-s/s**, {d['e2e']['ms_per_step']:.2f} ms / step")
        print(f"* RTF batch 1: {d['rtf_batch1']:.6f} device resident, {d['rtf_batch1_e2e']:.6f} host tensors")
        print(f"* parity (4 streams of the timed batch vs the oracle): {d['parity']['rms_vs_oracle_device']:.2e} RMS (bound 1e-4)")
        if d.get("cpu_baseline"):
            print(f"* CPU oracle port: {d['cpu_baseline']['value']:.1f} audio-s/s on {d['cpu_baseline']['cores']} threads ({d['cpu_baseline']['sample']})")
        print(f"* clocks: {d['clocks']}\n")
        print("| cfg | workload | ms / step | audio-s/s | e2e audio-s/s | dominant kernel (frac of peak) | parity RMS |")
        print("|---|---|---|---|---|---|---|")
        rows = [dict(cfg=2, workload=d["config"]["workload"], ms_per_step=d["ms_per_step"], value=d["value"], e2e=d["e2e"], roofline=d["roofline"],
                     parity=d["parity"])] + d["extra"]["configs"]
        for e in rows:
            if "error" in e:
                print(f"| {e['cfg']} | error: {e['error']} | | | | | |")
                continue
            r = e["roofline"]
            print(f"| {e['cfg']} | {e['workload'].split(' (')[0]} | {e['ms_per_step']:.1f} | {e['value'] / 1e3:.1f} k | {e['e2e']['value'] / 1e3:.1f} k | "
                  f"`{r['kernel']}` {r['frac']:.3f} ({r['bound']}) | {e['parity']['rms_vs_oracle_device']:.1e} |")
        print()
    s = load("r02_bench_serial.json")
    if s:
        r = s["roofline"]
        print("### Per-kernel times, each kernel alone on the GPU (`r02_bench_serial.json`: `DFB_SERIAL=1 DFB_DEVICE_CHUNKS=1`, 128 x 10 s)\n")
        print("| kernel | ms / step | fraction of its roofline (HBM 6564.5 GB/s or BF16 1440.7 TF/s sustained, measured) |")
        print("|---|---|---|")
        for k, v in r["kernel_ms_per_step"].items():
            print(f"| `{k}` | {v:.3f} | {r['kernel_frac_of_peak'].get(k)} |")
        print(f"\nsum {sum(r['kernel_ms_per_step'].values()):.2f} ms; step (serial) {s['ms_per_step']:.2f} ms\n")


if __name__ == "__main__":
    main()
