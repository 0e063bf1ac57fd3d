#!/usr/bin/env python
"""bench.py -- 48 kHz audio-seconds enhanced per wall-second (batched enhance()), BASELINE.json's
metric, on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole enhancement path (pad -> STFT -> features -> DNN -> mask + deep
filter -> ISTFT -> crop) over one batch of synthetic noisy streams.  Workload at every N:
BASELINE.json configs[1] per GPU -- DeepFilterNet3, 128 streams x 10 s @ 48 kHz -- i.e. weak scaling,
streams sharded across ranks with no data-path collective.

  value  : whole-job audio-s/s with the noisy batch already resident in HBM (device-pointer C ABI)
  e2e    : the same through the reference-facing call enhance(model, df_state, cpu_tensor) with
           pinned HOST buffers, H2D / D2H copies inside the timed region
  roofline: dominant kernel, CUDA-event timed on its launching stream during the timed steps
  cpu_baseline: the CPU oracle port (C DSP + torch-CPU DNN) on this box's host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SR, SECONDS, STREAMS_PER_GPU = 48000, 10, 128
# Algorithmic work per frame per stream (SURVEY.md 8d / BASELINE.md 2)
FLOP_PER_FRAME = {"DeepFilterNet3": 6_614_784, "DeepFilterNet2": 6_956_800, "DeepFilterNet3_ll": 21_846_784}
# Per-kernel algorithmic figures per frame per stream for the roofline line ("Kernels")
KERNEL_MODEL = {
    # name: (bound, algorithmic unit per frame per stream, "bytes"|"flops")   -- DESIGN.md section 4
    "k_analysis": ("hbm", 1920 + 3848 + 128, "bytes"),
    "k_feat_norm": ("hbm", 128 + 768 + 128 + 768, "bytes"),
    "k_apply_synthesis": ("hbm", 3848 + 128 + 3840 + 1920, "bytes"),
    "k_gru_tc": ("tensor", 2 * 5 * 256 * 768, "flops"),                      # 5 GRU layers, W_hh h (fp32-equivalent flops)
    "k_gru": ("tensor", 2 * 5 * 256 * 768, "flops"),
    "k_gemm_bf16x3[gru_proj]": ("tensor", 2 * 5 * 256 * 768, "flops"),      # 5 GRU layers, W_ih x
    "k_grouped_linear[gru_proj]": ("tensor", 2 * 5 * 256 * 768, "flops"),
    "k_grouped_linear": ("tensor", 2 * (3072 * 16 + 512 * 16 + 256 * 32 + 512 * 16 + 256 * 32 + 512 * 32 + 512 * 16 + 256 * 60), "flops"),
    "k_gl_ws": ("tensor", 2 * (3072 * 16 + 512 * 16 + 256 * 32 + 512 * 16 + 256 * 32 + 512 * 32 + 512 * 16 + 256 * 60), "flops"),  # experimental GL
    "k_dwpw": ("tensor", 2 * 64 * 64 * (16 + 8 + 8 + 48 + 8 + 16 + 32), "flops"),
    # fused depthwise -> tcgen05 1x1: rows read (input + pathway) + rows written, 256 B each, over the 7 separable blocks
    "k_dwpw_bx": ("hbm", 256 * ((96 + 48) + (32 + 16) + (16 + 8) + (8 + 8) + (8 + 8 + 8) + (8 + 8 + 16) + (16 + 16 + 32)), "bytes"),
    "k_conv_in[df_conv0]": ("hbm", 96 * 8 + 96 * 256, "bytes"),             # reads feat_spec, writes c0
    "k_conv_in[erb_conv0]": ("hbm", 128 + 32 * 256, "bytes"),
    "k_df_convp": ("hbm", 96 * 256 + 96 * 40, "bytes"),                     # reads c0 once, writes the pathway term of coefs
    "k_mask_out": ("hbm", 2 * 32 * 256 + 128, "bytes"),
}


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the `ncu --set full` capture of one
# forward at 32 streams x 10 s (profiles/r01_ncu_summary_final.md); every kernel's traffic is proportional to the
# number of streams, so `roofline.traffic` scales these by streams / 32.
NCU_TRAFFIC_32 = {
    "k_gru_tc": 99.5e6 + 19.2e6, "k_gemm_bf16x3[gru_proj]": 33.7e6 + 40.4e6, "k_analysis": 61.6e6 + 71.9e6,
    "k_apply_synthesis": 272.5e6 + 46.7e6, "k_df_convp": 793.2e6 + 115.1e6, "k_conv_in[df_conv0]": 24.7e6 + 728.5e6,
    "k_conv_in[erb_conv0]": 4.1e6 + 203.5e6, "k_feat_norm": 32.6e6 + 0.5e6,
}


def model_config(name: str):
    from deepfilternet_b200.config import ModelConfig, load_config
    p = os.path.join(ROOT, "models", "_ref", name, "config.ini")
    if os.path.isfile(p):
        return load_config(p, env={})
    if name == "DeepFilterNet3":
        return ModelConfig(model="deepfilternet3", conv_ch=64, conv_lookahead=2, df_lookahead=2, emb_num_layers=3,
                           df_num_layers=2, lin_groups=16, enc_lin_groups=32, df_gru_skip="groupedlinear",
                           df_pathway_kernel_size_t=5)
    raise SystemExit(f"no config for {name}")


def load_weights(name: str, cfg):
    """Pretrained weights when models/_ref travelled with the snapshot, else random-init weights of
    the same architecture (there is no network for checkpoints)."""
    from deepfilternet_b200.model import find_checkpoint, load_state_dict_file
    from deepfilternet_b200.weights import random_state_dict
    d = os.path.join(ROOT, "models", "_ref", name, "checkpoints")
    if os.path.isdir(d):
        p, _ = find_checkpoint(d)
        if p:
            return load_state_dict_file(p), "pretrained"
    return random_state_dict(cfg, seed=0), "random-init"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  The sampler is started before the warm-up (the tool
    needs ~100 ms to come up and the timed region of 10 steps is only ~140 ms long); `window()` brackets the timed
    regions with wall-clock stamps and only samples inside the window are reported."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def window_begin(self):
        self.t0 = time.time()

    def window_end(self):
        self.t1 = time.time()

    @staticmethod
    def parse(lines, t0=None, t1=None) -> dict:
        def collect(filtered):
            sm, mx, reasons = [], None, set()
            for ts, ln in lines:
                if filtered and t0 is not None and t1 is not None and not (t0 <= ts <= t1 + 0.06):
                    continue
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 10:
                    continue
                try:
                    sm.append(float(f[2])); mx = float(f[3])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        sm, mx, reasons = collect(True)
        scope = "timed regions"
        if not sm:  # nothing landed inside the window: fall back to everything sampled under load (warm-up included)
            sm, mx, reasons = collect(False)
            scope = "warm-up + timed regions"
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "scope": scope}

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return self.parse(self.lines, self.t0, self.t1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def host_threads() -> int:
    """Host cores this process may actually use (affinity mask and cgroup CPU quota respected)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        pass
    return max(1, n)


def cpu_reference_run(cfg, sd, steps: int, warmup: int, threads: int, budget_s: float):
    """The reference's CPU path restated (oracle/): C DSP + torch-CPU DNN with `threads` host
    threads.  The per-step sample (whole 10 s streams of the same synthetic workload) is sized from
    a 1-stream probe so that the run stays near `budget_s` seconds."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dfnet_oracle
    import libdf_oracle
    from tests_common import synth_audio
    libdf_oracle.build()
    torch.set_num_threads(threads)
    cfgd = cfg.as_dict()
    probe = synth_audio(1, SR * SECONDS, seed=1234)
    dfnet_oracle.enhance(sd, cfgd, probe[:, : SR * 2])
    t0 = time.perf_counter()
    dfnet_oracle.enhance(sd, cfgd, probe)
    t_probe = time.perf_counter() - t0
    streams = int(max(1, min(32, budget_s / max(steps + warmup, 1) / max(t_probe, 1e-3))))
    audio = synth_audio(streams, SR * SECONDS, seed=1234)
    for _ in range(warmup):
        dfnet_oracle.enhance(sd, cfgd, audio)
    t0 = time.perf_counter()
    for _ in range(steps):
        dfnet_oracle.enhance(sd, cfgd, audio)
    dt = time.perf_counter() - t0
    return streams * SECONDS * steps / dt, dt / steps, streams


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="DeepFilterNet3")
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU")
    ap.add_argument("--seconds", type=int, default=SECONDS)
    ap.add_argument("--roofline-kernel", default=None, help="kernel to report (default: the one with most time)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = a.gpus
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    cfg = model_config(a.model)
    sd, weights_kind = load_weights(a.model, cfg)
    frames = (SR * a.seconds + cfg.fft_size) // cfg.hop_size
    config = {"workload": f"{a.model}, batch={a.streams} x {a.seconds} s 48 kHz synthetic noisy streams per GPU "
                          f"(BASELINE.json configs[1]), pad=True, {frames} frames/stream",
              "streams_per_gpu": a.streams, "seconds": a.seconds, "frames_per_stream": frames,
              "global_streams": a.streams * n_gpus, "parallelism": f"stream-sharded x{n_gpus}, no collective",
              "weights": weights_kind,
              "l2": "inputs larger than L2 (batch audio 245 MB + >9 GB of activations per step)"}
    threads = host_threads()

    if a.impl == "reference":
        if rank != 0:
            return
        v, s_per_step, sample_streams = cpu_reference_run(cfg, sd, a.steps, min(a.warmup, 1), threads, budget_s=90.0)
        line = {"impl": "reference", "metric": "48kHz audio-sec/sec (batched enhance)", "value": v,
                "unit": "audio-s/s", "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                 "sample": f"{sample_streams} streams x {a.seconds} s per step (CPU oracle port: "
                                           "C libDF restatement + torch-CPU DfNet, all host threads)"},
                "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from deepfilternet_b200 import DfNet, _lib, enhance, enhance_device, libdf
    from tests_common import synth_audio

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    dev_index = local_rank % torch.cuda.device_count()
    torch.cuda.set_device(dev_index)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev_index))
    L = _lib.lib()
    st = libdf.DF(cfg.sr, cfg.fft_size, cfg.hop_size, cfg.nb_erb, cfg.min_nb_erb_freqs, device=dev_index)
    model = DfNet(cfg, sd, st, device=dev_index)
    # this rank's shard of the global batch: streams [rank * streams, (rank + 1) * streams)
    T = SR * a.seconds
    audio = synth_audio(a.streams, T, seed=1234 + rank * a.streams, device=f"cuda:{dev_index}")
    out = torch.empty_like(audio)
    host_in = audio.cpu().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident timing
    clocks = ClockSampler(dev_index)
    clocks.start()
    for _ in range(a.warmup):
        enhance_device(model, st, audio, out=out)
    sync_all()
    L.dfb_profile_report(ctypes_buf(), 1 << 16)  # drain
    L.dfb_profile_enable(1, None)
    clocks.window_begin()
    launches0 = L.dfb_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(a.steps):
        enhance_device(model, st, audio, out=out)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = int(L.dfb_kernel_launches() - launches0)
    buf = ctypes_buf()
    n = L.dfb_profile_report(buf, 1 << 16)
    L.dfb_profile_enable(0, None)
    prof = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, tot = ln.rsplit(" ", 2)
        prof[name] = (int(cnt), float(tot))
    # ---------------- end to end through the public API with host buffers
    for _ in range(2):
        enhance(model, st, host_in, out=host_out)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        enhance(model, st, host_in, out=host_out)   # synchronous: returns with the result on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks.window_end()
    clk = clocks.stop()
    times = torch.tensor([ms / 1e3, e2e_s], dtype=torch.float64, device=f"cuda:{dev_index}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(times[0]), float(times[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_audio_s = a.streams * n_gpus * a.seconds * a.steps
    value = total_audio_s / t_dev
    e2e = total_audio_s / t_e2e
    hbm, tf_burst, tf_sust, peak_src = peaks()
    # roofline of the dominant kernel (share of the step from the event timings)
    total_prof_ms = sum(v[1] for v in prof.values()) or 1.0
    kname = a.roofline_kernel or max(prof, key=lambda k: prof[k][1])
    cnt, tot_ms = prof.get(kname, (0, 0.0))
    bound, per_frame, kind = KERNEL_MODEL.get(kname, ("hbm", 0, "bytes"))
    frames_per_step = a.streams * frames
    work_per_step = per_frame * frames_per_step          # algorithmic bytes / flops of this kernel per step
    launches_per_step = cnt / a.steps if a.steps else 0
    avg_launch_s = (tot_ms / 1e3) / cnt if cnt else float("nan")
    per_launch = work_per_step / launches_per_step if launches_per_step else 0.0
    if kind == "bytes":
        achieved, peak, unit = per_launch / avg_launch_s / 1e9, hbm, "GB/s"
    else:
        achieved, peak, unit = per_launch / avg_launch_s / 1e12, tf_sust, "TFLOP/s"
    roofline = {"kernel": kname, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak if peak else None,
                "traffic": (NCU_TRAFFIC_32[kname] * a.streams / 32.0) if kname in NCU_TRAFFIC_32 else None,
                "traffic_source": "ncu --set full at 32 streams (profiles/r01_ncu_summary_final.md), scaled by streams / 32",
                "peak_source": peak_src,
                "launches_per_step": launches_per_step, "avg_launch_ms": avg_launch_s * 1e3,
                "share_of_step": tot_ms / total_prof_ms,
                "kernel_ms_per_step": {k: round(v[1] / a.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                "end_to_end_tensor_frac": (value * 100 * FLOP_PER_FRAME.get(a.model, 0) / n_gpus) / (tf_sust * 1e12)}
    line = {"metric": "48kHz audio-sec/sec (batched enhance)", "value": value, "unit": "audio-s/s", "n_gpus": n_gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": t_dev * 1e3 / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": int(host_in.numel() * 4),
                    "d2h_bytes_per_step": int(host_out.numel() * 4), "ms_per_step": t_e2e * 1e3 / a.steps},
            "rtf_batch1": None, "roofline": roofline}
    # RTF at batch = 1 (BASELINE.json metric, second half): one 10 s stream, device resident
    a1 = audio[:1].contiguous()
    o1 = torch.empty_like(a1)
    for _ in range(3):
        enhance_device(model, st, a1, out=o1)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(5):
        enhance_device(model, st, a1, out=o1)
    ev1.record()
    torch.cuda.synchronize()
    line["rtf_batch1"] = (ev0.elapsed_time(ev1) / 5 / 1e3) / a.seconds
    if n_gpus == 1 and not a.no_cpu_baseline:
        v, s_step, ns = cpu_reference_run(cfg, sd, 2, 0, threads, budget_s=25.0)
        line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                "sample": f"2 steps of {ns} streams x {a.seconds} s (CPU oracle port: C libDF restatement "
                                          f"+ torch-CPU DfNet, {threads} threads), {s_step:.2f} s/step"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ctypes_buf():
    import ctypes
    return ctypes.create_string_buffer(1 << 16)


if __name__ == "__main__":
    main()
