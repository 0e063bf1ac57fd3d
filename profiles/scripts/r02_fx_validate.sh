# Validation of the fused-projection recurrence (k_gru_fx) on one B200: the whole GPU suite, the bench with it (device
# chunk sweep), the clock64 anatomy and the per-launch timeline.  DFB_GRU_FX=0 is the projection GEMM + k_gru_tc pair.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --extra none --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/b_fx1.json 2> gpurun_out/b_fx1.err
for c in 1 3; do
DFB_DEVICE_CHUNKS=$c timeout 300 python bench.py --extra none --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/b_fx1_c$c.json 2> gpurun_out/b_fx1_c$c.err
done
DFB_WIDE_BRANCH=none timeout 300 python bench.py --extra none --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/b_fx1_wnone.json 2> gpurun_out/b_fx1_wnone.err
DFB_WIDE_BRANCH=both timeout 300 python bench.py --extra none --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/b_fx1_wboth.json 2> gpurun_out/b_fx1_wboth.err
python - <<'PY'
import json
for f in ("b_fx1", "b_fx1_c1", "b_fx1_c3", "b_fx1_wnone", "b_fx1_wboth"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "rtf1", d["rtf_batch1"], "parity", d["parity"]["rms_vs_oracle_device"], d["parity"]["ok"])
        print("   ", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), d["roofline"]["kernel_ms_per_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 120 python tests/gpu_gru_timing.py 128 2 > gpurun_out/fx_timing128.txt 2>&1; cat gpurun_out/fx_timing128.txt | tail -12
timeout 120 python tests/gpu_gru_timing.py 16 2 > gpurun_out/fx_timing16.txt 2>&1; cat gpurun_out/fx_timing16.txt | tail -12
DFB_PROF_TIMELINE=1 DFB_PROF_DETAIL=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra none > gpurun_out/bt.json 2> gpurun_out/bt.err
grep timeline gpurun_out/bt.err > gpurun_out/fx_timeline.txt; wc -l gpurun_out/fx_timeline.txt
