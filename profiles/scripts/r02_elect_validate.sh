# Validation of the elected-lane MMA issue loop of k_gru_tc (operands in uniform registers), the H = 512 W_lo column change
# and the L2 prefetch of the projection rows: whole GPU suite, bench configs 2 / 4, timeline, clock64 anatomy.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --extra none --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/b_el.json 2> gpurun_out/b_el.err
timeout 300 python bench.py --config 4 --extra none --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/b_el_cfg4.json 2> gpurun_out/b_el_cfg4.err
python - <<'PY'
import json
for f in ("b_el", "b_el_cfg4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "rtf1", d["rtf_batch1"], "parity", d["parity"]["rms_vs_oracle_device"], d["parity"]["ok"])
        print("   ", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), d["roofline"]["kernel_ms_per_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
DFB_PROF_TIMELINE=1 DFB_PROF_DETAIL=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra none > gpurun_out/bt.json 2> gpurun_out/bt.err
grep timeline gpurun_out/bt.err > gpurun_out/el_timeline.txt; grep "k_gru_tc" gpurun_out/el_timeline.txt
timeout 120 python tests/gpu_gru_timing.py 16 2 2>&1 | tail -9 | tee gpurun_out/el_timing16.txt
timeout 120 python tests/gpu_gru_timing.py 128 2 2>&1 | tail -9 | tee gpurun_out/el_timing128.txt
timeout 120 python tests/gpu_gru_timing.py 256 2 ll 2>&1 | tail -9 | tee gpurun_out/el_timing256_ll.txt
