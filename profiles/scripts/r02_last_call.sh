# Last GPU call of round 2: the whole GPU suite at the closing commit (log kept as profiles/r02_gputest_head.log) and one
# policy experiment: every recurrence at 32 streams per cluster (DFB_GRU_NS=32: enc 4 + erb 4 + df 4 clusters are co-resident,
# so the next chunk's encoder recurrence can overlap this chunk's decoders) with 2 / 3 / 4 device chunks.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_head.log 2>&1
tail -3 gpurun_out/r02_gputest_head.log
for c in 2 3 4; do
DFB_GRU_NS=32 DFB_DEVICE_CHUNKS=$c timeout 200 python bench.py --extra none --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/ns32_c$c.json 2> gpurun_out/ns32_c$c.err
done
python - <<'PY'
import json
for c in (2, 3, 4):
    try:
        d = json.load(open(f"gpurun_out/ns32_c{c}.json"))
        print("NS32 chunks", c, "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["parity"]["rms_vs_oracle_device"], d["parity"]["ok"], "gru", d["roofline"]["kernel_ms_per_step"].get("k_gru_tc"))
    except Exception as e:
        print(c, "unreadable:", e)
PY
