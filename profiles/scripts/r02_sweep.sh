# Schedule sweep after the faster recurrence step: per-launch timeline of the default schedule, device chunk count, which
# decoder branch runs 32 streams per cluster (128 x 10 s DeepFilterNet3).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
DFB_PROF_TIMELINE=1 DFB_PROF_DETAIL=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra none > gpurun_out/bt.json 2> gpurun_out/bt.err
grep timeline gpurun_out/bt.err > gpurun_out/el_timeline.txt; wc -l gpurun_out/el_timeline.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --extra none --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; }
run c1 DFB_DEVICE_CHUNKS=1
run c3 DFB_DEVICE_CHUNKS=3
run c4 DFB_DEVICE_CHUNKS=4
run wnone DFB_WIDE_BRANCH=none
run werb DFB_WIDE_BRANCH=erb
run wboth DFB_WIDE_BRANCH=both
run c1wboth DFB_DEVICE_CHUNKS=1 DFB_WIDE_BRANCH=both
run c3wboth DFB_DEVICE_CHUNKS=3 DFB_WIDE_BRANCH=both
python - <<'PY'
import json
for f in ("c1", "c3", "c4", "wnone", "werb", "wboth", "c1wboth", "c3wboth"):
    try:
        d = json.load(open(f"gpurun_out/sw_{f}.json"))
        print(f, "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "gru", d["roofline"]["kernel_ms_per_step"].get("k_gru_tc"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
