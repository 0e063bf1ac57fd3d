# Round-2 closing evidence at the benchmarked commit and batch (128 x 10 s DeepFilterNet3).  Run under gpurun; outputs land
# in gpurun_out/ and are condensed into profiles/ in the build container (summarize_ncu.py, make_r02_tables.py).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
# (0) the default bench line (all extra configs, CPU baseline) and the reference arm
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err
# (1) every kernel alone on the GPU, one time chunk
DFB_SERIAL=1 DFB_DEVICE_CHUNKS=1 timeout 300 python bench.py --extra none --no-cpu-baseline --steps 5 > gpurun_out/r02_bench_serial.json 2> gpurun_out/r02_serial.err
# (2) timeline of one step
DFB_PROF_TIMELINE=1 DFB_PROF_DETAIL=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra none > gpurun_out/bt.json 2> gpurun_out/bt.err
grep timeline gpurun_out/bt.err > gpurun_out/r02_timeline.txt
# (3) launch list of the default bench command: every launch of our kernels with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra none > gpurun_out/r02_ncu_l.log 2>&1
# (4) full sections of one forward at 128 streams, one time chunk: L launches per forward, skip the 3 warm-up steps
L=$(DFB_DEVICE_CHUNKS=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra none 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['gpu_launches'])")
echo "launches per forward: $L"
DFB_DEVICE_CHUNKS=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ -s $((3 * L)) -c $L -f -o gpurun_out/r02_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra none > gpurun_out/r02_ncu_f.log 2>&1
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
for k in k_gru_tc k_dwpw_bx k_apply_synthesis k_analysis; do
  ncu -i gpurun_out/r02_full.ncu-rep --page source --csv -k regex:$k > gpurun_out/r02_src_$k.csv 2>/dev/null
done
rm -f gpurun_out/r02_full.ncu-rep
ls -la gpurun_out | grep r02_ | head -20; tail -2 gpurun_out/r02_ncu_f.log | cut -c1-300
