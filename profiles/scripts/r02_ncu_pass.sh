# Round-2 evidence at the benchmarked commit and batch (128 x 10 s DeepFilterNet3, one time chunk so that one forward is
# one launch per layer).  Run under gpurun; outputs land in gpurun_out/ and are condensed into profiles/ here.
set -x
# (1) launch list of the default bench command: every launch of our kernels with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra none > gpurun_out/r02_ncu_l.log 2>&1
# (2) full sections of one forward at 128 streams (skip the 3 warm-up steps = 3 x 31 launches)
DFB_DEVICE_CHUNKS=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ -s 93 -c 31 -f -o gpurun_out/r02_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra none > gpurun_out/r02_ncu_f.log 2>&1
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
for k in k_gl_bx k_gru_tc k_dwpw_bx k_df_convp k_apply_synthesis k_analysis; do
  ncu -i gpurun_out/r02_full.ncu-rep --page source --csv -k regex:$k > gpurun_out/r02_src_$k.csv 2>/dev/null
done
rm -f gpurun_out/r02_full.ncu-rep
ls -la gpurun_out | grep r02_ | head; tail -2 gpurun_out/r02_ncu_f.log | cut -c1-300
