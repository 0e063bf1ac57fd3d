# launch list of the exact bench command (timed steps only are what matters; list everything our kernels)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
# full sections for one forward at 32 streams x 10 s
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 40 -c 40 -f -o gpurun_out/full python bench.py --streams 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/full.ncu-rep --page raw --csv > gpurun_out/full_raw.csv 2>/dev/null
for k in k_conv_in k_mask_out k_df_convp k_grouped_linear k_analysis k_apply_synthesis; do
ncu -i gpurun_out/full.ncu-rep --page source --csv -k regex:$k > gpurun_out/src_$k.csv 2>/dev/null
done
rm -f gpurun_out/full.ncu-rep gpurun_out/gemm.ncu-rep
ls -la gpurun_out | head -30; tail -1 gpurun_out/ncu_f.log | cut -c1-200
