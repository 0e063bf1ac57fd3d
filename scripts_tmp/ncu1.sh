export DFB_PRECISION=fp32+gru_tc+proj_tc+conv_tc
export DFB_SERIAL=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dwpw_bx -c 7 -f -o gpurun_out/dwpw python bench.py --streams 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/dwpw.ncu-rep --page raw --csv > gpurun_out/dwpw_raw.csv 2>/dev/null
ncu -i gpurun_out/dwpw.ncu-rep --page source --csv --kernel-id :::6 > gpurun_out/dwpw_src6.csv 2>/dev/null
ls -la gpurun_out/ | head -20
tail -3 gpurun_out/ncu_b.log
