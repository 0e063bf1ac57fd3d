export DFB_SERIAL=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_bf16x3" -c 2 -f -o gpurun_out/gemm python bench.py --streams 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/gemm.ncu-rep --page raw --csv > gpurun_out/gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/gemm.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/gemm_src.csv 2>/dev/null
rm -f gpurun_out/gl.ncu-rep
ls -la gpurun_out | head; tail -2 gpurun_out/ncu_b.log | cut -c1-200
