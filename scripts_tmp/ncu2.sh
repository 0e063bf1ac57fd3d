export DFB_SERIAL=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_bf16x3|k_grouped_linear" -c 13 -f -o gpurun_out/gl python bench.py --streams 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/gl.ncu-rep --page raw --csv > gpurun_out/gl_raw.csv 2>/dev/null
ncu -i gpurun_out/gl.ncu-rep --page source --csv --kernel-id :::2 > gpurun_out/gemm_src.csv 2>/dev/null
rm -f gpurun_out/dwpw.ncu-rep
DFB_PROF_DETAIL=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bd.json 2>gpurun_out/b.err
python -c "import json; d=json.load(open('gpurun_out/bd.json')); print('detail',d['value'], d['ms_per_step']); print(d['roofline']['kernel_ms_per_step'])"
ls -la gpurun_out | head; tail -2 gpurun_out/ncu_b.log | cut -c1-300
