DFB_PROF_TIMELINE=1 DFB_PROF_DETAIL=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bt.json 2>gpurun_out/bt.err
grep timeline gpurun_out/bt.err | head -60
