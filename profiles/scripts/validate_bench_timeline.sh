timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err
python -c "import json; d=json.load(open('gpurun_out/b.json')); print('bench',d['value'], d['ms_per_step'], d['e2e']['value'], d['rtf_batch1']); print(d['roofline']['kernel_ms_per_step'])"
DFB_PROF_TIMELINE=1 DFB_PROF_DETAIL=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bt.json 2>gpurun_out/bt.err
grep timeline gpurun_out/bt.err | head -60
tail -3 gpurun_out/b.err
