timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err
python -c "import json; d=json.load(open('gpurun_out/b.json')); print('bench',d['value'], d['ms_per_step'], d['e2e']['value'], d['rtf_batch1']); print(d['roofline']['kernel_ms_per_step'])"
DFB_SERIAL=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bs.json 2>gpurun_out/b.err
python -c "import json; d=json.load(open('gpurun_out/bs.json')); print('serial',d['value'], d['ms_per_step']); print(d['roofline']['kernel_ms_per_step'])"
tail -3 gpurun_out/b.err
